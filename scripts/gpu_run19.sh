#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "wcsph or golden or conservation or async or counting" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    else: print(l, end="")
'
for v in "global_lists=1" "global_lists=0" "global_lists=1 --opt tile_lcap=96" "global_lists=1 --opt tile_g=5" "global_lists=1 --real f32"; do
  echo "== $v"; timeout 600 $B --opt $v 2>&1 | python -c "$summ"
done
echo "== ncu full 1m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_tiled -s 3 -c 1 -o gpurun_out/prof_lists_1m -f python bench.py --workload wcsph3d_1m --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
