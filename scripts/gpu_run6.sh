#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],4), "stepfrac", round(d["roofline"]["step"]["frac"],4))
    else: print(l, end="")
'
for v in "2 tile_g=0" "2 tile_g=3" "2 tile_g=4" "2 tile_g=2"; do
  set -- $v
  echo "== wcsph 10m force_kernel=$1 $2"; timeout 600 $B --force-kernel $1 --opt $2 2>&1 | python -c "$summ"
done
echo "== f32 v2"; timeout 600 $B --force-kernel 2 --real f32 2>&1 | python -c "$summ"
echo "== ncu full 1m v2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_tiled -s 3 -c 1 -o gpurun_out/prof_lists_1m -f python bench.py --workload wcsph3d_1m --force-kernel 2 --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
