#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "launches", d["gpu_launches"])
    else: print(l, end="")
'
for impl in 1 0; do
echo "== wcsph 10m sort_impl=$impl"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --steps 10 --opt sort_impl=$impl 2>&1 | python -c "$summ"
echo "== dem 1m sort_impl=$impl"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_1m --opt sort_impl=$impl 2>&1 | python -c "$summ"
echo "== dem 8m sort_impl=$impl"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_8m --steps 10 --opt sort_impl=$impl 2>&1 | python -c "$summ"
echo "== 2d sort_impl=$impl"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --workload wcsph2d_20k --opt sort_impl=$impl 2>&1 | python -c "$summ"
done
