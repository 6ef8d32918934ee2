#!/bin/bash
mkdir -p gpurun_out
echo "== pytest dem"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "dem or smoke" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3))
    else: print(l, end="")
'
echo "== dem 1m new"; timeout 600 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_1m 2>&1 | python -c "$summ"
echo "== dem 1m old"; timeout 600 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_1m --opt dem_kernel=0 2>&1 | python -c "$summ"
