#!/bin/bash
mkdir -p gpurun_out
echo skip
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3), "stepfrac", round(d["roofline"]["step"]["frac"],3))
    else: print(l, end="")
'
echo "== dem 1m"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_1m 2>&1 | python -c "$summ"
echo "== dem 8m"; timeout 900 python bench.py --no-cpu-baseline --no-e2e --workload dem3d_8m --steps 10 2>&1 | python -c "$summ"
