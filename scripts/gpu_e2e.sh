#!/bin/bash
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "e2e", d["e2e"] and ("%.4g" % d["e2e"]["value"], round(d["e2e"]["ms_per_step"],2)))
    else: print(l, end="")
'
for i in 1 2; do
echo "== wcsph 10m"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | python -c "$summ"
echo "== dem 1m"; timeout 600 python bench.py --no-cpu-baseline --workload dem3d_1m 2>&1 | python -c "$summ"
done
echo "== 2d"; timeout 600 python bench.py --no-cpu-baseline --workload wcsph2d_20k 2>&1 | python -c "$summ"
