#!/bin/bash
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    else: print(l, end="")
'
for v in "global_lists=1" "global_lists=0"; do
  echo "== $v"; timeout 600 $B --opt $v 2>&1 | python -c "$summ"
done
