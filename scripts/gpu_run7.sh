#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 10 --force-kernel 2"
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
    else: print(l, end="")
'
for v in "tile_ta=2 tile_g=4" "tile_ta=2 tile_g=3" "tile_ta=2 tile_g=5" "tile_ta=3 tile_g=3" "tile_ta=2 tile_g=4 --opt tile_lcap=96" "tile_ta=2 tile_g=4 --opt tile_lcap=64"; do
  set -- $v
  echo "== $v"; timeout 600 $B --opt $1 --opt $2 $3 $4 2>&1 | python -c "$summ"
done
