#!/bin/bash
free -g | head -2
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "particles", d["config"]["particles"], "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
        open("gpurun_out/bench_wcsph3d_80m_1gpu.json","w").write(l)
    else: print(l, end="")
'
mkdir -p gpurun_out
timeout 1500 python bench.py --workload wcsph3d_80m --no-cpu-baseline --no-e2e --steps 5 2>&1 | python -c "$summ"
