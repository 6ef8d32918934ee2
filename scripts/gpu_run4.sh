#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest.log
B="python bench.py --no-cpu-baseline --steps 10"
summ='
import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l); print("value", "%.4g" % d["value"], "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"],4), "stepfrac", round(d["roofline"]["step"]["frac"],4), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "launches", d["gpu_launches"], "zbar", d["config"]["mean_contacts"])
    else: print(l, end="")
'
echo "== wcsph 10m"; timeout 600 $B 2>&1 | python -c "$summ"
echo "== dem 1m"; timeout 600 $B --workload dem3d_1m 2>&1 | python -c "$summ"
echo "== wcsph2d 20k"; timeout 600 $B --workload wcsph2d_20k 2>&1 | python -c "$summ"
echo "== wcsph 10m f32"; timeout 600 $B --real f32 2>&1 | python -c "$summ"
echo "== ncu launches 10m"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_wcsph3d_10m.csv python bench.py --no-cpu-baseline --no-e2e --steps 4 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu launches dem"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_dem3d_1m.csv python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 4 > gpurun_out/ncu_launches_dem.log 2>&1; echo "rc=$?"
echo "== ncu full dem"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dem_forces -s 3 -c 1 -o gpurun_out/prof_dem_1m -f python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/ncu_full_dem.log 2>&1; echo "rc=$?"
