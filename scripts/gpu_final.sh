#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_wcsph3d_10m.json 2> gpurun_out/bench.err; echo rc=$?; cut -c1-200 gpurun_out/bench_wcsph3d_10m.json
echo "== bench dem"; timeout 900 python bench.py --workload dem3d_1m > gpurun_out/bench_dem3d_1m.json 2>> gpurun_out/bench.err; echo rc=$?
echo "== bench 2d"; timeout 900 python bench.py --workload wcsph2d_20k > gpurun_out/bench_wcsph2d_20k.json 2>> gpurun_out/bench.err; echo rc=$?
echo "== launches dem"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_dem3d_1m.csv python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full dem 1m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dem_forces -s 3 -c 1 -o gpurun_out/prof_k_dem_forces_1m -f python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
python - <<'PY'
import json
for f in ["bench_wcsph3d_10m","bench_dem3d_1m","bench_wcsph2d_20k"]:
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value %.4g"%d["value"], "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e") and "%.4g"%d["e2e"]["value"], "cpu", d.get("cpu_baseline") and ("%.4g"%d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "frac", round(d["roofline"]["frac"],4), "traffic", d["roofline"]["traffic"], "launches", d["gpu_launches"])
PY
