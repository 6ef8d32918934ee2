#!/bin/bash
# Round-1 evidence: full default bench lines, per-launch list and full ncu capture of the dominant kernels at bench size.
mkdir -p gpurun_out
echo "== bench default (10M WCSPH f64)"; timeout 900 python bench.py > gpurun_out/bench_wcsph3d_10m.json 2> gpurun_out/bench.err; echo rc=$?; cut -c1-600 gpurun_out/bench_wcsph3d_10m.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo rc=$?; cut -c1-400 gpurun_out/bench_reference.json
echo "== bench dem"; timeout 900 python bench.py --workload dem3d_1m > gpurun_out/bench_dem3d_1m.json 2>> gpurun_out/bench.err; echo rc=$?; cut -c1-300 gpurun_out/bench_dem3d_1m.json
echo "== bench 2d"; timeout 900 python bench.py --workload wcsph2d_20k > gpurun_out/bench_wcsph2d_20k.json 2>> gpurun_out/bench.err; echo rc=$?; cut -c1-300 gpurun_out/bench_wcsph2d_20k.json
echo "== bench f32"; timeout 900 python bench.py --real f32 --no-cpu-baseline > gpurun_out/bench_wcsph3d_10m_f32.json 2>> gpurun_out/bench.err; echo rc=$?; cut -c1-300 gpurun_out/bench_wcsph3d_10m_f32.json
echo "== launches 10m"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_wcsph3d_10m.csv python bench.py --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== launches dem"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/launches_dem3d_1m.csv python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full pair kernel 10m"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_tiled -s 3 -c 1 -o gpurun_out/prof_k_wcsph_tiled_10m -f python bench.py --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full dem 1m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dem_forces -s 3 -c 1 -o gpurun_out/prof_k_dem_forces_1m -f python bench.py --workload dem3d_1m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full permute 10m"
timeout 900 ncu --set full --clock-control none -k regex:k_permute -s 3 -c 1 -o gpurun_out/prof_k_permute_10m -f python bench.py --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
tail -3 gpurun_out/bench.err
