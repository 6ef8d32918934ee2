#!/bin/bash
# coupled SPH-DEM evidence: launch list of one bench step at 20M, full ncu captures of the contact kernel and the coupled pair kernel
mkdir -p gpurun_out
echo "== launches coupled 20m"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 64 -c 32 --csv --log-file gpurun_out/launches_coupled3d_20m.csv python bench.py --workload coupled3d_20m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full contact kernel 20m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dem_forces -s 3 -c 1 -o gpurun_out/prof_k_dem_forces_nf_20m -f python bench.py --workload coupled3d_20m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full coupled pair kernel 2m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_tiled -s 3 -c 1 -o gpurun_out/prof_k_wcsph_tiled_coupled_2m -f python bench.py --workload coupled3d_2m --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_coupled3d_20m.csv
