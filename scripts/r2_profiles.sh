#!/bin/bash
# Round 2 profiles of the SHIPPED kernels (HEAD library): launch lists of the bench commands, ncu --set full (with source) of
# the pair kernel at 10 M and 80 M, of the DEM contact kernel at 1 M, and of the re-sort kernels on MOVED particles.
# Outputs under gpurun_out/ (copied to profiles/ by hand after reading).  One GPU.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
B="python bench.py --no-cpu-baseline --no-e2e --no-extra --no-parity"
PART=${1:-all}      # gpurun merges at most 64 MiB back: "a" = launch lists + pair kernel 10 M + DEM, "b" = pair kernel 80 M + re-sort kernels
if [ "$PART" != "b" ]; then
echo "== launch list, default bench command (80 M)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_wcsph3d_80m.csv $B --no-moving --steps 2 --warmup 3 > /dev/null 2>&1; echo rc=$?
echo "== launch list, 10 M"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_wcsph3d_10m.csv $B --no-moving --workload wcsph3d_10m --steps 2 --warmup 3 > /dev/null 2>&1; echo rc=$?
echo "== launch list, DEM 1 M"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_dem3d_1m.csv $B --no-moving --workload dem3d_1m --steps 2 --warmup 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full, pair kernel 10 M"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_zrun -s 4 -c 1 -o gpurun_out/r2_ncu_k_wcsph_zrun_10m -f $B --no-moving --workload wcsph3d_10m --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full, DEM contact kernel 1 M"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_dem_forces -s 4 -c 1 -o gpurun_out/r2_ncu_k_dem_forces_1m -f $B --no-moving --workload dem3d_1m --steps 3 > /dev/null 2>&1; echo rc=$?
fi
if [ "$PART" != "a" ]; then
echo "== ncu full, pair kernel 80 M"
timeout 1500 ncu --set full --clock-control none -k regex:k_wcsph_zrun -s 4 -c 1 -o gpurun_out/r2_ncu_k_wcsph_zrun_80m -f $B --no-moving --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full, re-sort kernels on MOVED particles (10 M, inside pst_step)"
timeout 1500 ncu --set full --clock-control none -k regex:"k_permute|k_keys_count|k_place|k_cell_order" --nvtx --nvtx-include "pst_step/" -s 24 -c 4 -o gpurun_out/r2_ncu_resort_moving_10m -f $B --workload wcsph3d_10m --steps 3 > /dev/null 2>&1; echo rc=$?
fi
ls -la gpurun_out/*.ncu-rep
