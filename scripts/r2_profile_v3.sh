#!/bin/bash
# ncu --set full (with source) of pair-kernel variant 3 at 10 M f64
mkdir -p gpurun_out
OPTS="${R2_OPTS:---force-kernel 3 --opt zsub=4}"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_zrun -s 3 -c 1 -o gpurun_out/r2_prof_k_wcsph_zrun_10m -f python bench.py --no-cpu-baseline --no-e2e --steps 3 $OPTS > gpurun_out/r2_prof_v3.log 2>&1; echo rc=$?
tail -3 gpurun_out/r2_prof_v3.log
