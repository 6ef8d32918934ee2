#!/bin/bash
# Round 2, call 1: ncu --set full (with source) of the SHIPPED pair kernel at 10 M f64, launch list, one bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
echo "== bench default"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo rc=$?; cut -c1-400 gpurun_out/r2_bench0.json
echo "== launches 10m"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 30 --csv --log-file gpurun_out/r2_launches_wcsph3d_10m_base.csv python bench.py --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
echo "== ncu full pair kernel 10m"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_tiled -s 3 -c 1 -o gpurun_out/r2_prof_k_wcsph_tiled_10m_base -f python bench.py --no-cpu-baseline --no-e2e --steps 3 > /dev/null 2>&1; echo rc=$?
ls -la gpurun_out | tail -8
