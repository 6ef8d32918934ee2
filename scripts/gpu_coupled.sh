#!/bin/bash
# coupled SPH-DEM on the GPU box: parity tests, then bench lines at 2M and 20M
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_coupled.py tests/test_gpu_parity.py -x -q -m gpu -k "coupled or dem" 2>&1 | tail -15 > gpurun_out/pytest_coupled.log
cat gpurun_out/pytest_coupled.log
timeout 300 python bench.py --workload coupled3d_2m --steps 10 --no-cpu-baseline > gpurun_out/bench_coupled3d_2m.json 2> gpurun_out/bench_coupled3d_2m.err
tail -c 1500 gpurun_out/bench_coupled3d_2m.json; tail -5 gpurun_out/bench_coupled3d_2m.err
timeout 600 python bench.py --workload coupled3d_20m --steps 10 > gpurun_out/bench_coupled3d_20m.json 2> gpurun_out/bench_coupled3d_20m.err
tail -c 3000 gpurun_out/bench_coupled3d_20m.json; tail -5 gpurun_out/bench_coupled3d_20m.err
