#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU parity tests, short + default bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest.log
echo "== bench 1m"; timeout 300 python bench.py --workload wcsph3d_1m --steps 5 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "rc=$?"; cat gpurun_out/bench_1m.json; tail -5 gpurun_out/bench_1m.err
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
