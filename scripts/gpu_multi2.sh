#!/bin/bash
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
N=2
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_n$N.err | cut -c1-400; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
