#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== multi-GPU parity"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30
N=${NGPU:-2}
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-1500; tail -5 gpurun_out/bench_n$N.err
echo "== bench N=1"; timeout 600 python bench.py --steps 10 --no-cpu-baseline | cut -c1-300
