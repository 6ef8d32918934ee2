#!/bin/bash
# 2-GPU check: slab parity tests (halo, migration, coupled), then 2-GPU bench lines (WCSPH weak, coupled strong)
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -25
N=2
for WL in wcsph3d_10m coupled3d_20m; do
for IMPL in ${IMPLS:-1}; do
echo "== bench $WL N=$N halo_impl=$IMPL"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --workload $WL --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --opt halo_impl=$IMPL 2> gpurun_out/bench_n$N.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g ms/step %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})
"; tail -3 gpurun_out/bench_n$N.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-300
done; done
