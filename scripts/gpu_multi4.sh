#!/bin/bash
# N-GPU verification of the default (peer-memory) halo: slab parity tests at world 2 and 4, then bench lines at N = #GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8
for WL in wcsph3d_10m coupled3d_20m; do
echo "== bench $WL N=$NG"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --workload $WL --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale2_${WL}_n$NG.json 2> gpurun_out/scale2_n$NG.err
python - <<PY
import json
for l in open("gpurun_out/scale2_${WL}_n$NG.json"):
    if l.startswith("{"):
        d=json.loads(l); print("value %.4g ms/step %.3f particles %d e2e %.4g" % (d["value"], d["ms_per_step"], d["config"]["particles"], d["e2e"]["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()})
PY
tail -3 gpurun_out/scale2_n$NG.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-300
done
