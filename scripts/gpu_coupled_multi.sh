#!/bin/bash
# coupled SPH-DEM on N GPUs of one box: slab parity test, then STRONG scaling of coupled3d_20m at every N <= #GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "== multi-GPU coupled parity"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -k coupled 2>&1 | tail -15
for N in 1 2 4 8; do
  if [ $N -le $NG ] && [ $N -ge ${NMIN:-1} ]; then
    echo "== bench coupled3d_20m N=$N"
    if [ $N -eq 1 ]; then timeout 900 python bench.py --workload coupled3d_20m --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_coupled_n$N.json 2> gpurun_out/scale_coupled_n$N.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --workload coupled3d_20m --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_coupled_n$N.json 2> gpurun_out/scale_coupled_n$N.err; fi
    echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/scale_coupled_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value %.4g ms/step %.3f particles %d e2e %s stages %s" % (d["value"], d["ms_per_step"], d["config"]["particles"], d["e2e"] and "%.4g"%d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
PY
    tail -3 gpurun_out/scale_coupled_n$N.err | cut -c1-400
  fi
done
