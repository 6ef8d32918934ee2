#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
NG=$(nvidia-smi -L | wc -l)
echo "== multi-GPU parity"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for N in 1 2 4 8; do
  if [ $N -le $NG ]; then
    echo "== bench N=$N"
    if [ $N -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
    else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; fi
    echo "rc=$?"; python - <<PY
import json
for l in open("gpurun_out/scale_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N value %.4g ms/step %.3f particles %d e2e %s stages %s" % (d["value"], d["ms_per_step"], d["config"]["particles"], d["e2e"] and "%.4g"%d["e2e"]["value"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
PY
    tail -2 gpurun_out/scale_n$N.err | cut -c1-300
  fi
done
