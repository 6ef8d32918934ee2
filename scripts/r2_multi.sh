#!/bin/bash
# Round 2 multi-GPU call: slab parity tests at every world size the box offers, then the bench line of the default (strong
# scaling of the 80 M tank) at N = $1 GPUs.   gpurun --gpus N -- ./scripts/r2_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r2_topo_n$N.txt
echo "== multi-GPU parity tests on $N GPUs"
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -rs 2>&1 | tail -25 | tee gpurun_out/r2_multigpu_parity_n$N.txt
echo "== bench N=$N (strong scaling, 80 M tank)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/r2_scale_n$N.out 2> gpurun_out/r2_scale_n$N.err
grep '^{' gpurun_out/r2_scale_n$N.out > gpurun_out/r2_scale_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_scale_n$N.json"))
print("value %.4g  ms/step %.3f  e2e %.4g (%.1f ms)  moving %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else 0, d["e2e"]["ms_per_step"] if d["e2e"] else 0, d["moving"]["ms_per_step"] if d["moving"] else 0))
print({k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()})
print("parity", d["parity_check"])
PY
tail -3 gpurun_out/r2_scale_n$N.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-400
