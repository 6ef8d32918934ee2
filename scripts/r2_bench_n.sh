#!/bin/bash
# bench line of the default workload (strong scaling of the 80 M tank) at N = $1 GPUs -> gpurun_out/r2_scale_n$N.json
N=${1:-2}; shift
mkdir -p gpurun_out
if [ "$N" = "1" ]; then python bench.py "$@" > gpurun_out/r2_scale_n1.out 2> gpurun_out/r2_scale_n1.err
else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > gpurun_out/r2_scale_n$N.out 2> gpurun_out/r2_scale_n$N.err; fi
grep '^{' gpurun_out/r2_scale_n$N.out > gpurun_out/r2_scale_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_scale_n$N.json"))
print("N=$N value %.4g  ms/step %.3f  e2e %.4g (%.1f ms)  moving %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"] if d["e2e"] else 0, d["e2e"]["ms_per_step"] if d["e2e"] else 0, d["moving"]["ms_per_step"] if d["moving"] else 0))
print({k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, "parity", d["parity_check"] and d["parity_check"]["ok"])
PY
tail -2 gpurun_out/r2_scale_n$N.err | grep -v "OMP_NUM\|\*\*\*\*" | cut -c1-300
