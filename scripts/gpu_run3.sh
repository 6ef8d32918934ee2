#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
B="python bench.py --no-cpu-baseline --no-e2e --steps 10"
for v in "1 tile_g=3"; do
  set -- $v
  echo "== bench 10m force_kernel=$1 $2"
  timeout 600 $B --force-kernel $1 --opt $2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('ms/step', round(d['ms_per_step'],3), 'stages', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'clk', d['clocks'])
    else: print(l, end='')
"
done
echo "== f32"; timeout 600 $B --real f32 | cut -c1-120
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_wcsph_cellwarp -s 3 -c 1 -o gpurun_out/prof_cellwarp_1m -f python bench.py --workload wcsph3d_1m --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
