#!/bin/bash
# kernel A/B experiments: bench the 10M WCSPH step with each library build named in $LIBS
mkdir -p gpurun_out
for L in $LIBS; do
  echo "== $L"
  PRESTIGE_B200_LIB=$PWD/prestige_b200/libprestige_b200$L.so timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-e2e $BENCH_ARGS 2> gpurun_out/exp$L.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})
"
  tail -2 gpurun_out/exp$L.err
done
