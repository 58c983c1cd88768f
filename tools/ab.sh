#!/bin/bash
# A/B several builds of the library on the 1B bench (qb=1). usage: tools_ab.sh lib1 lib2 ...
mkdir -p gpurun_out
for lib in "$@"; do
QADC_LIB=$PWD/$lib timeout 600 python bench.py --n-vectors ${N:-1000000000} --steps ${STEPS:-3} --warmup 3 --no-cpu --qb ${QB:-1} > gpurun_out/ab.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/ab.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$lib','value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'])
else:
    print(open('gpurun_out/ab.log').read()[-1500:])
PY
done
