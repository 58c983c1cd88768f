#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for qb in 1 2 4; do
timeout 600 python bench.py --n-vectors ${N:-134217728} --steps 3 --warmup 3 --no-cpu --qb $qb > gpurun_out/bench_qb$qb.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_qb$qb.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('qb',$qb,'value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'e2e %.1f'%(d['e2e']['value']/1e9), d['clocks'])
else:
    print(open('gpurun_out/bench_qb$qb.log').read()[-2000:])
PY
done
