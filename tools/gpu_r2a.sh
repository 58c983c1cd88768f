#!/bin/bash
# round 2, first GPU call: parity tests, then A/B of the byte-lane pre-filter on the 1B bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for f in 0 1 0 1; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --flat-filter $f > gpurun_out/bench_f$f.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_f$f.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('filter',$f,'value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'e2e %.1f'%(d['e2e']['value']/1e9), 'batched', d['batched'] and '%.1f'%(d['batched']['value']/1e9), d['verify'], d['clocks'])
else:
    print(open('gpurun_out/bench_f$f.log').read()[-2000:])
PY
done
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/prof_scan_r2a \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
