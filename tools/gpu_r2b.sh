#!/bin/bash
# round 2, second GPU call: new BASELINE-shape parity tests, then A/B of kernel build variants on the 1B bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_baseline_shapes.py -m gpu -q -x --timeout 1200 > gpurun_out/pytest_shapes.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_shapes.log
tail -15 gpurun_out/pytest_shapes.log
STEPS=${STEPS:-5}
for lib in "$@"; do
QADC_LIB=$PWD/build_ab/$lib.so timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu --verify 1 > gpurun_out/ab_$lib.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$lib.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('%-10s'%'$lib','value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'batched', d['batched'] and '%.1f'%(d['batched']['value']/1e9), d['verify']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
else:
    print('$lib', open('gpurun_out/ab_$lib.log').read()[-1500:])
PY
done
