#!/bin/bash
# A/B of library builds under build_ab/ on the 1B bench; NCU=<lib> adds one ncu --set full capture of that build.
# usage: [STEPS=10] [NCU=lib] [TESTS=1] tools/gpu_ab.sh lib1 lib2 ...
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
fi
STEPS=${STEPS:-10}
for lib in "$@"; do
QADC_LIB=$PWD/build_ab/$lib.so timeout 600 python bench.py --steps $STEPS --warmup 3 --no-cpu --verify 1 $BENCH_ARGS > gpurun_out/ab_$lib.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$lib.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('%-10s'%'$lib','value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'batched', d['batched'] and '%.1f'%(d['batched']['value']/1e9), d['verify'] and d['verify']['ok'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
else:
    print('$lib', open('gpurun_out/ab_$lib.log').read()[-1500:])
PY
done
if [ -n "$NCU" ]; then
export QADC_PROFILE_RANGE=1
QADC_LIB=$PWD/build_ab/$NCU.so ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/prof_$NCU \
    python bench.py --steps 1 --warmup 3 --no-cpu --verify 0 $BENCH_ARGS > gpurun_out/ncu_$NCU.log 2>&1
tail -2 gpurun_out/ncu_$NCU.log | cut -c1-200
fi
