#!/bin/bash
# round 2 final evidence run (one GPU): full GPU suite, same-box A/B of the round's kernel work, default bench, ncu captures.
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
echo "== same-box A/B (1e9 x 16x4, 16 queries/step, QB=1): round-1 configuration (CTA ring, exact core) vs pre-filter vs final"
BENCH_ARGS="--no-configs --flat-ring 0 --flat-filter 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 0 --flat-filter 1" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1 --flat-filter 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1 --flat-filter 1" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 0 --flat-filter 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1 --flat-filter 1" STEPS=10 bash tools/gpu_ab.sh cur
echo "== default bench"
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1
grep '^{' gpurun_out/bench_default.log | tail -1 > gpurun_out/bench_default.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'],'share %.3f'%d['roofline']['kernel_share_of_step'], 'e2e %.1f'%(d['e2e']['value']/1e9), d['verify']['ok'], d['clocks'])
    for k,v in (d.get('configs') or {}).items():
        if isinstance(v, dict):
            print(k, {kk: v[kk] for kk in ('value','queries_per_s','ms','e2e_ms','spot_check_ok','queries_per_pass','stage_ms','scan_kernel_ms') if kk in v}, v.get('roofline',{}).get('frac'))
        else: print(k, v)
    print('cpu', d.get('cpu_baseline')); print('recall', d.get('recall_check'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_default.log').read()[-3000:])
PY
grep real gpurun_out/bench_default.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; grep '^{' gpurun_out/bench_reference.log | cut -c1-600; grep real gpurun_out/bench_reference.log
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/r02_scan_flat \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02_ncu_scan.log 2>&1
tail -1 gpurun_out/r02_ncu_scan.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r02_launches_bench_1B.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench_1B.csv > gpurun_out/r02_launches_bench_1B.txt; cat gpurun_out/r02_launches_bench_1B.txt
