#!/bin/bash
# round 2, second session: final evidence run on one GPU — full GPU suite, same-box A/B (round-1 configuration vs final),
# default bench + reference arm, ncu captures (headline kernel, batched kernel), launch lists, sanitizer passes.
# Numbers printed under ncu / compute-sanitizer are never bench values.
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
echo "== same-box A/B (1e9 x 16x4, 16 queries/step, one query per pass)"
cp quick-adc_b200/libqadc_b200.so build_ab/cur.so
BENCH_ARGS="--no-configs --flat-ring 0 --flat-filter 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1 --flat-filter 1" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 0 --flat-filter 0" STEPS=10 bash tools/gpu_ab.sh cur
BENCH_ARGS="--no-configs --flat-ring 1 --flat-filter 1" STEPS=10 bash tools/gpu_ab.sh cur
rm -f build_ab/cur.so
echo "== default bench"
( time timeout 900 python bench.py ) > gpurun_out/bench_default.log 2>&1
grep '^{' gpurun_out/bench_default.log | tail -1 > gpurun_out/bench_default.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_default.json'))
    print('value %.1f G/s'%(d['value']/1e9),'ms/step %.3f'%d['ms_per_step'],'roof %.3f'%d['roofline']['frac'],'kernel_ms %.3f'%d['roofline']['kernel_ms'],'share %.3f'%d['roofline']['kernel_share_of_step'], 'e2e %.1f'%(d['e2e']['value']/1e9), d['verify']['ok'], d['clocks'], 'batched', d['batched'])
    for k,v in (d.get('configs') or {}).items():
        if isinstance(v, dict):
            print(k, {kk: v[kk] for kk in ('value','queries_per_s','ms','e2e_ms','spot_check_ok','queries_per_pass','stage_ms','scan_kernel_ms') if kk in v}, v.get('roofline',{}).get('frac'))
        else: print(k, v)
    print('cpu', d.get('cpu_baseline')); print('recall', d.get('recall_check'))
except Exception as e:
    print('ERR', e); print(open('gpurun_out/bench_default.log').read()[-3000:])
PY
grep real gpurun_out/bench_default.log
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; grep '^{' gpurun_out/bench_reference.log | cut -c1-700; grep real gpurun_out/bench_reference.log
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/r02b_scan_flat \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02b_ncu_scan.log 2>&1
tail -1 gpurun_out/r02b_ncu_scan.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r02b_launches_bench_1B.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02b_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02b_launches_bench_1B.csv > gpurun_out/r02b_launches_bench_1B.txt; cat gpurun_out/r02b_launches_bench_1B.txt
unset QADC_PROFILE_RANGE
for c in 1 2; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02b_launches_config$c.csv \
    python tools/bench_legs.py $c > gpurun_out/r02b_ncu_list_c$c.log 2>&1
python tools/launch_summary.py gpurun_out/r02b_launches_config$c.csv > gpurun_out/r02b_launches_config$c.txt; cat gpurun_out/r02b_launches_config$c.txt
done
QB=4 tools/gpu_ncu_kernel.sh r02b_c1_batched scan_flat 1 -- python tools/exp_c1.py ncu
echo "== racecheck of every flat-scan variant (qb m ring)"
for cfg in "1 16 1" "2 16 1" "4 16 1" "1 32 1" "2 32 1" "2 16 0"; do
  echo "== qb m ring = $cfg"
  timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_flat.py $cfg 2>&1 | grep -v "^=========     and" | grep "Race reported\|hazard\|RACECHECK\|^ok\|Error\|Warning" | cut -c1-230 | head -6
done > gpurun_out/r02b_racecheck_ring_variants.txt 2>&1
cat gpurun_out/r02b_racecheck_ring_variants.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_memcheck_smoke.txt 2>&1; echo "memcheck smoke exit $?"; tail -2 gpurun_out/r02b_memcheck_smoke.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x --timeout 1200 \
    -k "owner_computes or variants_agree or multi_ivf or test_search_flat_medium or fused" > gpurun_out/r02b_memcheck_tests.txt 2>&1; echo "memcheck tests exit $?"; tail -3 gpurun_out/r02b_memcheck_tests.txt
