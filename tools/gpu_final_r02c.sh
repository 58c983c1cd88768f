#!/bin/bash
# round 2, third session: final evidence run on one GPU — GPU suite, default bench + reference arm, ncu capture of the headline
# kernel, launch lists (bench step; rank 0 of an 8-way sharding), racecheck / memcheck of the paths that changed.
# Numbers printed under ncu / compute-sanitizer are never bench values.
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench_default.sh; cp gpurun_out/bench_default.json gpurun_out/r02c_bench_n1.json
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; grep '^{' gpurun_out/bench_reference.log | cut -c1-400; grep real gpurun_out/bench_reference.log
echo "== fixed cost, rank 0 of 8"; timeout 300 python tools/bench_fixed.py 2>&1 | tail -5
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/r02c_scan_flat \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02c_ncu_scan.log 2>&1
tail -1 gpurun_out/r02c_ncu_scan.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r02c_launches_bench_1B.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02c_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02c_launches_bench_1B.csv > gpurun_out/r02c_launches_bench_1B.txt; cat gpurun_out/r02c_launches_bench_1B.txt
unset QADC_PROFILE_RANGE
STEPS=2 OPT=flat_share timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"prefix|tables_kernel|quantize|scan_flat|merge_lists|flat_" -c 400 --csv --log-file gpurun_out/r02c_launches_fixed_g8.csv \
    python tools/bench_fixed.py > gpurun_out/ncu_fixed.log 2>&1
python tools/launch_summary.py gpurun_out/r02c_launches_fixed_g8.csv > gpurun_out/r02c_launches_fixed_g8.txt; cat gpurun_out/r02c_launches_fixed_g8.txt
echo "== racecheck (qb m ring), 4 chunks: the global-histogram path is active"
for cfg in "1 16 1" "4 16 1" "2 32 1"; do
  echo "== qb m ring = $cfg"
  timeout 240 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_flat.py $cfg 2>&1 | grep -v "^=========     and" | grep "Race reported\|hazard\|RACECHECK\|^ok\|Error\|Warning" | cut -c1-230 | head -6
done > gpurun_out/r02c_racecheck_ring_variants.txt 2>&1
cat gpurun_out/r02c_racecheck_ring_variants.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_memcheck_smoke.txt 2>&1; echo "memcheck smoke exit $?"; tail -2 gpurun_out/r02c_memcheck_smoke.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 500 \
    -k "long_prefix or many_full_lists or variants_agree or test_search_flat_medium" > gpurun_out/r02c_memcheck_tests.txt 2>&1; echo "memcheck tests exit $?"; tail -3 gpurun_out/r02c_memcheck_tests.txt
