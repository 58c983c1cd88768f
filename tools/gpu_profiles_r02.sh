#!/bin/bash
# round 2 evidence: ncu --set full of the dominant kernel + launch lists (bench step, config 2) + sanitizer passes.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export QADC_PROFILE_RANGE=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_flat -c 1 -f -o gpurun_out/r02_scan_flat \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02_ncu_scan.log 2>&1
tail -1 gpurun_out/r02_ncu_scan.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv --log-file gpurun_out/r02_launches_bench_1B.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --verify 0 > gpurun_out/r02_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_bench_1B.csv > gpurun_out/r02_launches_bench_1B.txt; cat gpurun_out/r02_launches_bench_1B.txt
unset QADC_PROFILE_RANGE
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_config2.csv \
    python tools/bench_legs.py 2 > gpurun_out/r02_ncu_list2.log 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_config2.csv > gpurun_out/r02_launches_config2.txt; cat gpurun_out/r02_launches_config2.txt
ncu --set full --clock-control none --import-source on -k regex:"scan_ivf|ivf_prepare|coarse_dist|coarse_select" -s 8 -c 4 -f -o gpurun_out/r02_ivf_kernels \
    python tools/bench_legs.py 2 > gpurun_out/r02_ncu_ivf.log 2>&1
tail -1 gpurun_out/r02_ncu_ivf.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_memcheck_smoke.txt 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r02_memcheck_smoke.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_racecheck_smoke.txt 2>&1; echo "racecheck exit $?"; tail -3 gpurun_out/r02_racecheck_smoke.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x --timeout 1200 \
    -k "fused or warp_ring or multi_ivf or test_search_flat_medium or flat_scan_with_tables_bit_exact" > gpurun_out/r02_racecheck_tests.txt 2>&1; echo "racecheck tests exit $?"; tail -4 gpurun_out/r02_racecheck_tests.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x --timeout 1200 \
    -k "fused or warp_ring or multi_ivf or test_search_flat_medium or add_vectors" > gpurun_out/r02_memcheck_tests.txt 2>&1; echo "memcheck tests exit $?"; tail -4 gpurun_out/r02_memcheck_tests.txt
