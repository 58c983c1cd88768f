#!/bin/bash
# round 2, third session, last run: GPU suite (incl. the 16-bit plain ADC), default bench, racecheck of the ring kernels
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/gpu_bench_default.sh 2>&1 | head -3; cp gpurun_out/bench_default.json gpurun_out/r02c_bench_n1_b.json
echo "== fixed cost, rank 0 of 8"; timeout 300 python tools/bench_fixed.py 2>&1 | tail -5
echo "== racecheck (qb m ring), 4 chunks: the global-histogram path is active"
for cfg in "1 16 1" "4 16 1" "2 32 1"; do
  echo "== qb m ring = $cfg"
  timeout 240 compute-sanitizer --tool racecheck --racecheck-report all python tools/racecheck_flat.py $cfg 2>&1 | grep -v "^=========     and" | grep "Race reported\|hazard\|RACECHECK\|^ok\|Error\|Warning" | cut -c1-230 | head -6
done > gpurun_out/r02c_racecheck_ring_variants.txt 2>&1
cat gpurun_out/r02c_racecheck_ring_variants.txt
