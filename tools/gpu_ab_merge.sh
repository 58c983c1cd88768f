#!/bin/bash
# A/B of two builds (build_ab/prev.so = previous commit, build_ab/new.so): fixed-cost shard (rank 0 of 8) and the 1e9 bench
mkdir -p gpurun_out
timeout 500 python -X faulthandler -m pytest tests -m gpu -q -x --timeout 400 2>&1 | tail -2
for l in prev new prev new; do
  echo "== $l"; QADC_LIB=$PWD/build_ab/$l.so OPT=flat_share timeout 300 python tools/bench_fixed.py gpurun_out/ab_$l.npz 2>&1 | grep "flat_share=1" | head -1
done
python tools/bench_fixed.py gpurun_out/ab_prev.npz gpurun_out/ab_new.npz
STEPS=10 bash tools/gpu_ab.sh prev new 2>&1 | tail -2
