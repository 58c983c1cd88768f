#!/bin/bash
# BASELINE config 5 on N GPUs (run under gpurun --gpus N): sharded lists + split coarse assignment.
# usage: tools/gpu_ivf_multi.sh N [n_vectors] [queries] [check]
mkdir -p gpurun_out
N=${1:-2}; NV=${2:-1000000000}; Q=${3:-10000}; CHECK=${4:-0}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/bench_ivf_sharded.py --n-vectors $NV --queries $Q --steps 5 --check $CHECK > gpurun_out/ivf5_n$N.log 2>&1
grep '^{' gpurun_out/ivf5_n$N.log | tail -1 | cut -c1-900
tail -4 gpurun_out/ivf5_n$N.log | grep -v '^{' | cut -c1-300
