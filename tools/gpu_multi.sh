#!/bin/bash
# multi-GPU bench: torchrun with N ranks (run under gpurun --gpus N)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
grep '^{' gpurun_out/bench_n$N.log | tail -1 | cut -c1-1500
tail -5 gpurun_out/bench_n$N.log | grep -v '^{' | cut -c1-300
