#!/bin/bash
# one ncu --set full capture of a kernel (regex) of a command.  usage: tools/gpu_ncu_kernel.sh name regex skip -- command...
mkdir -p gpurun_out
NAME=$1; REGEX=$2; SKIP=$3; shift 4
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c 1 -f -o gpurun_out/prof_$NAME "$@" > gpurun_out/ncu_$NAME.log 2>&1
tail -3 gpurun_out/ncu_$NAME.log | cut -c1-200
