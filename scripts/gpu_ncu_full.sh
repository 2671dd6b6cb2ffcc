#!/bin/bash
# full ncu capture of the first launches of one kernel family inside a training step
mkdir -p gpurun_out
KERN=${1:-conv_igemm_kernel}; SKIP=${2:-0}; COUNT=${3:-3}; B=${4:-32}; OUT=${5:-prof_$KERN}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KERN -s $SKIP -c $COUNT -f -o gpurun_out/$OUT \
    python scripts/profile_step.py --batch $B --warmup 0 --steps 1 > gpurun_out/ncu_full_$OUT.log 2>&1
tail -3 gpurun_out/ncu_full_$OUT.log; ls -la gpurun_out/*.ncu-rep
