#!/bin/bash
# launch list (per-launch device time) of one training step + one full ncu capture of the top kernels
mkdir -p gpurun_out
B=${1:-32}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --batch $B --warmup 1 --steps 1 > gpurun_out/prof_launch.log 2>&1
tail -2 gpurun_out/prof_launch.log
timeout 900 python bench.py --steps 5 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; tail -c 600 gpurun_out/bench_b$B.json
