#!/bin/bash
# launch list (per-launch device time) of one training step
mkdir -p gpurun_out
B=${1:-32}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python scripts/profile_step.py --batch $B --warmup 1 --steps 1 > gpurun_out/prof_launch.log 2>&1
tail -1 gpurun_out/prof_launch.log
python scripts/launch_summary.py gpurun_out/launches.csv "$2"
