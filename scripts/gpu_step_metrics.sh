#!/bin/bash
# per-launch time / DRAM bytes / tensor-pipe activity / L2 bytes of every kernel of one training step at batch $1
mkdir -p gpurun_out
B=${1:-128}
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/step_metrics_b$B.csv \
    python scripts/profile_step.py --batch $B --warmup 1 --steps 1 > gpurun_out/ev1.log 2>&1
tail -1 gpurun_out/ev1.log
python scripts/step_metrics_summary.py gpurun_out/step_metrics_b$B.csv | head -${2:-40}
