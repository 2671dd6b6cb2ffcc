#!/bin/bash
# full GPU test suite + smoke + launch list with per-launch metrics of every kernel of one step at batch $1
mkdir -p gpurun_out
B=${1:-128}
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=40 > gpurun_out/tests_gpu.log 2>&1; tail -8 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/step_metrics_b$B.csv \
    python scripts/profile_step.py --batch $B --warmup 1 --steps 1 > gpurun_out/ev1.log 2>&1
tail -1 gpurun_out/ev1.log
