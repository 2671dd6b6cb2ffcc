#!/bin/bash
# round evidence at the benchmark batch size: launch list, per-launch DRAM traffic / tensor-pipe metrics of the conv
# kernels, one --set full capture of the heaviest launches, and the bench line
mkdir -p gpurun_out
B=${1:-128}
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b$B.csv \
    python scripts/profile_step.py --batch $B --warmup 1 --steps 1 > gpurun_out/ev1.log 2>&1
python scripts/launch_summary.py gpurun_out/launches_b$B.csv | head -12
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum \
    --clock-control none -k regex:"conv_wgrad_kernel|conv_igemm_kernel" --csv --log-file gpurun_out/conv_metrics_b$B.csv \
    python scripts/profile_step.py --batch $B --warmup 0 --steps 1 > gpurun_out/ev2.log 2>&1
tail -1 gpurun_out/ev2.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_kernel" -s 12 -c 2 -f -o gpurun_out/wgrad_full_b$B \
    python scripts/profile_step.py --batch $B --warmup 0 --steps 1 > gpurun_out/ev3.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_kernel" -s 13 -c 2 -f -o gpurun_out/igemm_full_b$B \
    python scripts/profile_step.py --batch $B --warmup 0 --steps 1 > gpurun_out/ev4.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python bench.py --steps 10 --warmup 3 --batch $B > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; python scripts/bench_summary.py gpurun_out/bench_b$B.json
