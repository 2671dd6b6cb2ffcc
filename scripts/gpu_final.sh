#!/bin/bash
# end-of-round evidence: smoke, the default bench line, per-launch ncu metrics of one step, full ncu capture of the CTA-pair kernel
mkdir -p gpurun_out
B=${1:-128}
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err; python scripts/bench_summary.py gpurun_out/bench_final.json
bash scripts/gpu_step_metrics.sh $B 45
bash scripts/gpu_ncu_full.sh conv_igemm_pair_kernel 4 2 $B igemm_pair_r1
ncu -i gpurun_out/igemm_pair_r1.ncu-rep --page raw --csv > gpurun_out/igemm_pair_full_raw.csv 2>/dev/null; wc -c gpurun_out/igemm_pair_full_raw.csv
