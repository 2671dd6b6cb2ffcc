#!/bin/bash
# end-of-round evidence: full GPU suite, smoke, the default bench line, per-launch ncu metrics of one step
mkdir -p gpurun_out
B=${1:-128}
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/tests_gpu.log 2>&1; tail -6 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err; python scripts/bench_summary.py gpurun_out/bench_final.json
bash scripts/gpu_step_metrics.sh $B 45
