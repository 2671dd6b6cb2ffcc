#!/bin/bash
# full GPU suite (2-rank DP test included when two GPUs are visible), then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/tests_gpu.log 2>&1; tail -5 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err; python scripts/bench_summary.py gpurun_out/bench_final.json
