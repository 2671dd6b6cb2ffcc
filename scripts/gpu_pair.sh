#!/bin/bash
# group-norm / network / config tests, then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/tests_gpu.log 2>&1; tail -5 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_fold.json 2> gpurun_out/bench_fold.err; tail -c 300 gpurun_out/bench_fold.err; python scripts/bench_summary.py gpurun_out/bench_fold.json
