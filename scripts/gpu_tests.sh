#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=40 > gpurun_out/tests_gpu.log 2>&1; tail -${1:-25} gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
if [ -n "$2" ]; then timeout 900 python bench.py --steps 5 --warmup 3 --batch $2 --no-cpu-baseline > gpurun_out/bench_b$2.json 2> gpurun_out/bench_b$2.err; tail -c 300 gpurun_out/bench_b$2.err; python scripts/bench_summary.py gpurun_out/bench_b$2.json; fi
