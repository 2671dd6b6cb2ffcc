#!/bin/bash
# conv-level parity tests, then a bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_network.py -m gpu -q --tb=short --maxfail=20 > gpurun_out/tests_gpu2.log 2>&1; tail -6 gpurun_out/tests_gpu2.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_minmax.json 2> gpurun_out/bench_minmax.err; tail -c 300 gpurun_out/bench_minmax.err; python scripts/bench_summary.py gpurun_out/bench_minmax.json
