#!/bin/bash
# full GPU suite, then bench lines with / without the CTA-pair kernels (same box, back to back)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/tests_gpu.log 2>&1; tail -12 gpurun_out/tests_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; tail -c 300 gpurun_out/bench_pair.err; python scripts/bench_summary.py gpurun_out/bench_pair.json
CB200_CTA_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err; tail -c 300 gpurun_out/bench_nopair.err; python scripts/bench_summary.py gpurun_out/bench_nopair.json
