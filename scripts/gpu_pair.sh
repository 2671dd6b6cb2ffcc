#!/bin/bash
# first-layer / network parity tests, then bench lines with and without the TMA-store epilogue of the first layer
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_network.py -m gpu -q --tb=short --maxfail=10 -k "first or network or training or darknet" > gpurun_out/tests_first.log 2>&1; tail -6 gpurun_out/tests_first.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_tmastore.json 2> gpurun_out/bench_tmastore.err; tail -c 300 gpurun_out/bench_tmastore.err; python scripts/bench_summary.py gpurun_out/bench_tmastore.json
CB200_NO_TMA_STORE=1 timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_notmastore.json 2> gpurun_out/bench_notmastore.err; python scripts/bench_summary.py gpurun_out/bench_notmastore.json
