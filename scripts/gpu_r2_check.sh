#!/bin/bash
# round-2 check: new back-end tests first (verbose), then the full GPU suite, smoke, bench line at batch 128
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backends.py tests/test_gpu_shuffle.py tests/test_gpu_dropout.py -m gpu -q --tb=short -x --maxfail=12 > gpurun_out/tests_backends.log 2>&1; tail -40 gpurun_out/tests_backends.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=40 --deselect tests/test_gpu_backends.py --deselect tests/test_gpu_shuffle.py --deselect tests/test_gpu_dropout.py > gpurun_out/tests_gpu.log 2>&1; tail -25 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b128.json 2> gpurun_out/bench_b128.err; tail -c 1500 gpurun_out/bench_b128.json; tail -5 gpurun_out/bench_b128.err
