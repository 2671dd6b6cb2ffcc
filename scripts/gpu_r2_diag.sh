#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -k "first_layer" > gpurun_out/tests_first.log 2>&1; tail -3 gpurun_out/tests_first.log | cut -c1-300
timeout 300 python scripts/exp/first_layer_bench.py 2>&1 | tail -1
CB200_NO_TMA_STORE=1 timeout 300 python scripts/exp/first_layer_bench.py 2>&1 | tail -1
