#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=line -k "halo" > gpurun_out/tests_halo.log 2>&1; tail -3 gpurun_out/tests_halo.log | cut -c1-300
timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
timeout 300 python scripts/exp/conv_layer_bench.py 64 112 128 3
