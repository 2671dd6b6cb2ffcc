#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_strided.py tests/test_gpu_wholemap.py tests/test_gpu_gn_epilogue.py -m gpu -q --tb=line -x > gpurun_out/tests_ops.log 2>&1; tail -3 gpurun_out/tests_ops.log | cut -c1-300
for w in 1 0; do echo "wide stores $w"
CB200_WIDE_STORE=$w timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
CB200_WIDE_STORE=$w timeout 300 python scripts/exp/conv_layer_bench.py 64 112 128 3
CB200_WIDE_STORE=$w timeout 300 python scripts/exp/conv_layer_bench.py 128 112 64 1
CB200_WIDE_STORE=$w timeout 300 python scripts/exp/conv_layer_bench.py 256 28 512 3
done
