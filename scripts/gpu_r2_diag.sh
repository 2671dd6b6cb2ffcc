#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_strided.py -m gpu -q --tb=line -x > gpurun_out/tests_ops.log 2>&1; tail -2 gpurun_out/tests_ops.log | cut -c1-300
for v in 1 0 1 0; do echo "halo transposed stores $v"
CB200_HALO_XPOSE=$v timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
done
CB200_HALO_XPOSE=1 timeout 300 python scripts/exp/conv_layer_bench.py 64 112 64 3
CB200_HALO_XPOSE=0 timeout 300 python scripts/exp/conv_layer_bench.py 64 112 64 3
