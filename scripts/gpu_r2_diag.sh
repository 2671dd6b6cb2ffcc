#!/bin/bash
mkdir -p gpurun_out
for v in 1 0 1 0; do echo "bias once $v"
CB200_BIAS_ONCE=$v timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
CB200_BIAS_ONCE=$v timeout 300 python scripts/exp/conv_layer_bench.py 64 112 128 3
CB200_BIAS_ONCE=$v timeout 300 python scripts/exp/conv_layer_bench.py 128 112 64 1
done
