#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_gn_epilogue.py tests/test_gpu_network.py -m gpu -q --tb=short -x > gpurun_out/tests_ops.log 2>&1; tail -6 gpurun_out/tests_ops.log | cut -c1-300
timeout 900 python scripts/exp/gn_apply_sweep.py > gpurun_out/gn_apply_sweep2.txt 2>&1; cat gpurun_out/gn_apply_sweep2.txt
