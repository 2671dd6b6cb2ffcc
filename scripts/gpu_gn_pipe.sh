#!/bin/bash
# pipelined group-norm: parity tests, then the chunk-size / occupancy sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -k "group_norm" > gpurun_out/tests_gn.log 2>&1; tail -15 gpurun_out/tests_gn.log
timeout 900 python scripts/exp/gn_pipeline_sweep.py ${1:-128} > gpurun_out/gn_sweep.log 2>&1; tail -80 gpurun_out/gn_sweep.log | cut -c1-160
