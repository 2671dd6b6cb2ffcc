#!/bin/bash
# round-2 final (second pass): evidence first (tied to the kernel sources), then the GPU tests not run by
# scripts/gpu_r2_first_gn.sh on this build, then smoke
mkdir -p gpurun_out
bash scripts/gpu_r2_evidence.sh 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=30 --deselect tests/test_gpu_ops.py --deselect tests/test_gpu_strided.py --deselect tests/test_gpu_wholemap.py \
    --ignore tests/test_gpu_ops.py --ignore tests/test_gpu_strided.py --ignore tests/test_gpu_wholemap.py > gpurun_out/tests_gpu_rest.log 2>&1; tail -4 gpurun_out/tests_gpu_rest.log | cut -c1-300
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
