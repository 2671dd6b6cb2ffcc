#!/bin/bash
# round-2, final build: compute-sanitizer memcheck over the operator-level GPU tests (every kernel family at small shapes)
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest \
  tests/test_gpu_ops.py tests/test_gpu_strided.py tests/test_gpu_wholemap.py tests/test_gpu_yolo.py tests/test_gpu_lrn.py \
  tests/test_gpu_dropout.py tests/test_gpu_update_plan.py tests/test_gpu_gn_epilogue.py tests/test_gpu_shuffle.py \
  tests/test_gpu_generic_geometry.py -m gpu -q -x --tb=short -k "not network_uses" --durations=8 \
  > gpurun_out/r2_compute_sanitizer_memcheck_all.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_compute_sanitizer_memcheck_all.log | tail -8
tail -15 gpurun_out/r2_compute_sanitizer_memcheck_all.log | cut -c1-200
