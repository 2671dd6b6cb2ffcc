#!/bin/bash
# round-2, final build: compute-sanitizer memcheck over the network-level GPU tests (host library + every kernel in sequence),
# then synccheck over the operator tests (barrier / mbarrier misuse)
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest \
  tests/test_gpu_network.py tests/test_gpu_regression.py tests/test_gpu_perf_eval.py -m gpu -q -x --tb=short --durations=5 \
  > gpurun_out/r2_compute_sanitizer_memcheck_net.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_compute_sanitizer_memcheck_net.log | tail -6
timeout 240 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest \
  tests/test_gpu_ops.py tests/test_gpu_gn_epilogue.py tests/test_gpu_strided.py -m gpu -q -x --tb=short -k "not network_uses" \
  > gpurun_out/r2_compute_sanitizer_synccheck.log 2>&1
echo "synccheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Barrier|Error" gpurun_out/r2_compute_sanitizer_synccheck.log | tail -6
