#!/bin/bash
# round-2: compute-sanitizer memcheck over the tests of the kernels added this round, then a bench run (clock sampler)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_update_plan.py tests/test_gpu_gn_epilogue.py tests/test_gpu_shuffle.py tests/test_gpu_generic_geometry.py "tests/test_gpu_ops.py::test_group_norm_max_pool_fused_equals_unfused" "tests/test_gpu_ops.py::test_group_norm_vs_oracle" -m gpu -q -x --tb=short -k "not network_uses" > gpurun_out/r2_compute_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/r2_compute_sanitizer_memcheck.log | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_clk.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_clk.json') if l.startswith('{')][-1]); print(round(d['value']), d['ms_per_step'], d['clocks'], d['roofline']['traffic'], d['roofline']['ncu'])" | cut -c1-700
