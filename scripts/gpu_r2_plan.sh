#!/bin/bash
# round-2: update-plan bit-exactness, fused norm+pool backward from the pooled output, previously failing tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_update_plan.py tests/test_gpu_ops.py tests/test_gpu_backends.py tests/test_gpu_network.py tests/test_gpu_darknet19_full.py "tests/test_gpu_configs.py::test_darknet19_448_headline_batch_128" -m gpu -q --tb=short > gpurun_out/tests_plan.log 2>&1; tail -30 gpurun_out/tests_plan.log
CB200_GN_POOL_STATS_Y=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y_off.json 2> gpurun_out/bench_y_off.err; tail -3 gpurun_out/bench_y_off.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_y_on.json 2> gpurun_out/bench_y_on.err; tail -3 gpurun_out/bench_y_on.err
python - <<'PY'
import json
for f in ('off','on'):
    d=json.loads(open('gpurun_out/bench_y_%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()})
PY
