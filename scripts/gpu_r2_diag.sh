#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_update_plan.py tests/test_gpu_dropout.py tests/test_gpu_regression.py tests/test_gpu_darknet19_full.py -m gpu -q --tb=short -x > gpurun_out/tests_early.log 2>&1; tail -4 gpurun_out/tests_early.log | cut -c1-300
for v in 0 1 0 1; do CB200_EARLY_UPDATES=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_early_$v.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_early_$v.json') if l.startswith('{')][-1]); print('early=$v', round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'])"; done
