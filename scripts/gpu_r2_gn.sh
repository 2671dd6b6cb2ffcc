#!/bin/bash
# round-2: group-norm statistics from the conv epilogue: parity, then bench A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gn_epilogue.py tests/test_gpu_update_plan.py -m gpu -q --tb=short > gpurun_out/tests_gn.log 2>&1; tail -30 gpurun_out/tests_gn.log | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_backends.py tests/test_gpu_network.py tests/test_gpu_darknet19_full.py tests/test_gpu_configs.py -m gpu -q --tb=short > gpurun_out/tests_gn2.log 2>&1; tail -12 gpurun_out/tests_gn2.log | cut -c1-300
CB200_GN_EPILOGUE_STATS=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gn_off.json 2> gpurun_out/bench_gn_off.err; tail -3 gpurun_out/bench_gn_off.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gn_on.json 2> gpurun_out/bench_gn_on.err; tail -3 gpurun_out/bench_gn_on.err
python - <<'PY'
import json
for f in ('off','on'):
    d=json.loads(open('gpurun_out/bench_gn_%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d['gpu_launches'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()}, 'inf', round(d['inference']['value']), {k:round(v['ms_per_step'],3) for k,v in d['inference']['kernel_families'].items()})
PY
