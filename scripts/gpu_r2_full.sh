#!/bin/bash
# round-2: whole GPU suite + smoke + bench line of the current build
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --maxfail=30 > gpurun_out/tests_gpu.log 2>&1; tail -25 gpurun_out/tests_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b128.json 2> gpurun_out/bench_b128.err; tail -3 gpurun_out/bench_b128.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b128.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks'], 'frac', round(d['roofline']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()}, 'inf', round(d['inference']['value']), d['loss'], d['cpu_baseline'])
PY
