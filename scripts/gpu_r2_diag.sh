#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_strided.py tests/test_gpu_wholemap.py tests/test_gpu_network.py -m gpu -q --tb=line -x > gpurun_out/tests_ops.log 2>&1; tail -3 gpurun_out/tests_ops.log | cut -c1-300
timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
timeout 300 python scripts/exp/conv_layer_bench.py 64 112 128 3
timeout 300 python scripts/exp/conv_layer_bench.py 256 28 512 3
timeout 300 python scripts/exp/first_layer_bench.py | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prod.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_prod.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()}, round(d['inference']['value']))"
