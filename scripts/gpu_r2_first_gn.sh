#!/bin/bash
# round-2: group-norm sums in the first layer's epilogue + packed bias / activation in the conv epilogues - tests, per-layer timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_gn_epilogue.py tests/test_gpu_ops.py tests/test_gpu_strided.py tests/test_gpu_wholemap.py -m gpu -q -x --tb=short 2>&1 | tail -12
timeout 120 python scripts/exp/first_layer_bench.py 2>&1 | tail -2
timeout 200 python scripts/exp/gn_epilogue_bench.py 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value'],1), d['ms_per_step'], d['clocks']['sm_mhz'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()})"
