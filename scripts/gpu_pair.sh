#!/bin/bash
# the live-reference test repeated (time-seeded reference weights), the network tests, then the default bench line
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do sleep 1; timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q --tb=line -k "live_reference" 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_ops.py -m gpu -q --tb=short 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err; python scripts/bench_summary.py gpurun_out/bench_final.json
