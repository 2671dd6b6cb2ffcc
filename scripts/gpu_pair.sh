#!/bin/bash
# full GPU suite (the live-reference test three more times: its reference weights are time-seeded), then a bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/tests_gpu.log 2>&1; tail -8 gpurun_out/tests_gpu.log
for i in 1 2 3; do sleep 1; timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q --tb=short -k "live_reference" 2>&1 | tail -2; done
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -c 300 gpurun_out/bench_last.err; python scripts/bench_summary.py gpurun_out/bench_last.json
