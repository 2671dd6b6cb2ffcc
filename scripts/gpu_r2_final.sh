#!/bin/bash
# round-2 final: whole GPU suite, smoke, then the evidence script (launch list, step metrics, full capture, bench lines)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short --maxfail=30 > gpurun_out/tests_gpu.log 2>&1; tail -3 gpurun_out/tests_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
bash scripts/gpu_r2_evidence.sh 2>&1 | tail -6
