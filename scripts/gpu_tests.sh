#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=40 > gpurun_out/tests_gpu.log 2>&1; tail -25 gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
