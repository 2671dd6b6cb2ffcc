#!/bin/bash
# CTA-pair (cta_group::2) hardware probe, then the bit-exact tests of conv_igemm_pair_kernel (bounded: a hang costs 120 s)
mkdir -p gpurun_out
(cd scripts/exp && nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o cta_pair_probe cta_pair_probe.cu -lcuda 2>&1 | tail -3; timeout 60 ./cta_pair_probe) > gpurun_out/cta_pair_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/cta_pair_probe.log; cat gpurun_out/cta_pair_probe.log
if grep -q "mode 0 (K-major) K=256: exact" gpurun_out/cta_pair_probe.log; then
  timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -x -k "address_mapping" > gpurun_out/tests_pair.log 2>&1; tail -15 gpurun_out/tests_pair.log
fi
