#!/bin/bash
# first GPU bring-up: staged so that a hang in the tensor-core path cannot hide the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== A ops (simt/pool/norm)"; timeout 600 python -m pytest tests/test_gpu_ops.py -q --tb=short --maxfail=30 -k "simt_matches or pool or group_norm" > gpurun_out/A_ops_safe.log 2>&1; tail -5 gpurun_out/A_ops_safe.log
echo "== B network forced simt"; CB200_FORCE_SIMT=1 timeout 900 python -m pytest tests/test_gpu_network.py -q --tb=short --maxfail=40 > gpurun_out/B_net_simt.log 2>&1; tail -8 gpurun_out/B_net_simt.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_simt.json 2>/dev/null
echo "== C tcgen05 ops"; timeout 600 python -m pytest tests/test_gpu_ops.py -v --tb=short --maxfail=40 -k "tcgen05" > gpurun_out/C_ops_tc.log 2>&1; tail -30 gpurun_out/C_ops_tc.log
echo "== D network tc"; timeout 900 python -m pytest tests/test_gpu_network.py -q --tb=short --maxfail=40 > gpurun_out/D_net_tc.log 2>&1; tail -8 gpurun_out/D_net_tc.log
echo "== E smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/E_smoke.log 2>&1; tail -5 gpurun_out/E_smoke.log
echo "== F bench"; timeout 900 python bench.py --steps 3 --warmup 3 --batch 32 --no-cpu-baseline > gpurun_out/F_bench.log 2>&1; tail -3 gpurun_out/F_bench.log
