#!/bin/bash
# 8-GPU data-parallel check: DP tests (2 ranks) + bench at N = 8 (and 4)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_dp.py -q --tb=short > gpurun_out/dp_tests.log 2>&1; tail -3 gpurun_out/dp_tests.log
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 6 --warmup 3 --batch 128 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  tail -c 300 gpurun_out/scale_$n.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$n.json") if l.startswith("{")][-1]); print("N=$n value %.1f e2e %.1f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e: print("N=$n failed", e)
PY
done
