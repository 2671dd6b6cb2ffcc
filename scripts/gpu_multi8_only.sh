#!/bin/bash
# 8-GPU data-parallel bench only (one torchrun launch)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 8 --warmup 3 --batch 128 > gpurun_out/scale_8.json 2> gpurun_out/scale_8.err
tail -c 300 gpurun_out/scale_8.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_8.json") if l.startswith("{")][-1]); print("N=8 value %.1f e2e %.1f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e: print("N=8 failed", e)
PY
