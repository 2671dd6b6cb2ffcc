#!/bin/bash
# round-2, 8 GPUs of one box: data-parallel correctness at world 8 (scripts/dp_check.py), then the bench at N = 1 and 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for mode in FP16C_FP32A; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 scripts/dp_check.py $mode > gpurun_out/r2_dp_check_world8_$mode.log 2>&1
  grep -E "DP_CHECK|max rel|err" gpurun_out/r2_dp_check_world8_$mode.log | tail -12
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_1.json 2> gpurun_out/r2_scale_1.err
NCCL_DEBUG=INFO timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_scale_8.json 2> gpurun_out/r2_scale_8.err
grep -m3 -E "NVLS|Ring|Tree|via P2P" gpurun_out/r2_scale_8.err | cut -c1-200
python - <<'PY'
import json
r = {}
for n in (1, 8):
    try:
        d=json.loads([l for l in open("gpurun_out/r2_scale_%d.json" % n) if l.startswith("{")][-1]); r[n] = d
        print("N=%d value %.1f e2e %.1f ms/step %.3f per-rank %s clocks %s" % (n, d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("per_rank_ms_per_step"), d.get("clocks")))
    except Exception as e: print("N=%d failed" % n, e)
if 1 in r and 8 in r: print("weak-scaling efficiency %.4f" % (r[8]["value"] / (8 * r[1]["value"])))
PY
