#!/bin/bash
# 2 GPUs: does the number of NCCL channels (CTAs that share SMs with the persistent conv kernels) matter?
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/nccl_$name.json 2> gpurun_out/nccl_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/nccl_$name.json") if l.startswith("{")][-1]); print("$name value %.1f ms/step %.3f per-rank %s clocks %s" % (d["value"], d["ms_per_step"], d.get("per_rank_ms_per_step"), [c.get("sm_mhz") for c in d["clocks"].get("per_rank", [])]))
except Exception as e: print("$name failed", e)
PY
  grep -m2 -E "channels|nChannels|Channel 0" gpurun_out/nccl_$name.err | cut -c1-160
}
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/nccl_n1.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/nccl_n1.json') if l.startswith('{')][-1]); print('N=1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
run default NCCL_DEBUG=INFO
run ch2 NCCL_MAX_NCHANNELS=2
run ch4 NCCL_MAX_NCHANNELS=4
run ch8 NCCL_MAX_NCHANNELS=8
run ch4_ll NCCL_MAX_NCHANNELS=4 NCCL_PROTO=Simple
