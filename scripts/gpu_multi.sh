#!/bin/bash
# N-GPU checks: DP correctness + scaling bench
N=${1:-2}; B=${2:-128}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_dp.py -q --tb=short > gpurun_out/r2_dp_tests_world2.log 2>&1; tail -5 gpurun_out/r2_dp_tests_world2.log
for n in 1 $N; do
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 --steps 8 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/r2_scale2_$n.json 2> gpurun_out/r2_scale2_$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 8 --warmup 3 --batch $B > gpurun_out/r2_scale2_$n.json 2> gpurun_out/r2_scale2_$n.err; fi
  tail -c 400 gpurun_out/r2_scale2_$n.err; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_scale2_$n.json") if l.startswith("{")][-1]); print("N=$n value %.1f e2e %.1f ms/step %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]))
except Exception as e: print("N=$n failed", e)
PY
done
