#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient: nothing charged)
# usage: scripts/gpurun_retry.sh <timeout> <command string> [gpus]
T=$1; CMD=$2; G=${3:-1}
for i in $(seq 1 30); do
  if [ "$G" = "1" ]; then OUT=$(gpurun --timeout $T -- "$CMD" 2>&1); else OUT=$(gpurun --gpus $G --timeout $T -- "$CMD" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient\|no box\|busy"; then sleep 90; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; exit 3
