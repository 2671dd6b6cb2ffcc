#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do echo "== default $i"; timeout 300 python scripts/exp/first_net_debug.py 2>&1 | cut -c1-400 | tail -8; done
for i in 1 2 3; do echo "== STATSY0 $i"; CB200_GN_POOL_STATS_Y=0 timeout 300 python scripts/exp/first_net_debug.py 2>&1 | cut -c1-400| tail -5; done
for i in 1 2 3; do echo "== NO_WGRAD_STREAM $i"; CB200_NO_WGRAD_STREAM=1 timeout 300 python scripts/exp/first_net_debug.py 2>&1 | cut -c1-400| tail -5; done
