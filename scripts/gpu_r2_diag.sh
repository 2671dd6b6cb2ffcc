#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp/gn_apply_sweep.py 2>&1 | tail -16
