#!/bin/bash
# tests + smoke + bench + launch list in one call:  gpu_all.sh <batch> [regex for per-launch lines]
B=${1:-32}
./scripts/gpu_tests.sh 4 $B
./scripts/gpu_profile.sh $B "$2"
