#!/bin/bash
# full ncu captures of the early-layer kernels that sit 2-3x above their HBM bound (guidance for the next round)
mkdir -p gpurun_out
bash scripts/gpu_ncu_full.sh conv_halo_kernel 0 6 128 halo_r1b
ncu -i gpurun_out/halo_r1b.ncu-rep --page raw --csv > gpurun_out/halo_full_raw.csv 2>/dev/null
bash scripts/gpu_ncu_full.sh conv_first 0 2 128 first_r1b
ncu -i gpurun_out/first_r1b.ncu-rep --page raw --csv > gpurun_out/first_full_raw.csv 2>/dev/null
wc -c gpurun_out/halo_full_raw.csv gpurun_out/first_full_raw.csv
