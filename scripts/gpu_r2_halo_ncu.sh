#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp/conv_layer_bench.py 32 224 64 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -s 2 -c 1 -f -o gpurun_out/halo_f python scripts/exp/conv_layer_bench.py 32 224 64 3 > gpurun_out/ev_h.log 2>&1
ncu -i gpurun_out/halo_f.ncu-rep --page raw --csv > gpurun_out/halo_f_raw.csv 2>/dev/null
ncu -i gpurun_out/halo_f.ncu-rep --page source --csv --print-source sass > gpurun_out/halo_f_sass.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/halo_f_raw.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel" -s 8 -c 1 -f -o gpurun_out/halo_d python scripts/exp/conv_layer_bench.py 32 224 64 3 > gpurun_out/ev_h.log 2>&1
ncu -i gpurun_out/halo_d.ncu-rep --page raw --csv > gpurun_out/halo_d_raw.csv 2>/dev/null
ncu -i gpurun_out/halo_d.ncu-rep --page source --csv --print-source sass > gpurun_out/halo_d_sass.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/halo_d_raw.csv
rm -f gpurun_out/halo_f.ncu-rep gpurun_out/halo_d.ncu-rep
