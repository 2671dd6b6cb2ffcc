#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_pair_kernel" -s 2 -c 1 -f -o gpurun_out/pair_f python scripts/exp/conv_layer_bench.py 512 28 256 1 > gpurun_out/ev_p.log 2>&1
ncu -i gpurun_out/pair_f.ncu-rep --page raw --csv > gpurun_out/pair_f_raw.csv 2>/dev/null
ncu -i gpurun_out/pair_f.ncu-rep --page source --csv --print-source sass > gpurun_out/pair_f_sass.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/pair_f_raw.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad_kernel" -s 2 -c 1 -f -o gpurun_out/wg_f python scripts/exp/conv_layer_bench.py 256 28 512 3 > gpurun_out/ev_p.log 2>&1
ncu -i gpurun_out/wg_f.ncu-rep --page raw --csv > gpurun_out/wg_f_raw.csv 2>/dev/null
ncu -i gpurun_out/wg_f.ncu-rep --page source --csv --print-source sass > gpurun_out/wg_f_sass.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/wg_f_raw.csv
rm -f gpurun_out/pair_f.ncu-rep gpurun_out/wg_f.ncu-rep
