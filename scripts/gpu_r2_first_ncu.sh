#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_first_fwd" -s 1 -c 1 -f -o gpurun_out/first_full \
    python scripts/exp/first_layer_bench.py 128 > gpurun_out/ev_first.log 2>&1
ncu -i gpurun_out/first_full.ncu-rep --page raw --csv > gpurun_out/first_full_raw.csv 2>/dev/null
ncu -i gpurun_out/first_full.ncu-rep --page source --csv --print-source sass > gpurun_out/first_full_source_sass.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/first_full_raw.csv
rm -f gpurun_out/first_full.ncu-rep
