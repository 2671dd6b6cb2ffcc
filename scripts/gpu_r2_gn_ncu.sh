#!/bin/bash
# round-2: ncu --set full of the group-norm launches of one Darknet19 training step (batch 128)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"norm_" -c 80 -f -o gpurun_out/gn_full_b128 \
    python scripts/profile_step.py --batch 128 --warmup 0 --steps 1 > gpurun_out/ev_gn.log 2>&1
tail -2 gpurun_out/ev_gn.log
ncu -i gpurun_out/gn_full_b128.ncu-rep --page raw --csv > gpurun_out/gn_full_b128_raw.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/gn_full_b128_raw.csv > gpurun_out/gn_full_b128_digest.txt; wc -l gpurun_out/gn_full_b128_digest.txt
rm -f gpurun_out/gn_full_b128.ncu-rep
timeout 900 python scripts/exp/gn_apply_sweep.py > gpurun_out/gn_apply_sweep.txt 2>&1; cat gpurun_out/gn_apply_sweep.txt
