#!/bin/bash
# chunked group-norm: block-count / chunk-size sweep on the large layers, then DRAM bytes per launch under ncu
mkdir -p gpurun_out
GN_SWEEP_SHAPES=0,1,2,3,4 timeout 900 python scripts/exp/gn_pipeline_sweep.py ${1:-128} > gpurun_out/gn_sweep2.log 2>&1; tail -3 gpurun_out/gn_sweep2.log | cut -c1-160
GN_SWEEP_REPS=1 GN_SWEEP_SHAPES=0,2 GN_SWEEP_VARIANTS=0:0,48:2,24:2 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --csv --log-file gpurun_out/gn_chunk_ncu.csv python scripts/exp/gn_pipeline_sweep.py ${1:-128} > gpurun_out/gn_ncu.log 2>&1; tail -3 gpurun_out/gn_ncu.log | cut -c1-160
