#!/bin/bash
# halo tests, then per-launch times of the early-layer forward kernels with / without the TMA-store epilogues (ncu, same box)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short --maxfail=10 -k "halo or saturation or first" > gpurun_out/tests_halo.log 2>&1; tail -4 gpurun_out/tests_halo.log
for v in on off; do
  if [ $v = off ]; then export CB200_NO_TMA_STORE=1; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed -k regex:"conv_halo_kernel|conv_first_fwd" --clock-control none --csv --log-file gpurun_out/early_$v.csv python scripts/profile_step.py --batch 128 --warmup 1 --steps 1 > gpurun_out/early_$v.log 2>&1
  echo "== TMA store $v"; grep -v "^==" gpurun_out/early_$v.csv | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value'); i=h.index('ID')
d={}
for row in r[1:]:
    d.setdefault(row[i],{'k':row[k][:52]})[row[m]]=row[v]
for e in list(d.values())[-7:]: print(e['k'], e.get('gpu__time_duration.sum'), e.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'))
"
done
