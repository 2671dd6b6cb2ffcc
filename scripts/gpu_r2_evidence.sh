#!/bin/bash
# round-2 evidence on the FINAL build (batch 128): launch list of the bench command, per-launch metrics of one step,
# one --set full capture of the dominant kernel family, bench line.  Run last: profiles/r2_conv_metrics_summary.json is
# tied to the kernel sources by bench.kernel_source_id().
mkdir -p gpurun_out
# (1) launch list of the bench command itself (B200_PROFILING.md recipe)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev0.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_bench.csv 2>/dev/null | head -14
# (2) every launch of one training step: time, DRAM bytes, L2 bytes, tensor-pipe activity
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none --csv --log-file gpurun_out/r2_step_metrics_b128.csv \
    python scripts/profile_step.py --batch 128 --warmup 1 --steps 1 > gpurun_out/ev1.log 2>&1
python scripts/step_metrics_summary.py gpurun_out/r2_step_metrics_b128.csv > gpurun_out/r2_step_metrics_b128_summary.txt; head -30 gpurun_out/r2_step_metrics_b128_summary.txt
python scripts/make_conv_metrics_summary.py gpurun_out/r2_step_metrics_b128.csv 128 gpurun_out/r2_conv_metrics_summary.json | tail -25
# (3) --set full of the dominant family: two CTA-pair launches inside a step (a deep 3x3 layer and a 1x1 layer)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_pair_kernel" -s 4 -c 2 -f -o gpurun_out/r2_igemm_pair_full \
    python scripts/profile_step.py --batch 128 --warmup 0 --steps 1 > gpurun_out/ev3.log 2>&1
ncu -i gpurun_out/r2_igemm_pair_full.ncu-rep --page raw --csv > gpurun_out/r2_igemm_pair_full_b128_raw.csv 2>/dev/null
python scripts/ncu_raw_digest.py gpurun_out/r2_igemm_pair_full_b128_raw.csv
rm -f gpurun_out/r2_igemm_pair_full.ncu-rep
# (4) the bench line, with the fresh summary in place
cp gpurun_out/r2_conv_metrics_summary.json profiles/r2_conv_metrics_summary.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_b128_final.json 2> gpurun_out/r2_bench_b128_final.err; tail -2 gpurun_out/r2_bench_b128_final.err
timeout 900 python bench.py --steps 10 --warmup 3 --precision BF16C_FP32A --no-cpu-baseline > gpurun_out/r2_bench_b128_bf16.json 2>/dev/null
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; cat gpurun_out/r2_bench_reference_arm.json | cut -c1-600
python - <<'PY'
import json
for f in ('r2_bench_b128_final', 'r2_bench_b128_bf16'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks'], 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], {k:round(v['ms_per_step'],3) for k,v in d['kernel_families'].items()}, 'inf', round(d['inference']['value']), d.get('cpu_baseline'))
PY
