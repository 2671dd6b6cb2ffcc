"""Compact per-launch digest of an `ncu --page raw --csv` export: time, the busiest units and the top warp stall reasons.

    python scripts/ncu_raw_digest.py gpurun_out/halo_full_raw.csv
"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.avg.per_second", "sm clk"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("smsp__inst_executed.sum", "warp insts"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "smem")]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("==", d.get("Kernel Name", "")[:70], "grid", d.get("Grid Size", d.get("launch__grid_size", "")))
    print("   " + "  ".join("%s %s%s" % (name, d[k], (" " + u[k]) if u[k] and u[k] != "%" else "") for k, name in KEYS if k in d and d[k]))
    stalls = sorted(((float(v.replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, v in d.items()
                     if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v), reverse=True)
    print("   stalls per issue: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:5]))
