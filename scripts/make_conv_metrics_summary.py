"""profiles/r2_conv_metrics_summary.json from an ncu per-launch metrics log of one step (scripts/gpu_step_metrics.sh):
mean DRAM bytes per launch and time-weighted tensor-pipe activity of the forward/data-gradient kernels and of the
weight-gradient kernels. bench.py reports these next to its live roofline numbers."""
import collections
import csv
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_id  # noqa: E402

src, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
by = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = by.setdefault(int(d["ID"]), {"k": d["Kernel Name"]})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(by)
# the last step starts at the last launch of the network's first kernel (first-layer forward, or the input import)
starts = [i for i in ids if "conv_first_fwd_kernel" in by[i]["k"] or "import_input" in by[i]["k"]]
last = [by[i] for i in ids if i >= starts[-1]] if starts else [by[i] for i in ids[len(ids) // 2:]]
groups = {"igemm": ("conv_igemm_kernel", "conv_igemm_pair_kernel", "conv_halo_kernel", "conv_first_fwd_kernel"), "wgrad": ("conv_wgrad_kernel", "conv_wgrad_pair_kernel", "conv_wgrad_swap_kernel", "conv_first_wgrad_kernel")}
out = {"batch": batch, "kernel_source_id": kernel_source_id(), "note": "ncu per-launch metrics of one Darknet19-448 FP16C_FP32A training step (scripts/gpu_step_metrics.sh -> %s)" % src}
for key, names in groups.items():
    sel = [e for e in last if any(n in e["k"] for n in names)]
    t = sum(e["gpu__time_duration.sum"] for e in sel)
    out[key] = {"kernels": list(names), "launches": len(sel), "total_us": t / 1e3,
                "avg_dram_bytes_per_launch": sum(e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"] for e in sel) / len(sel),
                "tensor_pipe_pct_time_weighted": sum(e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * e["gpu__time_duration.sum"] for e in sel) / t}
json.dump(out, open(sys.argv[3] if len(sys.argv) > 3 else "profiles/r2_conv_metrics_summary.json", "w"), indent=1)
print(json.dumps(out, indent=1))
