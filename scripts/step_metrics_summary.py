"""Per-kernel totals of the LAST step in an ncu --csv metrics log (scripts/gpu_step_metrics.sh)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
by = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    e = by.setdefault(int(d["ID"]), {"k": d["Kernel Name"], "grid": d.get("Grid Size")})
    e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
ids = sorted(by)
# the last step starts at the last launch of the network's first kernel (first-layer forward, or the input import)
starts = [i for i in ids if "conv_first_fwd_kernel" in by[i]["k"] or "import_input" in by[i]["k"]]
last = [i for i in ids if i >= starts[-1]] if starts else ids[len(ids) // 2:]
per_launch = len(sys.argv) > 2
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for i in last:
    e = by[i]
    name = e["k"].split("(")[0].replace("void ", "").replace("cb200::", "")[:64]
    t = tot[name]
    t[0] += 1
    t[1] += e["gpu__time_duration.sum"] / 1e3
    t[2] += e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
    t[3] += e["lts__t_bytes.sum"]
    t[4] += e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * e["gpu__time_duration.sum"] / 1e3
    if per_launch and sys.argv[2] in name:
        print("  #%d %-50s %8.1f us dram %6.2f GB l2 %6.2f GB tensor %5.1f%% grid %s" % (i, name, e["gpu__time_duration.sum"] / 1e3,
              (e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]) / 1e9, e["lts__t_bytes.sum"] / 1e9,
              e["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], e["grid"]))
tt = sum(v[1] for v in tot.values())
print("launches in step %d, total %.1f us" % (len(last), tt))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-64s n=%3d %8.1f us %5.1f%% dram %6.2f GB (%5.2f TB/s) l2 %6.2f GB tensor %4.1f%%" % (
        k, v[0], v[1], 100 * v[1] / tt, v[2] / 1e9, v[2] / 1e12 / (v[1] / 1e6) if v[1] else 0, v[3] / 1e9, v[4] / v[1] if v[1] else 0))
