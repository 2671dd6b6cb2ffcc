import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f img/s  e2e %.1f  ms/step %.2f  infer %.1f  launches %d  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["inference"]["value"], d["gpu_launches"], d["clocks"]))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items()})
for k, v in d["kernel_families"].items():
    print("  %-22s %7.3f ms/step  %s" % (k, v["ms_per_step"], {a: round(b, 1) for a, b in v.items() if a != "ms_per_step"}))
