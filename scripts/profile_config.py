"""Runs W + K training steps of one of the other BASELINE configurations (cianna_b200/configs.py) on a device-resident
synthetic batch, for `ncu` launch lists:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
      --clock-control none --csv --log-file gpurun_out/cfg.csv python scripts/profile_config.py --config extinction --batch 128
  python scripts/step_metrics_summary.py gpurun_out/cfg.csv
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cianna_b200 import CIANNA as cnn, configs, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="extinction", choices=["extinction", "sdc1", "coco", "mnist"])
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--precision", default="FP16C_FP32A")
a = ap.parse_args()
spec = {"extinction": lambda: configs.extinction_profile(a.batch), "sdc1": lambda: configs.sdc1_yolo(a.batch, 512),
        "coco": lambda: configs.darknet19_yolo(a.batch, 416), "mnist": lambda: configs.lenet(a.batch, dropout=True)}[a.config]()
with utils.Quiet():
    utils.build_network(cnn, spec, "C_CUDA", a.precision, network=0)
cnn.set_TC_scale_factor(16.0, network=0)
rng = np.random.default_rng(0)
dim = spec["in_dim"][0] * spec["in_dim"][1] * spec["in_ch"]
x = np.zeros((a.batch, dim + 1), np.float32)
x[:, :dim] = rng.standard_normal((a.batch, dim)).astype(np.float32)
t = np.zeros((a.batch, spec["out_dim"]), np.float32)
if "yolo" in spec:
    y = spec["yolo"]
    per = 7 + y.get("nb_param", 0) + y.get("diff_flag", 0)
    n_obj = min(40, y["max_nb_obj_per_image"])
    for b in range(a.batch):
        t[b, 0] = n_obj
        for j in range(n_obj):
            cx, cy = rng.uniform(20, spec["in_dim"][0] - 20, 2)
            w, h = rng.uniform(6, 30, 2)
            row = t[b, 1 + j * per: 1 + (j + 1) * per]
            row[0] = rng.integers(1, max(1, y.get("nb_class", 0)) + 1)
            row[1:7] = (cx - w / 2, cy - h / 2, 0.0, cx + w / 2, cy + h / 2, 1.0)
    cnn.set_iter(10, 100000, network=0)
else:
    t[:] = rng.random(t.shape, dtype=np.float32)
cnn.load_batch(x, t, network=0)
for _ in range(a.warmup + a.steps):
    cnn.forward_batch(network=0)
    cnn.backward_batch(1e-5, 0.9, network=0)
print("loss", cnn.batch_loss(network=0))
