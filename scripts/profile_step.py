"""Runs W warm-up + K timed training steps of Darknet19-448 (device-resident batches) - the workload of bench.py's
`value` - with nothing else around it, for `ncu` captures:

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python scripts/profile_step.py --batch 32 --warmup 1 --steps 1
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cianna_b200 import CIANNA as cnn, cabi, utils  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=448)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--precision", default="FP16C_FP32A")
ap.add_argument("--infer", action="store_true")
a = ap.parse_args()
L = cabi.lib()
cabi.check(L.cb200_init(0))
H = cnn.host()
H.cb_train_steps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
H.cb_forward_steps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
with utils.Quiet():
    utils.build_network(cnn, bench.darknet19_spec(a.batch, a.size, 1000), "C_CUDA", a.precision, network=0)
net = cnn._net(0)
cnn.set_TC_scale_factor(256.0, network=0)
x, t = bench.synth_batches(1, a.batch, a.size, 1000, 7)
with utils.Quiet():
    cnn.create_dataset("TRAIN", a.batch, x, t, network=0, silent=1)
    cnn.create_dataset("TEST", a.batch, x, t, network=0, silent=1)
h = bench.HYPER
if a.infer:
    H.cb_forward_steps(net, a.warmup + a.steps, 1, 1)
else:
    H.cb_train_steps(net, a.warmup + a.steps, h["learning_rate"], h["momentum"], h["weight_decay"], 1, 1)
cabi.check(L.cb200_device_sync())
print("done")
