"""Golden fixtures of the YOLO output head, produced by the UNMODIFIED reference CPU back-end
(oracle/_ref/serial: src/activ_functions.c compiled with float abs(), i.e. the arithmetic of the reference's CUDA
kernels - SURVEY.md 8c).

Run where /root/reference exists:   python tests/golden/make_golden_yolo.py
Per case of tests/netdefs.YOLO_HEAD_CASES: raw head values x, target rows t, then what the reference makes of them:
activated output a, error signal delta (TC_scale_factor 1), box_locked states, per-element loss, IoU monitor.
A second set (suffix _h) repeats delta / state / loss / monitor for a and t rounded to FP16 first, which is what the
mixed-precision product reads.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402
from tests import netdefs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def rz16(a):
    """FP32 -> FP16 round toward zero -> FP32 (the dataset cast of upstream, src/cuda/cuda_main.cu:790-813)"""
    a = np.asarray(a, dtype=np.float32)
    h = a.astype(np.float16)
    over = np.abs(h.astype(np.float32)) > np.abs(a)
    h = np.where(over, np.nextafter(h, np.float16(0)), h)
    return h.astype(np.float32)


def run_case(ref, a, t, nb_box):
    ref.set_last_output(a)
    ref.last_deriv_error(t)
    last = ref.n_layers - 1
    delta, state = ref.delta(last), ref.yolo_box_state(nb_box)
    ref.set_last_output(a)
    loss = ref.loss(t)
    return delta, state, loss, ref.yolo_monitor(nb_box)


def capture(name, seed):
    spec = netdefs.yolo_head(name)
    ref = rd.RefNet(spec, "C_BLAS")
    ref.set_iter(1, spec["batch"])
    rng = np.random.default_rng(seed)
    last = ref.n_layers - 1
    shape = ref.out_shape(last)
    x = (1.5 * rng.standard_normal(shape)).astype(np.float32)
    n_obj = None
    t = rd.make_yolo_targets(spec, seed + 1, n_obj)
    if name == "giou_default":
        t[0, 0] = -1.0       # "class only" image: one target, no geometry fit
    ref.set_last_output(x)
    ref.last_activation()
    a = ref.output(last)
    nb_box = spec["yolo"]["nb_box"]
    out = dict(x=x, t=t, a=a)
    out["delta"], out["state"], out["loss"], out["monitor"] = run_case(ref, a, t, nb_box)
    ah, th = a.astype(np.float16).astype(np.float32), rz16(t)
    out["delta_h"], out["state_h"], out["loss_h"], out["monitor_h"] = run_case(ref, ah, th, nb_box)
    path = os.path.join(HERE, "yolo_%s.npz" % name)
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "assoc", int((out["state"] == 2).sum()),
          "good-not-best", int((out["state"] == 1).sum()), "targets", t[:, 0])


if __name__ == "__main__":
    for i, name in enumerate(netdefs.YOLO_HEAD_CASES):
        capture(name, 300 + 10 * i)
