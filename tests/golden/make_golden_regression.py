"""Golden fixtures of quadratic-loss networks with LIN / RELU / LOGI output layers and logistic hidden layers, produced by
the UNMODIFIED reference CPU back-end (oracle/_ref/serial).

Run where /root/reference exists:   python tests/golden/make_golden_regression.py
One fixture per case of tests/netdefs.REGRESSION_CASES -> tests/golden/regress_<act>_<head>.npz: initial weights, one
batch (targets in [0, 1]), every layer's output, the per-sample loss, and after one SGD step every layer's delta, the
updated weights and momentum buffers.  A tail batch (length < batch) is used so the zeroing branches are in.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver as rd  # noqa: E402
from tests import netdefs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HYPER = dict(lr=0.02, momentum=0.9, weight_decay=0.0005)
LENGTH = 4


def capture(out_act, head, seed):
    spec = netdefs.regression_net(out_act, head)
    ref = rd.RefNet(spec, "C_BLAS")
    n = ref.n_layers
    out = {}
    for l in range(n):
        if ref.layer_type(l) in (rd.CONV, rd.DENSE):
            out["w0_%d" % l] = ref.weights_view(l).copy()
    rng = np.random.default_rng(seed)
    x, _ = rd.make_inputs(spec, seed)
    t = rng.random((spec["batch"], spec["out_dim"]), dtype=np.float32)
    out["x"], out["t"] = x, t
    ref.forward(x, LENGTH)
    for l in range(n):
        out["out_%d" % l] = ref.output(l)
    out["loss_per_sample"] = ref.loss(t).reshape(-1)      # reduced below, layout differs between dense and conv heads
    shape = ref.out_shape(n - 1)
    e = ref.loss(t)
    out["loss_per_sample"] = (e.sum(axis=1) if ref.layer_type(n - 1) == rd.DENSE else e.sum(axis=(0, 2))).astype(np.float32)
    ref.backward(t, HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"])
    for l in range(n):
        out["delta_%d" % l] = ref.delta(l)
        if ref.layer_type(l) in (rd.CONV, rd.DENSE):
            out["w1_%d" % l] = ref.weights_view(l).copy()
            out["m1_%d" % l] = ref.moment_view(l).copy()
    out["length"] = np.array([LENGTH])
    out["steps"] = np.array([1])
    path = os.path.join(HERE, "regress_%s_%s.npz" % (out_act.lower(), head))
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "shape", shape, "loss", out["loss_per_sample"])


if __name__ == "__main__":
    for k, (act, head) in enumerate(netdefs.REGRESSION_CASES):
        capture(act, head, 40 + k)
