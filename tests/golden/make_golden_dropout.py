"""Golden fixture of a training step WITH dropout, produced by the UNMODIFIED reference CPU back-end (oracle/_ref/serial).

Run where /root/reference exists:   python tests/golden/make_golden_dropout.py
tests/netdefs.dropout_net (dropout on conv, pool and dense layers).  The masks are random draws of the reference
(rand()), so the fixture stores the ones it used next to everything that follows from them: with the masks GIVEN, the
step is deterministic and the oracle / the product must reproduce it.  Also stored: the outputs of an inference pass
in AVG_MODEL (no randomness: outputs scaled by 1 - rate) taken before the step.
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


def main():
    spec = netdefs.dropout_net()
    ref = rd.RefNet(spec, "C_BLAS")
    out = {}
    n = ref.n_layers
    for l in range(n):
        if ref.layer_type(l) in (rd.CONV, rd.DENSE):
            out["w0_%d" % l] = ref.weights_view(l).copy()
    x, t = rd.make_inputs(spec, 21)
    out["x"], out["t"] = x, t
    ref.forward(x, None, is_inference=1)
    for l in range(n):
        out["inf_out_%d" % l] = ref.output(l)
    ref.forward(x, None)
    for l in range(n):
        out["out_%d" % l] = ref.output(l)
        m = ref.dropout_mask(l)
        if m is not None:
            out["mask_%d" % l] = m
            print("layer", l, "kept fraction %.3f" % m.mean())
        if ref.layer_type(l) == rd.POOL:
            out["map_%d" % l] = ref.pool_map(l)
    ref.backward(t, HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"])
    for l in range(n):
        out["delta_%d" % l] = ref.delta(l)
        if ref.layer_type(l) in (rd.CONV, rd.DENSE):
            out["w1_%d" % l] = ref.weights_view(l).copy()
            out["m1_%d" % l] = ref.moment_view(l).copy()
    out["length"] = np.array([spec["batch"]])
    out["steps"] = np.array([1])
    path = os.path.join(HERE, "dropout_net_blas.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
