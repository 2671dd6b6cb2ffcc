"""Helpers shared by the test modules."""
import os

import numpy as np

from oracle import ref_loader
from oracle.oracle_net import OracleNet
from tests import netdefs

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HYPER = dict(lr=0.02, momentum=0.9, weight_decay=0.0005)   # must match tests/golden/make_golden.py

GOLDEN_SPECS = {
    "mini_darknet_blas": lambda: netdefs.mini_darknet(),
    "mini_darknet_naiv_tail": lambda: netdefs.mini_darknet(),
    "mini_darknet_2steps": lambda: netdefs.mini_darknet(),
    "tc_darknet_blas": lambda: netdefs.tc_darknet(batch=4, size=8),
    "lenet_small_blas": lambda: netdefs.lenet(batch=8, size=16, d1=64, d2=32),
}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def rel_err(a, b):
    """max |a-b| normalised by the largest reference magnitude (robust for tensors with many zeros)"""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.size == b.size, (a.size, b.size)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


def rel_l2(a, b):
    """relative Frobenius error ||a-b|| / ||b||.  Used for mixed-precision deltas / gradients: there a single
    element whose pre-activation changes sign under 16-bit rounding flips its leaky-ReLU slope (x1 <-> x0.05) or moves a
    max-pool argmax, which makes the POINTWISE error of a delta tensor ill-conditioned while the tensor as a whole is
    accurate to the storage precision."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.size == b.size, (a.size, b.size)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def oracle_from_golden(spec, g):
    net = OracleNet(spec)
    for L in net.layers:
        key = "w0_%d" % L["idx"]
        if L["kind"] in ("conv", "dense"):
            L["weights"] = g[key].copy()
        elif L["kind"] == "norm":
            G = L["gamma"].size
            L["gamma"], L["beta"] = g[key][:G].copy(), g[key][G:].copy()
    return net


def ref_available():
    return ref_loader.available("serial")


def rel_q(a, b, q=0.98):
    """q-quantile of |a-b| normalised by max|ref|.  Mixed-precision backward tensors are compared with this at the
    2e-2 tolerance: 16-bit rounding of the forward activations flips a small fraction (measured ~0.1-0.5 %) of DISCRETE
    decisions with respect to the FP32 reference - a max-pool argmax between two nearly equal candidates, the leaky-ReLU
    slope of a value next to zero - and each flip is a full-size error on one element (rel. L2 = sqrt(2 f) ~ 5-10 % for
    f = 0.3 %), the same on the reference's own FP16 CUDA path.  The quantile checks that everything else is within
    tolerance; grad_cos() bounds the damage of the flips on the summed gradients; the identical-input operator tests
    (tests/test_gpu_ops.py) hold each backward kernel itself to 2e-2 in max-norm."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.size == b.size, (a.size, b.size)
    return float(np.quantile(np.abs(a - b), q) / max(np.abs(b).max(), 1e-30))


def grad_cos(a, b):
    """1 - cosine similarity"""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(1.0 - a.dot(b) / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
