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
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


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
