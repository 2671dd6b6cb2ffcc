"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference CPU back-end
(oracle/_ref, built by oracle/build_ref.sh from /root/reference/src) on seeded inputs.

Run where /root/reference exists:   python tests/golden/make_golden.py
Each fixture holds: the seeded input/target batch, the initial weights the reference drew, and, after ONE
training step (forward, output error, backprop + SGD update), every layer's output, delta, pool argmax map,
group-norm statistics and updated weights / momentum buffers, all in the reference's own layouts.
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


def randomize_norms(ref, rng):
    for l in range(ref.n_layers):
        if ref.layer_type(l) == rd.NORM:
            g = ref.norm_view(l, "gamma")
            b = ref.norm_view(l, "beta")
            g[:] = 1.0 + 0.2 * rng.standard_normal(g.shape).astype(np.float32)
            b[:] = 0.1 * rng.standard_normal(b.shape).astype(np.float32)


def capture(name, spec, comp_meth, seed, length=None, steps=1):
    ref = rd.RefNet(spec, comp_meth)
    rng = np.random.default_rng(seed + 1000)
    randomize_norms(ref, rng)
    out = {}
    for l in range(ref.n_layers):
        t = ref.layer_type(l)
        if t in (rd.CONV, rd.DENSE):
            out["w0_%d" % l] = ref.weights_view(l).copy()
        if t == rd.NORM:
            out["w0_%d" % l] = np.concatenate([ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")]).copy()
    for s in range(steps):
        x, tgt = rd.make_inputs(spec, seed + s)
        ref.forward(x, length)
        if s == steps - 1:
            out["x"], out["t"] = x, tgt
            for l in range(ref.n_layers):
                out["out_%d" % l] = ref.output(l)
                if ref.layer_type(l) == rd.POOL and ref.geom(l)[1] == 0:
                    out["map_%d" % l] = ref.pool_map(l)
                if ref.layer_type(l) == rd.NORM:
                    out["mean_%d" % l] = ref.norm_view(l, "mean").copy()
                    out["var_%d" % l] = ref.norm_view(l, "var").copy()
            last = ref.n_layers - 1
            if ref.layer_type(last) != rd.DENSE:
                out["loss"] = ref.loss(tgt)
        else:
            out["x_prev%d" % s], out["t_prev%d" % s] = x, tgt
        ref.backward(tgt, HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"])
    for l in range(ref.n_layers):
        t = ref.layer_type(l)
        out["delta_%d" % l] = ref.delta(l)
        if t in (rd.CONV, rd.DENSE):
            out["w1_%d" % l] = ref.weights_view(l).copy()
            out["m1_%d" % l] = ref.moment_view(l).copy()
        if t == rd.NORM:
            out["w1_%d" % l] = np.concatenate([ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")]).copy()
            out["dgamma_%d" % l] = ref.norm_view(l, "d_gamma").copy()
            out["dbeta_%d" % l] = ref.norm_view(l, "d_beta").copy()
    out["length"] = np.array([spec["batch"] if length is None else length])
    out["steps"] = np.array([steps])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    capture("mini_darknet_blas", netdefs.mini_darknet(), "C_BLAS", seed=11)
    capture("mini_darknet_naiv_tail", netdefs.mini_darknet(), "C_NAIV", seed=12, length=3)
    capture("mini_darknet_2steps", netdefs.mini_darknet(), "C_BLAS", seed=13, steps=2)
    capture("tc_darknet_blas", netdefs.tc_darknet(batch=4, size=8), "C_BLAS", seed=14)
    capture("lenet_small_blas", netdefs.lenet(batch=8, size=16, d1=64, d2=32), "C_BLAS", seed=15)
