"""CPU checks of oracle/lrn_oracle.py: the fixture upstream's OWN CUDA kernels produced on a B200
(tests/golden/lrn_refcuda.npz, written by tests/test_gpu_backends.py::test_lrn_of_upstream_cuda_path_pins_the_lrn_oracle
from oracle/_ref/cuda, the unmodified src/cuda/cuda_lrn_layer.cu compiled for sm_100), closed-form values on a tiny
case and the backward formula against central finite differences of the forward one."""
import os

import numpy as np

from oracle import cianna_oracle as co
from oracle import lrn_oracle as lo
from tests.common import GOLDEN_DIR, rel_err


def test_lrn_oracle_matches_upstream_cuda_fixture():
    """both LRN layers of tests/netdefs.lrn_net (explicit parameters / upstream's defaults), forward and backward
    (upstream's backward ends with the leaky-ReLU derivative of the conv layer below, cuda_lrn_layer.cu:207-214)"""
    g = np.load(os.path.join(GOLDEN_DIR, "lrn_refcuda.npz"))
    layers = sorted(int(k.split("_")[1]) for k in g.files if k.startswith("param_"))
    assert layers == [1, 4]
    for l in layers:
        r, k, alpha, beta = g["param_%d" % l]
        x, y_ref, dy, dx_ref = g["x_%d" % l], g["y_%d" % l], g["dy_%d" % l], g["dx_%d" % l]
        y, scale = lo.lrn_forward(x, int(r), k, alpha, beta)
        assert rel_err(y, y_ref) < 1e-5
        dx = co.relu_deriv(lo.lrn_backward(x, y_ref, dy, scale, int(r), alpha, beta), x, x.shape[1])
        assert np.abs(dx_ref).max() > 0
        assert rel_err(dx, dx_ref) < 1e-4


def test_lrn_forward_known_values():
    # 3 channels, one pixel, range 3 (one neighbour each side), k=1, alpha=3, beta=1 -> s_c = 1 + sum(window x^2)
    x = np.array([1.0, 2.0, 3.0], np.float32).reshape(3, 1, 1)
    y, s = lo.lrn_forward(x, 3, 1.0, 3.0, 1.0)
    assert np.allclose(s.ravel(), [1 + 5, 1 + 14, 1 + 13])
    assert np.allclose(y.ravel(), [1 / 6, 2 / 15, 3 / 14])
    # even range: 4 // 2 = 2 channels on both sides, still divided by 4
    y, s = lo.lrn_forward(x, 4, 2.0, 4.0, 0.5)
    assert np.allclose(s.ravel(), [2 + 14, 2 + 14, 2 + 14])
    assert np.allclose(y.ravel(), np.array([1, 2, 3]) / 4.0)


def test_lrn_backward_is_the_gradient_of_forward():
    rng = np.random.default_rng(0)
    C, B, A, r, k, alpha, beta = 7, 2, 3, 5, 1.5, 0.8, 0.75
    x = rng.standard_normal((C, B, A))
    dy = rng.standard_normal((C, B, A))

    def f64(xv):      # the forward pass in float64, same window rule
        out = np.empty_like(xv)
        for c in range(C):
            lo_, hi = lo._window(c, C, r)
            out[c] = xv[c] / (k + alpha * (xv[lo_:hi + 1] ** 2).sum(axis=0) / r) ** beta
        return out

    y, s = lo.lrn_forward(x.astype(np.float32), r, k, alpha, beta)
    dx = lo.lrn_backward(x, y, dy, s, r, alpha, beta)
    num = np.zeros_like(x)
    eps = 1e-5
    for idx in np.ndindex(*x.shape):
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        num[idx] = ((f64(xp) - f64(xm)) * dy).sum() / (2 * eps)
    assert np.abs(dx - num).max() < 2e-5 * max(1.0, np.abs(num).max())
