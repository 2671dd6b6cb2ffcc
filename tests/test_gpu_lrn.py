"""GPU parity of the local response normalisation layer and of the architecture table export.

LRN: cb200_lrn_forward / _backward through the C-ABI against oracle/lrn_oracle.py (pinned to upstream's own CUDA
kernels by tests/golden/lrn_refcuda.npz, see that file), then the host layer (lrn_create / save / load, python lrn()) inside a network.
Architecture table: the .tex written by print_arch_tex must equal, byte for byte, the files the unmodified reference
wrote for the same networks (tests/golden/arch_tex/, tests/golden/make_golden_arch_tex.py).
"""
import os

import numpy as np
import pytest

from oracle import cianna_oracle as co
from oracle import lrn_oracle as lo
from oracle import ref_driver as rd
from tests import netdefs
from tests.common import GOLDEN_DIR, rel_err

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-5
TOL_MIXED = 2e-2


@pytest.fixture(scope="module")
def cabi():
    from cianna_b200 import cabi as m
    m.init_device(0)
    return m


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    return m


def _representable(x, dtype_name):
    if dtype_name == "FP16":
        return x.astype(np.float16).astype(np.float32)
    if dtype_name == "BF16":
        return (x.view(np.uint32) & 0xFFFF0000).view(np.float32)
    return x


# (batch, channels, size, range, k, alpha, beta): ragged channel counts, even and odd ranges, range > channels
LRN_CASES = [(3, 16, 6, 5, 1.0, 1.0, 0.5), (2, 13, 5, 4, 2.0, 1e-1, 0.75), (2, 40, 7, 7, 1.0, 0.5, 0.75), (2, 3, 4, 9, 1.5, 2.0, 0.5)]


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", LRN_CASES)
def test_lrn_vs_oracle(cabi, cfg, dtype_name):
    B, C, S, rng_, k, alpha, beta = cfg
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    rng = np.random.default_rng(11)
    x = _representable((rng.standard_normal((C, B, S * S)) * 1.3).astype(np.float32), dtype_name)
    dy = _representable(rng.standard_normal((C, B, S * S)).astype(np.float32), dtype_name)
    layer = cabi.LrnLayer(dtype, B, C, S, S, rng_, k, alpha, beta)
    xb, dyb = cabi.upload_act(x, dtype, B, C, S, S), cabi.upload_act(dy, dtype, B, C, S, S)
    y = cabi.download_act(layer.forward(xb), dtype, B, C, S, S)
    ref_y, ref_s = lo.lrn_forward(x, rng_, k, alpha, beta)
    assert rel_err(y, ref_y) < tol
    assert rel_err(layer.scale_ref_layout(), ref_s) < TOL_FP32     # the scale is kept in FP32 in every mode
    dx = cabi.download_act(layer.backward(xb, dyb), dtype, B, C, S, S)
    # the kernel reads the stored (rounded) output back, so does the checker
    assert rel_err(dx, lo.lrn_backward(x, y, dy, ref_s, rng_, alpha, beta)) < tol
    # derivative hook of the layer in front (leaky ReLU on its stored output, here x itself)
    act = cabi.activ(cabi.RELU)
    dxh = cabi.download_act(layer.backward(xb, dyb, prev_act=act, prev_out=xb), dtype, B, C, S, S)
    assert rel_err(dxh, co.relu_deriv(lo.lrn_backward(x, y, dy, ref_s, rng_, alpha, beta), x, B)) < tol
    # inference: no scale buffer
    y2 = cabi.download_act(layer.forward(xb, keep_scale=False), dtype, B, C, S, S)
    assert np.array_equal(y, y2)


_lrn_net = netdefs.lrn_net


def _batch(spec, seed):
    rng = np.random.default_rng(seed)
    B = spec["batch"]
    dim = spec["in_dim"][0] * spec["in_dim"][1] * spec["in_ch"]
    x = np.zeros((B, dim + 1), np.float32)
    x[:, :dim] = rng.standard_normal((B, dim)).astype(np.float32)
    t = np.zeros((B, spec["out_dim"]), np.float32)
    t[np.arange(B), rng.integers(0, spec["out_dim"], B)] = 1
    return x, t


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_lrn_layer_inside_a_network(cnn, mode, tmp_path, monkeypatch):
    """host layer: forward / backward through python lrn(), checked layer-locally against the oracle on the tensors read
    back around it; then the save -> load round trip in both file formats (src/lrn_layer.c:216-277)."""
    monkeypatch.chdir(tmp_path)
    spec = _lrn_net()
    B = spec["batch"]
    tol = TOL_FP32 if mode == "off" else TOL_MIXED
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)
    x, t = _batch(spec, 5)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    loss0 = cnn.batch_loss(network=0)
    assert np.isfinite(loss0)
    for conv_i, lrn_i, (r, k, a, b) in ((0, 1, (5, 2.0, 0.3, 0.75)), (3, 4, (5, 1.0, 1.0, 0.5))):
        xin, y = cnn.layer_output(conv_i, network=0), cnn.layer_output(lrn_i, network=0)
        ref_y, ref_s = lo.lrn_forward(xin, r, k, a, b)
        assert rel_err(y, ref_y) < tol
    cnn.backward_batch(0.0, network=0)          # lr 0: deltas only
    for conv_i, lrn_i, (r, k, a, b) in ((0, 1, (5, 2.0, 0.3, 0.75)), (3, 4, (5, 1.0, 1.0, 0.5))):
        xin, y = cnn.layer_output(conv_i, network=0), cnn.layer_output(lrn_i, network=0)
        _, ref_s = lo.lrn_forward(xin, r, k, a, b)
        dy, dx = cnn.layer_delta(lrn_i, network=0), cnn.layer_delta(conv_i, network=0)
        assert np.abs(dy).max() > 0
        assert rel_err(dx, co.relu_deriv(lo.lrn_backward(xin, y, dy, ref_s, r, a, b), xin, B)) < tol
    # a few SGD steps on the same batch must fit it
    for _ in range(30):
        cnn.load_batch(x, t, network=0)
        cnn.forward_batch(network=0)
        cnn.backward_batch(0.05, 0.9, network=0)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    assert cnn.batch_loss(network=0) < 0.5 * loss0
    out_before = cnn.layer_output(5, network=0)
    weights = {i: cnn.layer_weights(i, network=0) for i in (0, 3, 5)}
    for use_bin in (1, 0):
        name = "lrn_net.dat" if use_bin else "lrn_net.txt"
        with rd._Quiet():
            cnn.save(name, network=0, bin=use_bin)
            cnn.init(in_dim=rd.i_ar(spec["in_dim"]), in_nb_ch=3, out_dim=5, bias=0.1, b_size=B, comp_meth="C_CUDA",
                     mixed_precision=mode, no_logo=1, network=0)
            cnn.load(name, 0, network=0, bin=use_bin)
        assert cnn.layer_shape(1, network=0)[3] == 4 and cnn.layer_shape(4, network=0)[3] == 4      # LRN
        for i, w in weights.items():
            if use_bin:
                assert np.array_equal(cnn.layer_weights(i, network=0), w)
            else:
                assert rel_err(cnn.layer_weights(i, network=0), w) < 1e-5      # text format keeps ~6 digits
        cnn.load_batch(x, t, network=0)
        cnn.forward_batch(network=0)
        assert rel_err(cnn.layer_output(5, network=0), out_before) < (1e-6 if use_bin else 1e-3)


ARCH_NETS = {"mini_darknet": lambda: netdefs.mini_darknet(), "lenet": lambda: netdefs.lenet(batch=4)}
ARCH_SELECTIONS = {
    "default": {},
    "all": dict(size=1, in_size=1, f_size=1, out_size=1, stride=1, padding=1, in_padding=1, activation=1, bias=1, dropout=1,
                param_count=1),
    "sparse": dict(size=0, in_size=1, f_size=0, out_size=1, stride=0, padding=0, activation=1, param_count=1),
}


@pytest.mark.parametrize("net_name", sorted(ARCH_NETS))
def test_print_arch_tex_equals_reference_file(cnn, net_name, tmp_path):
    with rd._Quiet():
        rd.build_network(cnn, ARCH_NETS[net_name](), "C_CUDA", "off", network=0)
    for sel, kw in ARCH_SELECTIONS.items():
        with rd._Quiet():
            cnn.print_arch_tex(str(tmp_path) + "/", "arch_" + sel, network=0, **kw)
        with open(os.path.join(str(tmp_path), "arch_%s.tex" % sel), "rb") as f:
            mine = f.read()
        with open(os.path.join(GOLDEN_DIR, "arch_tex", "%s_%s.tex" % (net_name, sel)), "rb") as f:
            theirs = f.read()
        assert mine == theirs, (net_name, sel)


def test_print_arch_tex_lrn_row(cnn, tmp_path):
    """no reference file can exist for an LRN network (CUDA-only layer upstream): check the row against src/auxil.c:1019-1056"""
    with rd._Quiet():
        rd.build_network(cnn, _lrn_net(), "C_CUDA", "off", network=0)
        cnn.print_arch_tex(str(tmp_path) + "/", "arch", network=0, activation=1, dropout=1)
    rows = open(os.path.join(str(tmp_path), "arch.tex")).read().splitlines()
    assert "2 & LRN\\_1 & 12x12x1 & ch\\_range: 5& & & & 12x12x1 & LIN & \\\\" in rows
    assert "5 & LRN\\_2 & 6x6x1 & ch\\_range: 5& & & & 6x6x1 & LIN & \\\\" in rows
