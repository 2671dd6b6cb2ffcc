"""GPU parity, operator level: every kernel family is called through the C-ABI (include/cianna_b200.h) with
host buffers and checked against the oracle (oracle/cianna_oracle.py, pinned by tests/test_oracle.py).

Tolerances (BASELINE.json north_star): bit-exact for the implicit-GEMM address mapping (integer-valued
tensors) and pool argmax; 1e-5 relative for FP32 activations / gradients; 2e-2 for mixed precision.
"""
import json
import os

import numpy as np
import pytest

from oracle import cianna_oracle as co
from tests.common import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-5
TOL_MIXED = 2e-2


@pytest.fixture(scope="module")
def cabi():
    from cianna_b200 import cabi as m
    m.init_device(0)
    return m


def _int_tensor(rng, shape, p_zero=0.5):
    v = rng.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=shape, p=[(1 - p_zero) / 2, p_zero, (1 - p_zero) / 2])
    return v.astype(np.float32)


# geometry: (batch, in_c, size, out_c, f, pad)
TC_SHAPES = [
    (8, 32, 14, 64, 3, 1),     # BK=32 (64B swizzle), BN=64
    (3, 64, 13, 48, 3, 1),     # odd image size / batch, partial tiles in every dimension
    (8, 64, 14, 256, 3, 1),    # BN=256
    (4, 128, 28, 128, 1, 0),   # 1x1, two channel blocks
    (16, 16, 7, 16, 3, 1),     # BK=16 (32B swizzle), BN=16
    (2, 256, 14, 1000, 1, 0),  # the Darknet19 head: N=1000 -> four 256-wide tiles, last one partial
    (5, 64, 10, 32, 5, 2),     # 5x5 filter
    (128, 64, 1, 40, 1, 0),    # dense-like: 1x1 map, tile spans 128 images
    (4, 32, 12, 32, 1, 0),     # 32 -> 32 channels: the shape of a first layer run on patch rows (weight-gradient slab wider than the tensor)
    (2, 16, 9, 24, 3, 1),      # 16-channel operands everywhere
    (65, 64, 28, 128, 1, 0),   # 399 M tiles (odd): enough work for the 2-CTA cluster variant (multicast filter halves, dummy last tile)
    (49, 64, 28, 256, 3, 1),   # cluster variant with BN=256, 3x3, 301 M tiles; CTA-pair kernel (forward)
    (49, 256, 28, 256, 1, 0),  # CTA-pair kernel forward and data gradient (256 channels both sides), odd number of M tiles
    (50, 192, 28, 512, 1, 0),  # CTA-pair kernel: two N tiles of 256, partial last K block (192 channels = 3 blocks of 64)
    (49, 256, 28, 256, 3, 1),  # CTA-pair kernels with 3x3 filters: weight gradient with two taps per pair (last group: one)
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tcgen05_address_mapping_bit_exact(cabi, shape, dtype_name):
    """integer-valued inputs / weights: every product and partial sum is exactly representable, so the
    tensor-core implicit GEMM must reproduce the reference im2col + GEMM bit for bit (forward, data
    gradient and weight gradient)."""
    B, C, S, N, f, pad = shape
    dtype = cabi.FP16 if dtype_name == "FP16" else cabi.BF16
    rng = np.random.default_rng(hash(shape) % 2**31)
    x = _int_tensor(rng, (C, B, S * S), 0.6)
    w = _int_tensor(rng, (N, f * f * C + 1), 0.7)
    w[:, -1] = rng.integers(-2, 3, N)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, 1, pad, bias_value=1.0)
    layer.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = S + 2 * pad - f + 1
    cabi.lib().cb200_force_simt(16)          # one-SM kernels first (the CTA-pair kernel is checked against the same bits below)
    y = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
    assert cabi.lib().cb200_last_conv_impl() == b"tcgen05"
    ref, col = co.conv_forward(x, w, False, B, C, S, S, f, 1, pad, 1.0)
    assert np.abs(ref).max() < 256, "test values must stay exactly representable in bf16"
    assert np.array_equal(y, ref)

    dy = _int_tensor(rng, (N, B, So * So), 0.8)
    dyb = cabi.upload_act(dy, dtype, B, N, So, So)
    dx = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
    assert cabi.lib().cb200_last_conv_impl() == b"tcgen05"
    cabi.lib().cb200_force_simt(0)
    ref_dx = co.conv_backward_data(dy, w, B, C, S, S, f, 1, pad)
    assert np.abs(ref_dx).max() < 256
    assert np.array_equal(dx, ref_dx)

    layer.backward_weights(xb, dyb)
    got = layer.grad_ref_layout()
    assert cabi.lib().cb200_last_conv_impl() == b"tcgen05"
    ref_g = co.conv_weight_grad(col, dy).astype(np.float32)
    assert np.array_equal(got, ref_g)
    if C > 128 and N > 128:
        cabi.lib().cb200_force_simt(8)           # the weight gradient on CTA pairs (cta_group::2; optional): same bits
        try:
            layer.backward_weights(xb, dyb)
            assert cabi.lib().cb200_last_conv_impl() == b"tcgen05-pair"
            assert np.array_equal(layer.grad_ref_layout(), ref_g)
        finally:
            cabi.lib().cb200_force_simt(0)
    if B * So * So >= 128 * 2 * 148 and N >= 96:
        # enough M tiles for the optional 2-CTA cluster variant (filter halves multicast to both CTAs): same bits
        cabi.lib().cb200_force_simt(4)
        try:
            y2 = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
        finally:
            cabi.lib().cb200_force_simt(0)
        assert np.array_equal(y2, ref)
    if B * So * So >= 128 * 2 * 148 and N > 128:
        # the CTA-pair kernel (cta_group::2, M = 256 over two SMs, half of the filter block per SM): same bits
        cabi.lib().cb200_force_simt(8)
        try:
            y3 = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
            assert cabi.lib().cb200_last_conv_impl() == b"tcgen05-pair"
            if C > 128:
                dx3 = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
                assert cabi.lib().cb200_last_conv_impl() == b"tcgen05-pair"
        finally:
            cabi.lib().cb200_force_simt(0)
        assert np.array_equal(y3, ref)
        if C > 128:
            assert np.array_equal(dx3, ref_dx)
    layer.free(); xb.free(); dyb.free()


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", [(3, 32, 12, 40, 3, 1), (2, 64, 40, 64, 3, 1), (4, 64, 9, 256, 1, 0)])
@pytest.mark.parametrize("sat", [800.0, 6.0, 2.5])
def test_conv_relu_epilogue_saturation_bit_exact(cabi, shape, sat, dtype_name):
    """The leaky-saturated ReLU fused in the tensor-core epilogues (per-tap and halo kernels) runs as max(z, z*leak) with a
    packed check for elements above the saturation and falls back to the select form for those: integer pre-activations
    and leak = 1/4 make every branch exact, so the output must equal the oracle bit for bit - with the saturation far away
    (800, upstream's value: fast path only), in the middle of the value range (6) and between two integers (2.5: every
    element from 3 up takes the saturated branch)."""
    B, C, S, N, f, pad = shape
    dtype = cabi.FP16 if dtype_name == "FP16" else cabi.BF16
    rng = np.random.default_rng(31)
    x = _int_tensor(rng, (C, B, S * S), 0.6)
    w = _int_tensor(rng, (N, f * f * C + 1), 0.7)
    w[:, -1] = rng.integers(-2, 3, N)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, 1, pad, bias_value=1.0, act=cabi.activ(cabi.RELU, 0.25, sat), length=B - 1)
    layer.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = S + 2 * pad - f + 1
    y = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
    assert cabi.lib().cb200_last_conv_impl().startswith(b"tcgen05")
    pre, _ = co.conv_forward(x, w, False, B, C, S, S, f, 1, pad, 1.0)
    assert np.abs(pre).max() < 64
    ref = co.relu_forward(pre, B - 1, saturation=sat, leak=0.25)
    if sat < 800:
        assert (pre > sat).mean() > 0.01, "the test must reach the saturated branch"
    # quarter-integers up to 64 are exact in both storage types
    assert np.array_equal(y, ref), "%d wrong" % int((y != ref).sum())
    layer.free(); xb.free()


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(4, 3, 16, 32, 3, 1, 1), (3, 1, 12, 8, 5, 2, 1), (2, 3, 9, 16, 3, 0, 2)])
def test_first_layer_patch_rows(cabi, cfg, dtype_name):
    """first layer on few input channels: the layout import unrolls receptive fields (reference column order, bias
    input as a column) and the layer runs as a 1x1 GEMM; forward, weight gradient and SGD update against the oracle"""
    import ctypes
    B, C, S, N, f, pad, stride = cfg
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    L = cabi.lib()
    L.cb200_patch_width.restype = ctypes.c_int
    L.cb200_import_input_patches.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 13 + [ctypes.c_float, ctypes.c_void_p]
    rng = np.random.default_rng(9)
    x = np.empty((B, C * S * S + 1), np.float32)
    x[:, :-1] = rng.integers(-3, 4, (B, C * S * S)) / 4.0          # exactly representable in every dtype
    x[:, -1] = 0.5
    w = (rng.integers(-4, 5, (N, f * f * C + 1)) / 8.0).astype(np.float32)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, stride, pad, bias_value=0.5, act=cabi.activ(cabi.RELU))
    layer.d.input_is_patches = 1
    layer.set_weights(w)
    So = (S + 2 * pad - f) // stride + 1
    kp = L.cb200_patch_width(C, f, f)
    es = L.cb200_dtype_size(dtype)
    xt = np.empty(x.size, np.uint16 if es == 2 else np.float32)
    cabi.check(L.cb200_host_cast_from_f32(xt.ctypes.data, dtype, x.ctypes.data, x.size))
    src = cabi.DevBuf.from_numpy(xt)
    patches = cabi.DevBuf(B * So * So * kp * es)
    cabi.check(L.cb200_import_input_patches(patches.ptr, src.ptr, dtype, B, C, S, S, f, f, stride, stride, pad, pad, So, So, 0.5, None))
    y = cabi.download_act(layer.forward(patches), dtype, B, N, So, So)
    pre, col = co.conv_forward(x, w, True, B, C, S, S, f, stride, pad, 0.5)
    assert rel_err(y, co.relu_forward(pre, B)) < tol
    dy = (rng.integers(-4, 5, (N, B, So * So)) / 4.0).astype(np.float32)
    dyb = cabi.upload_act(dy, dtype, B, N, So, So)
    layer.backward_weights(patches, dyb)
    g = layer.bufs["grad"].to_numpy(np.float32, (N, kp))[:, : f * f * C + 1]
    assert rel_err(g, co.conv_weight_grad(col, dy)) < tol
    assert np.array_equal(layer.bufs["grad"].to_numpy(np.float32, (N, kp))[:, f * f * C + 1:], np.zeros((N, kp - f * f * C - 1), np.float32))


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("shape", [(4, 3, 16, 8, 3, 1, 1), (2, 8, 9, 12, 3, 0, 2), (3, 5, 12, 7, 5, 2, 1), (2, 16, 8, 16, 2, 0, 2)])
def test_conv_simt_matches_oracle(cabi, shape, dtype_name):
    """generic kernels (FP32 mode, strides, tiny channel counts) against the oracle on random data"""
    B, C, S, N, f, pad, stride = shape
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    rng = np.random.default_rng(1)
    x = rng.standard_normal((C, B, S * S)).astype(np.float32)
    w = (rng.standard_normal((N, f * f * C + 1)) * 0.2).astype(np.float32)
    cabi.lib().cb200_force_simt(1)
    try:
        layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, stride, pad, bias_value=0.1, act=cabi.activ(cabi.RELU))
        layer.set_weights(w)
        xb = cabi.upload_act(x, dtype, B, C, S, S)
        So = (S + 2 * pad - f) // stride + 1
        y = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
        pre, col = co.conv_forward(x, w, False, B, C, S, S, f, stride, pad, 0.1)
        ref = co.relu_forward(pre, B)
        assert rel_err(y, ref) < tol
        dy = rng.standard_normal((N, B, So * So)).astype(np.float32)
        dyb = cabi.upload_act(dy, dtype, B, N, So, So)
        dx = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
        assert rel_err(dx, co.conv_backward_data(dy, w, B, C, S, S, f, stride, pad)) < tol
        layer.backward_weights(xb, dyb)
        assert rel_err(layer.grad_ref_layout(), co.conv_weight_grad(col, dy)) < tol
    finally:
        cabi.lib().cb200_force_simt(0)


@pytest.mark.parametrize("shape", [(8, 64, 14, 128, 3, 1), (4, 32, 28, 64, 3, 1), (6, 128, 7, 64, 1, 0)])
def test_conv_tcgen05_vs_simt_random(cabi, shape):
    """same 16-bit operands through both kernel families: only the FP32 summation order may differ"""
    B, C, S, N, f, pad = shape
    rng = np.random.default_rng(2)
    x = rng.standard_normal((C, B, S * S)).astype(np.float32)
    w = (rng.standard_normal((N, f * f * C + 1)) * 0.1).astype(np.float32)
    outs = []
    for force in (0, 1):
        cabi.lib().cb200_force_simt(force)
        layer = cabi.ConvLayer(cabi.FP16, B, C, S, S, N, f, 1, pad, bias_value=0.1, act=cabi.activ(cabi.RELU))
        layer.set_weights(w)
        xb = cabi.upload_act(x, cabi.FP16, B, C, S, S)
        So = S + 2 * pad - f + 1
        y = cabi.download_act(layer.forward(xb), cabi.FP16, B, N, So, So)
        dy = rng.standard_normal((N, B, So * So)).astype(np.float32) if force == 0 else dy
        dyb = cabi.upload_act(dy, cabi.FP16, B, N, So, So)
        dx = cabi.download_act(layer.backward_data(dyb, cabi.activ(cabi.RELU), xb), cabi.FP16, B, C, S, S)
        layer.backward_weights(xb, dyb)
        outs.append((y, dx, layer.grad_ref_layout(), cabi.lib().cb200_last_conv_impl()))
    cabi.lib().cb200_force_simt(0)
    assert outs[0][3] == b"tcgen05" and outs[1][3] == b"simt"
    assert rel_err(outs[0][0], outs[1][0]) < 2e-3
    assert rel_err(outs[0][1], outs[1][1]) < 2e-3
    assert rel_err(outs[0][2], outs[1][2]) < 1e-4


def test_pool_argmax_bit_exact_vs_golden(cabi):
    """max-pool on the reference's own input tensor: values and argmax map must be identical"""
    g = load_golden("mini_darknet_blas")
    x = g["out_1"]                      # GN output feeding the first max pool
    C, B, A = x.shape
    S = int(round(A ** 0.5))
    pool = cabi.PoolLayer(cabi.FP32, B, C, S, S, 2)
    xb = cabi.upload_act(x, cabi.FP32, B, C, S, S)
    y = cabi.download_act(pool.forward(xb), cabi.FP32, B, C, S // 2, S // 2)
    assert np.array_equal(y, g["out_2"])
    assert np.array_equal(pool.map_ref_layout(), g["map_2"])


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(2, 2, 0, "MAX"), (3, 2, 1, "MAX"), (3, 1, 1, "MAX"), (2, 2, 0, "AVG"), (3, 2, 1, "AVG")])
def test_pool_forward_backward_vs_oracle(cabi, cfg, dtype_name):
    """window / stride / padding variants incl. ties (values quantised to halves, exactly representable in every
    dtype) - first maximum must win, argmax bit-exact in all three storage types"""
    p, s, pad, kind = cfg
    B, C, S = 3, 10, 9
    DT = getattr(cabi, dtype_name)
    rng = np.random.default_rng(5)
    x = np.round(rng.standard_normal((C, B, S * S)) * 2).astype(np.float32) / 2      # many exact ties
    pool = cabi.PoolLayer(DT, B, C, S, S, p, s, pad, cabi.POOL_MAX if kind == "MAX" else cabi.POOL_AVG)
    xb = cabi.upload_act(x, DT, B, C, S, S)
    So = (S + 2 * pad - p) // s + 1
    y = cabi.download_act(pool.forward(xb), DT, B, C, So, So)
    ref_y, ref_m = co.pool_forward(x, B, C, S, S, p, s, pad, kind)
    if kind == "MAX":
        assert np.array_equal(y, ref_y)
        assert np.array_equal(pool.map_ref_layout(), ref_m)
    else:
        assert rel_err(y, ref_y) < (TOL_FP32 if dtype_name == "FP32" else TOL_MIXED)
    dy = (np.round(rng.standard_normal((C, B, So * So)) * 8) / 8).astype(np.float32)
    dyb = cabi.upload_act(dy, DT, B, C, So, So)
    dx = cabi.download_act(pool.backward(dyb), DT, B, C, S, S)
    assert rel_err(dx, co.pool_backward(dy, ref_m, B, C, S, S, p, s, pad, kind)) < (TOL_FP32 if dtype_name == "FP32" else TOL_MIXED)


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(4, 32, 12, 4, 0, 4), (3, 16, 7, 8, 1, 2), (2, 64, 40, 16, 0, 2), (2, 8, 5, 8, 0, 2)])
def test_group_norm_vs_oracle(cabi, cfg, dtype_name):
    B, C, S, gs, set_off, length = cfg
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((C, B, S * S)) * 1.5 + 0.7).astype(np.float32)
    if dtype_name != "FP32":   # make the input exactly representable so both sides see the same tensor
        x = x.astype(np.float16).astype(np.float32) if dtype_name == "FP16" else (x.view(np.uint32) & 0xFFFF0000).view(np.float32)
    G = C // gs
    gamma = (1 + 0.3 * rng.standard_normal(G)).astype(np.float32)
    beta = (0.2 * rng.standard_normal(G)).astype(np.float32)
    nl = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
    nl.set_params(gamma, beta)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    y = cabi.download_act(nl.forward(xb), dtype, B, C, S, S)
    ref_y, mean, var = co.group_norm_forward(x, gamma, beta, gs, set_off, length)
    assert rel_err(y, ref_y) < tol
    m, v, _, _ = nl.stats()
    assert rel_err(m, mean) < TOL_FP32 * 5 and rel_err(v, var) < TOL_FP32 * 5
    dy = rng.standard_normal((C, B, S * S)).astype(np.float32)
    dy[:, length:, :] = 0
    if dtype_name != "FP32":
        dy = dy.astype(np.float16).astype(np.float32) if dtype_name == "FP16" else (dy.view(np.uint32) & 0xFFFF0000).view(np.float32)
    dyb = cabi.upload_act(dy, dtype, B, C, S, S)
    dx = cabi.download_act(nl.backward(xb, dyb), dtype, B, C, S, S)
    ref_dx, dgam, dbet = co.group_norm_backward(x, dy, gamma, mean, var, gs, set_off, length)
    assert rel_err(dx, ref_dx) < tol
    _, _, dg, db = nl.stats()
    assert rel_err(dg, dgam) < 1e-4 and rel_err(db, dbet) < 1e-4
    # fused by-product: per-channel sums of dx (bias-column gradient of the convolution in front of the norm layer)
    cs = nl.colsum.to_numpy(np.float32, (C,))
    assert rel_err(cs, ref_dx.astype(np.float64).sum(axis=(1, 2))) < (1e-4 if dtype_name == "FP32" else TOL_MIXED)


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(4, 32, 12, 4, 0, 4, "RELU"), (3, 16, 8, 8, 1, 2, "LIN"), (2, 64, 40, 16, 0, 2, "RELU"), (2, 8, 6, 8, 0, 2, "RELU"), (2, 20, 10, 4, 0, 2, "LIN")])
def test_group_norm_max_pool_fused_equals_unfused(cabi, cfg, dtype_name):
    """cb200_norm_pool_forward / _backward against cb200_norm_forward + cb200_pool_forward and cb200_pool_backward +
    cb200_norm_backward on the same tensors: same pooled values and argmax map (the fused kernel rounds the normalised
    values to the storage type before comparing them, like the un-fused pair does by storing them), statistics and dx
    equal to accumulation order"""
    B, C, S, gs, set_off, length, prev = cfg
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((C, B, S * S)) * 1.5 + 0.3).astype(np.float32)
    x[:, :, ::7] = np.round(x[:, :, ::7])            # some exact ties inside windows
    G = (C + gs - 1) // gs
    gamma = (1 + 0.3 * rng.standard_normal(G)).astype(np.float32)
    beta = (0.2 * rng.standard_normal(G)).astype(np.float32)
    pa = cabi.activ(cabi.RELU) if prev == "RELU" else None
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = S // 2
    dp = rng.standard_normal((C, B, So * So)).astype(np.float32)
    dp[:, length:, :] = 0
    dpb = cabi.upload_act(dp, dtype, B, C, So, So)
    # un-fused pair
    n1 = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
    n1.set_params(gamma, beta)
    p1 = cabi.PoolLayer(dtype, B, C, S, S, 2, 2, 0, cabi.POOL_MAX, length=length)
    y1 = cabi.download_act(p1.forward(n1.forward(xb)), dtype, B, C, So, So)
    m1 = p1.map_ref_layout()
    dyb = p1.backward(dpb)
    dx1 = cabi.download_act(n1.backward(xb, dyb, pa), dtype, B, C, S, S)
    st1 = n1.stats()
    cs1 = n1.colsum.to_numpy(np.float32, (C,))
    # fused pair
    n2 = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
    n2.set_params(gamma, beta)
    p2 = cabi.PoolLayer(dtype, B, C, S, S, 2, 2, 0, cabi.POOL_MAX, length=length)
    y2 = cabi.download_act(n2.forward_pool(xb, p2), dtype, B, C, So, So)
    m2 = p2.map_ref_layout()
    dx2 = cabi.download_act(n2.backward_pool(xb, dpb, p2, pa), dtype, B, C, S, S)
    st2 = n2.stats()
    cs2 = n2.colsum.to_numpy(np.float32, (C,))
    # (the two runs accumulate their statistics with atomics, so mean / var may differ in the last bit between them: the
    #  pooled values then agree to one unit of the storage type and a map entry can only move between two window values
    #  that are equal to within that bit)
    ulp = {"FP32": 2.0 ** -22, "FP16": 2.0 ** -10, "BF16": 2.0 ** -7}[dtype_name]
    assert rel_err(y2, y1) <= ulp, "pooled values differ"
    assert (m1 != m2).mean() < 2e-3, "argmax map differs"
    for a, b in zip(st1, st2):
        assert rel_err(b, a) < 1e-5
    assert rel_err(dx2, dx1) < (1e-5 if dtype_name == "FP32" else tol)
    assert rel_err(cs2, cs1) < (1e-4 if dtype_name == "FP32" else tol)
    # backward reductions from the pooled delta and the pooled OUTPUT (cb200_norm_pool_backward_ex: x = (y - shift) / scale
    # at the selected position, one rounding of y to the storage type instead of one of x): same d_gamma / d_beta / dx
    dx3 = cabi.download_act(n2.backward_pool(xb, dpb, p2, pa, from_pooled_output=True), dtype, B, C, S, S)
    st3 = n2.stats()
    for a, b in zip(st1, st3):
        assert rel_err(b, a) < (1e-5 if dtype_name == "FP32" else tol)
    assert rel_err(dx3, dx1) < (1e-5 if dtype_name == "FP32" else tol)
    assert rel_err(n2.colsum.to_numpy(np.float32, (C,)), cs1) < (1e-4 if dtype_name == "FP32" else tol)
    # ... groups where that inversion is ill-conditioned (tiny gamma, |beta| >> |gamma|) keep the gather from x
    g_bad, b_bad = gamma.copy(), beta.copy()
    g_bad[0] = 3e-5
    b_bad[-1] = 40.0 * abs(g_bad[-1])
    outs = []
    for from_y in (False, True):
        n4 = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
        n4.set_params(g_bad, b_bad)
        p4 = cabi.PoolLayer(dtype, B, C, S, S, 2, 2, 0, cabi.POOL_MAX, length=length)
        n4.forward_pool(xb, p4)
        dx4 = cabi.download_act(n4.backward_pool(xb, dpb, p4, pa, from_pooled_output=from_y), dtype, B, C, S, S)
        outs.append((dx4, n4.stats()))
    assert rel_err(outs[1][0], outs[0][0]) < (1e-5 if dtype_name == "FP32" else tol)
    for a, b in zip(outs[0][1], outs[1][1]):
        assert rel_err(b, a) < (1e-5 if dtype_name == "FP32" else tol)
    # and against the oracle, end to end
    ref_y, mean, var = co.group_norm_forward(x if dtype_name == "FP32" else cabi.download_act(xb, dtype, B, C, S, S), gamma, beta, gs, set_off, length)
    ref_p, ref_m = co.pool_forward(ref_y, B, C, S, S, 2, 2, 0, "MAX")
    assert rel_err(y2, ref_p) < tol
    if dtype_name == "FP32":
        assert (m2 != ref_m).mean() < 1e-3      # only exact FP32 ties may resolve differently after rounding


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(7, 32, 24, 16, 0, 7, 96, 3), (5, 64, 16, 16, 0, 4, 40, 2), (9, 24, 12, 8, 1, 9, 16, 3), (3, 136, 10, 8, 0, 3, 4096, 3)])
def test_group_norm_pipelined_equals_two_launch(cabi, cfg, dtype_name):
    """The chunked launches (launch i = statistics blocks of chunk i next to apply blocks of chunk i-1, mean / var and
    d_gamma / d_beta finalised inside the apply blocks) against the statistics / finalize / apply launches over the whole
    batch, plain and fused with the max-pool, forward and backward: chunks of one to a few samples, ragged last chunk,
    dead samples, one chunk for the whole batch.  Same arithmetic per block; only the order of the atomic accumulation
    differs."""
    B, C, S, gs, set_off, length, chunk_kb, ctas = cfg
    dtype = getattr(cabi, dtype_name)
    tol = TOL_FP32 if dtype_name == "FP32" else TOL_MIXED
    L = cabi.lib()
    rng = np.random.default_rng(23)
    x = (rng.standard_normal((C, B, S * S)) * 1.5 + 0.3).astype(np.float32)
    G = (C + gs - 1) // gs
    gamma = (1 + 0.3 * rng.standard_normal(G)).astype(np.float32)
    beta = (0.2 * rng.standard_normal(G)).astype(np.float32)
    pa = cabi.activ(cabi.RELU)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = S // 2
    dy = rng.standard_normal((C, B, S * S)).astype(np.float32)
    dy[:, length:, :] = 0
    dyb = cabi.upload_act(dy, dtype, B, C, S, S)
    dp = rng.standard_normal((C, B, So * So)).astype(np.float32)
    dp[:, length:, :] = 0
    dpb = cabi.upload_act(dp, dtype, B, C, So, So)
    res = []
    try:
        for on in (0, 1):
            L.cb200_norm_set_pipeline(on, chunk_kb, ctas)
            n = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
            n.set_params(gamma, beta)
            y = cabi.download_act(n.forward(xb), dtype, B, C, S, S)
            dx = cabi.download_act(n.backward(xb, dyb, pa), dtype, B, C, S, S)
            st = n.stats()
            cs = n.colsum.to_numpy(np.float32, (C,))
            n2 = cabi.NormLayer(dtype, B, C, S, S, gs, set_off, length)
            n2.set_params(gamma, beta)
            p2 = cabi.PoolLayer(dtype, B, C, S, S, 2, 2, 0, cabi.POOL_MAX, length=length)
            yp = cabi.download_act(n2.forward_pool(xb, p2), dtype, B, C, So, So)
            mp = p2.map_ref_layout()
            dxp = cabi.download_act(n2.backward_pool(xb, dpb, p2, pa), dtype, B, C, S, S)
            stp = n2.stats()
            csp = n2.colsum.to_numpy(np.float32, (C,))
            res.append((y, dx, st, cs, yp, mp, dxp, stp, csp))
    finally:
        L.cb200_norm_set_pipeline(0, 24 * 1024, 6)
    (y0, dx0, st0, cs0, yp0, mp0, dxp0, stp0, csp0), (y1, dx1, st1, cs1, yp1, mp1, dxp1, stp1, csp1) = res
    # (mean / var of the two runs may differ in the last bit: a couple of units of the storage type on y)
    ulp = {"FP32": 2.0 ** -21, "FP16": 2.0 ** -10, "BF16": 2.0 ** -7}[dtype_name]
    assert rel_err(y1, y0) <= ulp and rel_err(yp1, yp0) <= ulp
    assert (mp0 != mp1).mean() < 2e-3
    for a, b in zip(st0 + stp0, st1 + stp1):
        assert rel_err(b, a) < 1e-5
    assert rel_err(dx1, dx0) < (1e-5 if dtype_name == "FP32" else tol)
    assert rel_err(dxp1, dxp0) < (1e-5 if dtype_name == "FP32" else tol)
    assert rel_err(cs1, cs0) < (1e-4 if dtype_name == "FP32" else tol)
    assert rel_err(csp1, csp0) < (1e-4 if dtype_name == "FP32" else tol)
    # and the pipelined forward against the oracle
    xq = x if dtype_name == "FP32" else cabi.download_act(xb, dtype, B, C, S, S)
    ref_y, _, _ = co.group_norm_forward(xq, gamma, beta, gs, set_off, length)
    assert rel_err(y1, ref_y) < tol


# geometry: (batch, in_c, height, width, out_c, f, pad, stride)
FIRST_DIRECT = [
    (4, 3, 16, 20, 32, 3, 1, 1),    # the Darknet19 first layer in small: RGB, 3x3, 32 filters (KP=32, 64B swizzle)
    (3, 3, 13, 11, 24, 3, 1, 1),    # odd sizes, partial tiles, filters not a multiple of 32
    (2, 1, 28, 28, 16, 5, 2, 1),    # grey 5x5 (KP=32)
    (5, 1, 12, 12, 8, 3, 0, 1),     # grey 3x3, no padding (KP=16, 32B swizzle)
    (2, 2, 10, 14, 64, 3, 1, 1),    # two channels, 64 filters (BN=64)
    (3, 3, 17, 17, 40, 3, 1, 2),    # stride 2
    (130, 3, 4, 4, 32, 3, 1, 1),    # tiles spanning many images
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("cfg", FIRST_DIRECT)
def test_first_layer_direct_bit_exact(cabi, cfg, dtype_name):
    """input_is_patches = 2: the first layer reads the dataset batch and builds its patch rows in shared memory
    (conv_first.cu). Integer-valued tensors -> every product and partial sum is exact: forward and weight gradient must
    equal the im2col oracle BIT FOR BIT (receptive-field addressing, zero padding, swizzled operand layout, bias column)."""
    import ctypes
    B, C, H, W, N, f, pad, stride = cfg
    dtype = getattr(cabi, dtype_name)
    L = cabi.lib()
    L.cb200_patch_width.restype = ctypes.c_int
    rng = np.random.default_rng(13)
    x = np.empty((B, C * H * W + 1), np.float32)
    x[:, :-1] = _int_tensor(rng, (B, C * H * W), 0.3)
    x[:, -1] = 1.0
    w = _int_tensor(rng, (N, f * f * C + 1), 0.4)
    Ho, Wo = (H + 2 * pad - f) // stride + 1, (W + 2 * pad - f) // stride + 1
    d = cabi.ConvDesc(dtype, B, B, C, H, W, N, Ho, Wo, f, f, stride, stride, pad, pad, 1.0, cabi.activ(cabi.LINEAR), 1)
    assert L.cb200_conv_first_direct(ctypes.byref(d)) == 1
    d.input_is_patches = 2
    kp = L.cb200_patch_width(C, f, f)
    es = L.cb200_dtype_size(dtype)
    bufs = dict(master=cabi.DevBuf.from_numpy(w), moment=cabi.DevBuf(w.nbytes), w_fwd=cabi.DevBuf(N * kp * es),
                w_bwd=cabi.DevBuf(C * f * f * cabi.round8(N) * es), bias_w=cabi.DevBuf(N * 4), grad=cabi.DevBuf(N * kp * 4), grad_b=cabi.DevBuf(N * 4))
    wts = cabi.ConvWeights(*[bufs[k].ptr for k in ("master", "moment", "w_fwd", "w_bwd", "bias_w", "grad", "grad_b")])
    cabi.check(L.cb200_conv_prepare_weights(ctypes.byref(d), ctypes.byref(wts), None))
    xt = np.empty(x.size, np.uint16)
    cabi.check(L.cb200_host_cast_from_f32(xt.ctypes.data, dtype, x.ctypes.data, x.size))
    src = cabi.DevBuf.from_numpy(xt)
    y = cabi.DevBuf(B * Ho * Wo * cabi.round8(N) * es)
    cabi.check(L.cb200_conv_forward(ctypes.byref(d), ctypes.byref(wts), src.ptr, y.ptr, None))
    assert L.cb200_last_conv_impl().decode() == "tcgen05"
    # oracle on a square-agnostic im2col: reuse the generic routine through explicit patches
    xi = x[:, :-1].reshape(B, C, H, W)
    xp = np.pad(xi, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    cols = np.empty((B, Ho * Wo, C * f * f + 1), np.float32)
    for oy in range(Ho):
        for ox in range(Wo):
            cols[:, oy * Wo + ox, :-1] = xp[:, :, oy * stride: oy * stride + f, ox * stride: ox * stride + f].reshape(B, -1)
    cols[:, :, -1] = 1.0
    ref = np.einsum("bpk,nk->nbp", cols.astype(np.float64), w.astype(np.float64)).astype(np.float32)
    got = cabi.download_act(y, dtype, B, N, Ho, Wo)
    assert np.array_equal(got, ref), "forward differs: max |d| = %g" % np.abs(got - ref).max()
    dy = _int_tensor(rng, (N, B, Ho * Wo), 0.6)
    dyb = cabi.upload_act(dy, dtype, B, N, Ho, Wo)
    cabi.check(L.cb200_conv_backward_weights(ctypes.byref(d), ctypes.byref(wts), src.ptr, dyb.ptr, None))
    g = bufs["grad"].to_numpy(np.float32, (N, kp))
    gref = np.einsum("nbp,bpk->nk", dy.astype(np.float64), cols.astype(np.float64)).astype(np.float32)
    assert np.array_equal(g[:, : C * f * f + 1], gref), "weight gradient differs: max |d| = %g" % np.abs(g[:, : C * f * f + 1] - gref).max()
    assert not g[:, C * f * f + 1:].any()
    for b in list(bufs.values()) + [src, y, dyb]:
        b.free()


def test_first_layer_direct_activation_and_tail(cabi):
    """ReLU epilogue, samples beyond `length` forced to zero, pad output channels zero"""
    import ctypes
    B, C, H, W, N, f, pad = 6, 3, 12, 12, 20, 3, 1
    dtype = cabi.FP16
    L = cabi.lib()
    L.cb200_patch_width.restype = ctypes.c_int
    rng = np.random.default_rng(14)
    x = np.empty((B, C * H * W + 1), np.float32)
    x[:, :-1] = (rng.standard_normal((B, C * H * W)) * 0.5).astype(np.float16).astype(np.float32)
    x[:, -1] = 0.1
    w = (rng.standard_normal((N, f * f * C + 1)) * 0.3).astype(np.float32)
    layer = cabi.ConvLayer(dtype, B, C, H, W, N, f, 1, pad, bias_value=0.1, act=cabi.activ(cabi.RELU), length=4)
    layer.d.input_is_patches = 2
    layer.set_weights(w)
    xt = np.empty(x.size, np.uint16)
    cabi.check(L.cb200_host_cast_from_f32(xt.ctypes.data, dtype, x.ctypes.data, x.size))
    src = cabi.DevBuf.from_numpy(xt)
    y = cabi.download_act(layer.forward(src), dtype, B, N, H, W)
    xq = x.copy(); xq[:, -1] = np.float32(np.float16(0.1))
    pre, _ = co.conv_forward(xq, w.astype(np.float16).astype(np.float32), True, B, C, H, H, f, 1, pad, float(np.float16(0.1)))
    assert rel_err(y, co.relu_forward(pre, 4)) < TOL_MIXED
    assert not y[:, 4:, :].any()


# geometry: (batch, in_c, size, out_c, f, pad) - maps large enough for the 8x16 halo tile
HALO_SHAPES = [
    (2, 32, 32, 64, 3, 1),     # the Darknet19 second layer in small: BK=32 (64B swizzle), BN=64
    (2, 64, 32, 128, 3, 1),    # third / fifth layer: BK=64 (128B swizzle), BN=128
    (1, 128, 16, 64, 3, 1),    # two channel blocks (their data gradients: 128 -> 64)
    (2, 64, 30, 32, 5, 2),     # 5x5 filter, partial tiles in both directions
    (3, 40, 32, 24, 3, 1),     # channel counts that are not multiples of the block (zero-filled tails)
    (2, 32, 34, 48, 3, 0),     # no padding (output 32x32; its data gradient, 34x34, stays on the per-tap kernel)
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", HALO_SHAPES)
def test_conv_halo_kernel_bit_exact(cabi, shape, dtype_name):
    """filters larger than 1x1 on large maps go through conv_halo_kernel (one halo tile per channel block, all taps read
    from it through shifted descriptors, filter bank resident): bit for bit the im2col oracle, and identical to the
    per-tap kernel on the same data"""
    B, C, S, N, f, pad = shape
    dtype = cabi.FP16 if dtype_name == "FP16" else cabi.BF16
    L = cabi.lib()
    rng = np.random.default_rng(hash(shape) % 2**31)
    x = _int_tensor(rng, (C, B, S * S), 0.6)
    w = _int_tensor(rng, (N, f * f * C + 1), 0.7)
    w[:, -1] = rng.integers(-2, 3, N)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, 1, pad, bias_value=1.0)
    layer.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = S + 2 * pad - f + 1
    y = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
    assert L.cb200_last_conv_impl() == b"tcgen05-halo"
    ref, col = co.conv_forward(x, w, False, B, C, S, S, f, 1, pad, 1.0)
    assert np.abs(ref).max() < 256
    assert np.array_equal(y, ref), "forward: %d wrong" % int((y != ref).sum())
    dy = _int_tensor(rng, (N, B, So * So), 0.8)
    dyb = cabi.upload_act(dy, dtype, B, N, So, So)
    dx = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
    halo_dgrad = L.cb200_last_conv_impl() == b"tcgen05-halo"
    ref_dx = co.conv_backward_data(dy, w, B, C, S, S, f, 1, pad)
    assert np.abs(ref_dx).max() < 256
    assert np.array_equal(dx, ref_dx), "data gradient: %d wrong" % int((dx != ref_dx).sum())
    assert halo_dgrad or pad == 0 or cabi.round8(N) < 32      # (dy with < 32 channels runs 16-channel blocks: per-tap kernel)
    # optional output path: staging tile + one TMA store per warp (CB200_HALO_TMA_STORE=1; used when N fills the block): same bits
    os.environ["CB200_HALO_TMA_STORE"] = "1"
    try:
        y_tma = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
        assert L.cb200_last_conv_impl() == b"tcgen05-halo"
    finally:
        del os.environ["CB200_HALO_TMA_STORE"]
    assert np.array_equal(y_tma, y)
    # A/B against the per-tap kernel
    L.cb200_force_simt(2)
    try:
        y2 = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
        assert L.cb200_last_conv_impl() == b"tcgen05"
    finally:
        L.cb200_force_simt(0)
    assert np.array_equal(y, y2)
    layer.free(); xb.free(); dyb.free()
