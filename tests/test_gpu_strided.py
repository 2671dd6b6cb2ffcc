"""Strided convolutions on the tcgen05 implicit-GEMM kernels (stride 2: what upstream's SKA SDC1 detector uses in place of
pooling, examples/SKAO_SDC1/train_network.py:118-132).  Forward and weight gradient read the input through a TMA box with
a traversal stride; the data gradient of a filter that tiles its input (f == stride, no padding) is one 1x1 GEMM per tap
scattered with that stride; other strided data gradients stay on the CUDA-core kernels.  Integer-valued tensors: bit-exact
against the oracle's im2col + GEMM."""
import numpy as np
import pytest

from oracle import cianna_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from cianna_b200 import cabi as m
    m.init_device(0)
    return m


def _int_tensor(rng, shape, p_zero=0.5):
    v = rng.choice(np.array([-1.0, 0.0, 1.0], np.float32), size=shape, p=[(1 - p_zero) / 2, p_zero, (1 - p_zero) / 2])
    return v.astype(np.float32)


# (batch, in_c, size, out_c, f, stride, pad, data gradient on tensor cores?)
STRIDED = [
    (4, 32, 16, 16, 2, 2, 0, True),      # SDC1 layer 2 shape: 2x2 stride 2, 32 -> 16 channels
    (3, 64, 14, 128, 2, 2, 0, True),     # odd batch, partial tiles
    (2, 128, 28, 64, 2, 2, 0, True),     # two channel blocks
    (5, 16, 10, 24, 2, 2, 0, True),      # 16-channel operands, 24 filters
    (2, 384, 8, 512, 2, 2, 0, True),     # SDC1 layer 13 shape (wide)
    (4, 32, 15, 32, 3, 2, 1, False),     # overlapping 3x3 stride 2: forward / weight gradient strided, data gradient on CUDA cores
    (2, 32, 15, 32, 2, 2, 0, False),     # input not covered by the filters (15 = 2*7 + 1): last row / column gets no gradient
    # channel counts between the operand slab widths (SDC1 layers 3-4: 16 -> 24 -> 32 channels): the weight gradient reads a
    # 32- / 64-channel slab from a 24- / 40-channel tensor (TMA zero-fills the rest)
    (4, 24, 12, 32, 3, 1, 1, True),
    (3, 40, 9, 48, 3, 1, 1, True),
    (2, 24, 16, 64, 2, 2, 0, True),
    # thin 3x3 layers on large maps (>= 4 x 148 pixel tiles): weight gradient with swapped operands (conv_wgrad_swap_kernel:
    # M = (filter row, input channel), N = filters)
    (8, 32, 72, 64, 3, 1, 1, True),      # Darknet19 layer 2 channel counts
    (6, 24, 90, 40, 3, 1, 1, True),      # 24 / 40 channels: slabs wider than the tensors
    (5, 16, 96, 16, 3, 1, 0, True),      # no padding, 16 channels both sides
    (4, 64, 100, 128, 3, 1, 1, True),    # Darknet19 layers 3 / 5 channel counts: two tap groups (5 + 4 taps), 64-channel slabs
    (3, 40, 120, 96, 3, 1, 1, True),     # 40 / 96 channels in the same class
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", STRIDED)
def test_strided_conv_bit_exact(cabi, shape, dtype_name):
    B, C, S, N, f, stride, pad, tc_dgrad = shape
    dtype = cabi.FP16 if dtype_name == "FP16" else cabi.BF16
    rng = np.random.default_rng(hash(shape) % 2**31)
    x = _int_tensor(rng, (C, B, S * S), 0.6)
    w = _int_tensor(rng, (N, f * f * C + 1), 0.7)
    w[:, -1] = rng.integers(-2, 3, N)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, f, stride, pad, bias_value=1.0)
    layer.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    So = (S + 2 * pad - f) // stride + 1
    y = cabi.download_act(layer.forward(xb), dtype, B, N, So, So)
    assert cabi.lib().cb200_last_conv_impl().startswith(b"tcgen05")          # (the halo variant on large stride-1 maps)
    ref, col = co.conv_forward(x, w, False, B, C, S, S, f, stride, pad, 1.0)
    assert np.abs(ref).max() < 256
    assert np.array_equal(y, ref)

    dy = _int_tensor(rng, (N, B, So * So), 0.8)
    dyb = cabi.upload_act(dy, dtype, B, N, So, So)
    dx = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
    assert cabi.lib().cb200_last_conv_impl().startswith(b"tcgen05") == tc_dgrad
    ref_dx = co.conv_backward_data(dy, w, B, C, S, S, f, stride, pad)
    assert np.abs(ref_dx).max() < 256
    assert np.array_equal(dx, ref_dx)
    # the derivative hook of the layer in front, on the strided scatter as well
    prev = cabi.upload_act(x, dtype, B, C, S, S)
    dxh = cabi.download_act(layer.backward_data(dyb, prev_act=cabi.activ(cabi.RELU), prev_out=prev), dtype, B, C, S, S)
    want = co.relu_deriv(ref_dx, x, B)
    assert np.abs(dxh - want).max() <= np.abs(want).max() * (2.0 ** -8)          # x 0.05 is not exact in 16 bits

    layer.backward_weights(xb, dyb)
    got = layer.grad_ref_layout()
    assert cabi.lib().cb200_last_conv_impl() == b"tcgen05"
    assert np.array_equal(got, co.conv_weight_grad(col, dy).astype(np.float32))
    layer.free(); xb.free(); dyb.free(); prev.free()
