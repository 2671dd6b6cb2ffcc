"""Filters that cover their whole input map - dense layers behind a conv / pool layer (upstream flattens the maps,
src/cuda/cuda_dense_layer.cu:330-368).  In the channels-last layout they run as ONE 1x1 GEMM over f*f*Cp "channels" in
all three passes (conv_tc.cu: tc_view) with the data-gradient operand kept in the input tensor's own row order (conv.cu:
wbwd_row); the CUDA-core kernels read the same operand.  Integer-valued tensors: bit-exact against the oracle."""
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


# (batch, in_c, map size = filter size, neurons)
WHOLE = [
    (8, 16, 4, 32),
    (130, 12, 8, 40),      # 12 channels padded to 16: the pad lanes are part of the collapsed K axis (zero rows / columns)
    (5, 64, 3, 256),
    (6, 3, 5, 24),         # Cp = 8: 200 collapsed channels, last 64-channel block partial
    (128, 12, 32, 128),    # the first dense layer of the extinction-profile network (12288 inputs), fewer neurons
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16", "FP32"])
@pytest.mark.parametrize("shape", WHOLE)
def test_whole_map_filter_bit_exact(cabi, shape, dtype_name):
    B, C, S, N = shape
    if dtype_name == "FP32" and S == 32:
        pytest.skip("large case on the tensor-core path only")
    dtype = getattr(cabi, dtype_name)
    rng = np.random.default_rng(hash(shape) % 2**31)
    sparse = 0.9 if S == 32 else 0.6          # keep sums of 12k products exactly representable in 16 bits
    x = _int_tensor(rng, (C, B, S * S), sparse)
    w = _int_tensor(rng, (N, S * S * C + 1), 0.9 if S == 32 else 0.7)
    w[:, -1] = rng.integers(-2, 3, N)
    layer = cabi.ConvLayer(dtype, B, C, S, S, N, S, 1, 0, bias_value=1.0)
    layer.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    impl = b"tcgen05" if dtype_name != "FP32" else b"simt"
    y = cabi.download_act(layer.forward(xb), dtype, B, N, 1, 1)
    assert cabi.lib().cb200_last_conv_impl().startswith(impl)      # (tcgen05 / tcgen05-pair)
    ref, col = co.conv_forward(x, w, False, B, C, S, S, S, 1, 0, 1.0)
    assert np.abs(ref).max() < 256
    assert np.array_equal(y, ref)

    dy = _int_tensor(rng, (N, B, 1), 0.8)
    dyb = cabi.upload_act(dy, dtype, B, N, 1, 1)
    dx = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
    assert cabi.lib().cb200_last_conv_impl().startswith(impl)      # (tcgen05 / tcgen05-pair)
    ref_dx = co.conv_backward_data(dy, w, B, C, S, S, S, 1, 0)
    assert np.abs(ref_dx).max() < 256
    assert np.array_equal(dx, ref_dx)

    layer.backward_weights(xb, dyb)
    assert cabi.lib().cb200_last_conv_impl().startswith(impl)      # (tcgen05 / tcgen05-pair)
    assert np.array_equal(layer.grad_ref_layout(), co.conv_weight_grad(col, dy).astype(np.float32))

    if dtype_name != "FP32":
        # the CUDA-core kernels read the same operands (w_bwd in (tap, c) row order)
        cabi.lib().cb200_force_simt(1)
        try:
            y2 = cabi.download_act(layer.forward(xb), dtype, B, N, 1, 1)
            dx2 = cabi.download_act(layer.backward_data(dyb), dtype, B, C, S, S)
            assert cabi.lib().cb200_last_conv_impl() == b"simt"
        finally:
            cabi.lib().cb200_force_simt(0)
        assert np.array_equal(y2, ref) and np.array_equal(dx2, ref_dx)
    layer.free(); xb.free(); dyb.free()
