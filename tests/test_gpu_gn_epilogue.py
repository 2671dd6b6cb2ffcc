"""Group-norm statistics out of the convolution epilogue (cb200_conv_forward_stats + cb200_norm_forward_ex /
cb200_norm_pool_forward_ex, csrc/conv_tc.cu: epilogue_loop<GN>) against the statistics pass they replace
(norm_stats_kernel; upstream: cuda_group_mean / cuda_group_var, src/cuda/cuda_norm_layer.cu:65-138).

Both forms sum the SAME numbers - the convolution output as rounded to the 16-bit storage type - so mean / variance agree
to FP32 accumulation order (1e-5 relative on mean, on var and on every normalised value in units of the storage type),
the convolution output itself is bit-identical (contiguous tile runs instead of strided ones move no arithmetic), and the
NumPy oracle holds both to the usual mixed tolerance.  Shapes walk the kernels and lane layouts: per-tap kernel with
1 / 2 / 32 / 64 samples per 128-pixel tile (lane groups of 32 / 4 / 2 pixels per sample), the CTA-pair kernel (two N tiles),
partial last tiles and samples, dead samples, every supported group size, and the refusals (halo kernel, group size 4, tile
rows narrower than the sums of a chunk, CUDA-core kernel) where the flag must stay 0 and nothing changes.
"""
import numpy as np
import pytest

from oracle import cianna_oracle as co
from tests.common import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    from cianna_b200 import cabi as m
    m.init_device(0)
    m.lib().cb200_set_gn_epilogue_stats(2)        # everywhere the arithmetic allows (default: only behind long K loops)
    yield m
    m.lib().cb200_set_gn_epilogue_stats(1)


# (batch, in_c, size, out_c, filter, pad, group size, length, expected flag, expected kernel)
SHAPES = [
    (3, 128, 16, 64, 1, 0, 8, 3, 1, b"tcgen05"),            # per-tap kernel: 16 x 8 tile, one sample per tile
    (3, 32, 24, 64, 3, 1, 8, 3, 1, b"tcgen05"),             # 3x3 on the per-tap kernel (map too ragged for halo tiles), partial tiles
    (5, 64, 8, 128, 1, 0, 16, 4, 1, b"tcgen05"),            # 8 x 8 x 2 tiles: two samples per tile, last tile partial, a dead sample
    (32, 64, 28, 128, 1, 0, 16, 32, 1, b"tcgen05"),         # 4 x 1 x 32 -> lane groups of 4: exactly the 4 sums of gs = 16
    (64, 32, 14, 64, 1, 0, 32, 61, 1, b"tcgen05"),          # 2 x 1 x 64: lane groups of 2, gs = 32
    (4, 256, 7, 1024, 1, 0, 64, 4, 1, b"tcgen05"),          # four N tiles of 256, group of 64 = two chunks
    (2, 64, 12, 136, 1, 0, 8, 2, 1, b"tcgen05"),            # 136 filters: two N tiles, the second almost empty
    (49, 64, 28, 512, 3, 1, 16, 49, 1, b"tcgen05-pair"),    # CTA-pair kernel, two N tiles, odd number of M tiles
    (50, 192, 28, 512, 1, 0, 32, 47, 1, b"tcgen05-pair"),   # CTA-pair kernel, 1x1, dead samples
    (2, 64, 32, 96, 3, 1, 32, 2, 0, b"tcgen05-halo"),       # halo kernel: keeps the statistics pass (see run_igemm)
    (3, 32, 24, 64, 3, 1, 4, 3, 0, b"tcgen05"),             # group size 4: refused
    (64, 32, 14, 64, 1, 0, 8, 64, 0, b"tcgen05"),           # lane groups of 2 < 8 sums per chunk: refused
    (2, 8, 10, 24, 3, 1, 8, 2, 0, b"simt"),                 # CUDA-core kernel: refused
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", SHAPES)
def test_epilogue_statistics_equal_the_statistics_pass(cabi, shape, dtype_name):
    B, C, S, N, f, pad, gs, length, want_flag, want_impl = shape
    dtype = getattr(cabi, dtype_name)
    rng = np.random.default_rng(abs(hash(shape)) % 2**31)
    x = (rng.standard_normal((C, B, S * S)) * 0.8).astype(np.float32)
    w = (rng.standard_normal((N, f * f * C + 1)) * (1.5 / np.sqrt(f * f * C))).astype(np.float32)
    So = S + 2 * pad - f + 1
    act = cabi.activ(cabi.RELU)
    conv = cabi.ConvLayer(dtype, B, C, S, S, N, f, 1, pad, bias_value=0.1, act=act, length=length)
    conv.set_weights(w)
    xb = cabi.upload_act(x, dtype, B, C, S, S)
    G = (N + gs - 1) // gs
    gamma = (1 + 0.2 * rng.standard_normal(G)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(G)).astype(np.float32)

    # the two passes
    n1 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
    n1.set_params(gamma, beta)
    y_conv1 = cabi.download_act(conv.forward(xb), dtype, B, N, So, So)
    y1 = cabi.download_act(n1.forward(conv.y), dtype, B, N, So, So)
    mean1, var1 = n1.stats()[:2]

    # statistics from the epilogue
    n2 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
    n2.set_params(gamma, beta)
    done = conv.forward_stats(xb, n2)
    impl = cabi.lib().cb200_last_conv_impl()
    assert impl == want_impl, impl
    assert done == want_flag
    y_conv2 = cabi.download_act(conv.y, dtype, B, N, So, So)
    assert np.array_equal(y_conv2, y_conv1)
    y2 = cabi.download_act(n2.forward(conv.y, stats_ready=done), dtype, B, N, So, So)
    mean2, var2 = n2.stats()[:2]
    assert rel_err(mean2, mean1) < 1e-5 and rel_err(var2, var1) < 1e-5
    ulp = {"FP16": 2.0 ** -10, "BF16": 2.0 ** -7}[dtype_name]
    assert rel_err(y2, y1) <= ulp

    # oracle on the stored convolution output
    ref_y, ref_mean, ref_var = co.group_norm_forward(y_conv2, gamma, beta, gs, 0, length)
    assert rel_err(mean2[:length], ref_mean[:length]) < 1e-4 and rel_err(var2[:length], ref_var[:length]) < 1e-4
    assert rel_err(y2, ref_y) < 2e-2

    # fused with the max-pool
    if So % 2 == 0:
        outs = []
        for use in (0, 1):
            n3 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
            n3.set_params(gamma, beta)
            p3 = cabi.PoolLayer(dtype, B, N, So, So, 2, 2, 0, cabi.POOL_MAX, length=length)
            flag = conv.forward_stats(xb, n3) if use else (conv.forward(xb), 0)[1]
            outs.append((cabi.download_act(n3.forward_pool(conv.y, p3, stats_ready=flag), dtype, B, N, So // 2, So // 2), p3.map_ref_layout()))
        assert rel_err(outs[1][0], outs[0][0]) <= ulp
        assert (outs[1][1] != outs[0][1]).mean() < 2e-3


# first layer straight from the dataset batch (conv_first.cu): (batch, in_c, size, out_c, filter, pad, group size, length, flag)
FIRST_SHAPES = [
    (3, 3, 32, 32, 3, 1, 4, 3, 1),        # the Darknet19 first layer in small: 32 x 4 tiles of one sample, group size 4
    (3, 3, 40, 32, 3, 1, 8, 2, 1),        # 16 x 8 tiles, partial in both directions (rows outside the map must not count), a dead sample
    (3, 1, 32, 32, 3, 1, 16, 3, 1),       # grey input (KP = 16), groups of 16
    (2, 3, 32, 32, 3, 1, 32, 2, 1),       # one group; tiles of 32 x 2 pixels of TWO samples (the headline's tile is 64 x 1 x 2)
    (4, 3, 64, 32, 3, 1, 4, 3, 1),        # 64 x 1 x 2 tiles, a dead sample sharing tiles with a live one
    (3, 1, 28, 32, 5, 2, 4, 3, 1),        # 5x5 on one channel, partial tiles
    (3, 3, 32, 24, 3, 1, 8, 3, 0),        # 24 filters: rows narrower than the staging tile (plain stores), refused
    (2, 2, 32, 64, 3, 1, 8, 2, 0),        # 64 filters (BN = 64): refused
    (130, 3, 4, 32, 3, 1, 4, 130, 0),     # 16 pixels per sample: a warp's rows span two samples, refused
    (3, 3, 32, 32, 3, 1, 2, 3, 0),        # group size 2 (the sums are kept per 4 columns): refused
]


@pytest.mark.parametrize("dtype_name", ["FP16", "BF16"])
@pytest.mark.parametrize("shape", FIRST_SHAPES)
def test_first_layer_epilogue_statistics(cabi, shape, dtype_name):
    """conv_first_fwd_kernel<GN>: the sums of the first layer's output (the largest statistics pass of Darknet19) from
    its epilogue.  They are taken from the activated FP32 values BEFORE the rounding to 16 bit, the statistics pass reads
    the rounded tensor.  Two checks: (i) tight - inputs and filters are representable in the 16-bit type, so every product
    is exact and the float64 oracle of the UNROUNDED layer output must give the same mean / variance to FP32 accumulation
    order (1e-5); (ii) against the statistics pass they differ by the mean rounding error of a group (~ulp / sqrt(n): 2e-4
    FP16, 2e-3 BF16 for the 4096 values of the smallest group here; 3e-7 for the headline's 800k), the normalised output
    by one unit of the storage type; the convolution output itself is bit-identical."""
    import ctypes
    B, C, S, N, f, pad, gs, length, want_flag = shape
    dtype = getattr(cabi, dtype_name)
    L = cabi.lib()
    rng = np.random.default_rng(abs(hash(shape)) % 2**31)
    def q16(a):      # representable in the storage type: the casts on the way to the device are exact
        a = np.ascontiguousarray(a, np.float32)
        return a.astype(np.float16).astype(np.float32) if dtype_name == "FP16" else (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    x = np.empty((B, C * S * S + 1), np.float32)
    x[:, :-1] = q16(rng.standard_normal((B, C * S * S)) * 0.7)
    x[:, -1] = 0.125
    w = q16(rng.standard_normal((N, f * f * C + 1)) * (1.5 / np.sqrt(f * f * C)))
    conv = cabi.ConvLayer(dtype, B, C, S, S, N, f, 1, pad, bias_value=0.125, act=cabi.activ(cabi.RELU), length=length)
    conv.d.input_is_patches = 2
    assert L.cb200_conv_first_direct(ctypes.byref(conv.d)) == 1
    conv.set_weights(w)
    xt = np.empty(x.size, np.uint16)
    cabi.check(L.cb200_host_cast_from_f32(xt.ctypes.data, dtype, x.ctypes.data, x.size))
    src = cabi.DevBuf.from_numpy(xt)
    So = S + 2 * pad - f + 1
    G = (N + gs - 1) // gs
    gamma = (1 + 0.2 * rng.standard_normal(G)).astype(np.float32)
    beta = (0.1 * rng.standard_normal(G)).astype(np.float32)

    n1 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
    n1.set_params(gamma, beta)
    y_conv1 = cabi.download_act(conv.forward(src), dtype, B, N, So, So)
    y1 = cabi.download_act(n1.forward(conv.y), dtype, B, N, So, So)
    mean1, var1 = n1.stats()[:2]

    n2 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
    n2.set_params(gamma, beta)
    done = conv.forward_stats(src, n2)
    assert L.cb200_last_conv_impl() == b"tcgen05"
    assert done == want_flag
    y_conv2 = cabi.download_act(conv.y, dtype, B, N, So, So)
    assert np.array_equal(y_conv2, y_conv1)
    y2 = cabi.download_act(n2.forward(conv.y, stats_ready=done), dtype, B, N, So, So)
    mean2, var2 = n2.stats()[:2]
    tol = {"FP16": 2e-4, "BF16": 2e-3}[dtype_name] if done else 1e-6
    assert rel_err(mean2, mean1) < tol and rel_err(var2, var1) < tol, (rel_err(mean2, mean1), rel_err(var2, var1))
    if done:
        pre, _ = co.conv_forward(x, w, True, B, C, S, S, f, 1, pad, 0.125)
        _, mean_u, var_u = co.group_norm_forward(co.relu_forward(pre, length), gamma, beta, gs, 0, length)
        e_m, e_v = rel_err(mean2[:length], mean_u[:length]), rel_err(var2[:length], var_u[:length])
        assert e_m < 1e-5 and e_v < 1e-5, (e_m, e_v)
    ulp = {"FP16": 2.0 ** -10, "BF16": 2.0 ** -7}[dtype_name]
    assert rel_err(y2, y1) <= ulp
    ref_y, ref_mean, ref_var = co.group_norm_forward(y_conv2, gamma, beta, gs, 0, length)
    assert rel_err(mean2[:length], ref_mean[:length]) < max(tol, 1e-4) and rel_err(var2[:length], ref_var[:length]) < max(tol, 1e-4)
    assert rel_err(y2, ref_y) < 2e-2
    if length < B:
        assert not y2[:, length:].any()

    # fused with the max-pool, as the network runs it
    outs = []
    for use in (0, 1):
        n3 = cabi.NormLayer(dtype, B, N, So, So, gs, 0, length)
        n3.set_params(gamma, beta)
        p3 = cabi.PoolLayer(dtype, B, N, So, So, 2, 2, 0, cabi.POOL_MAX, length=length)
        flag = conv.forward_stats(src, n3) if use else (conv.forward(src), 0)[1]
        outs.append((cabi.download_act(n3.forward_pool(conv.y, p3, stats_ready=flag), dtype, B, N, So // 2, So // 2), p3.map_ref_layout()))
    assert rel_err(outs[1][0], outs[0][0]) <= ulp
    assert (outs[1][1] != outs[0][1]).mean() < 2e-3


def test_network_uses_the_epilogue_statistics(cabi):
    """host library: a conv -> group-norm (-> max-pool) chain takes its statistics from the epilogue (one statistics
    launch less per pair), results unchanged against the run with the switch off"""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from cianna_b200 import CIANNA as cnn, utils, cabi
from tests import netdefs
spec = netdefs.tc_darknet(batch=8, size=32)
with utils.Quiet():
    utils.build_network(cnn, spec, "C_CUDA", "FP16C_FP32A", network=0)
kinds = [k for k, _ in spec["layers"]]
rng = np.random.default_rng(5)
for i, k in enumerate(kinds):
    if k == "conv":
        w = cnn.layer_weights(i)
        cnn.set_layer_weights(i, (rng.standard_normal(w.shape) * 0.05).astype(np.float32))
x = np.empty((8, 3 * 32 * 32 + 1), np.float32)      # dataset rows: the bias slot comes last
x[:, :-1] = rng.standard_normal((8, 3 * 32 * 32))
x[:, -1] = spec["bias"]
t = np.zeros((8, spec["out_dim"]), np.float32); t[:, 0] = 1
cnn.load_batch(x, t)
L = cabi.lib()
import ctypes
L.cb200_launch_count.restype = ctypes.c_longlong
n0 = L.cb200_launch_count(0)
cnn.forward_batch()
cabi.check(L.cb200_stream_sync(None))
print("LAUNCHES", L.cb200_launch_count(0) - n0)
out = [cnn.layer_output(i) for i in range(len(kinds)) if kinds[i] in ("conv", "pool") or i == len(kinds) - 1]
np.savez(sys.argv[1], *out)
""" % root
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for flag in ("0", "1"):
            path = os.path.join(tmp, "o%s.npz" % flag)
            env = dict(os.environ, CB200_GN_EPILOGUE_STATS="2" if flag == "1" else "0")
            r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-2000:]
            launches = int([ln for ln in r.stdout.splitlines() if ln.startswith("LAUNCHES")][0].split()[1])
            res[flag] = (launches, dict(np.load(path)))
    # tc_darknet at 32 px: the 1x1 layer runs on the per-tap kernel and loses its statistics launch (the first layer and the
    # two 3x3 layers on the 16 px map keep theirs: first-layer / halo kernels)
    assert res["0"][0] - res["1"][0] >= 1, (res["0"][0], res["1"][0])
    for k in res["0"][1]:
        assert rel_err(res["1"][1][k], res["0"][1][k]) < 2e-3, k
