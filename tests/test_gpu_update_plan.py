"""The optimizer sweep in three launches (cianna_b200/csrc/update_plan.cu) against the layer-by-layer launches it
replaces (cb200_conv_update / cb200_norm_reduce_update / cb200_norm_update; upstream: cuda_update_weights +
cuda_master_weight_copy per layer, src/cuda/cuda_conv_layer.cu:559-562, and the host-side gamma / beta update,
src/cuda/cuda_norm_layer.cu:437-459).

C-ABI level: the SAME master weights, momentum buffers and raw gradients are given to both forms; master weights,
momentum, both 16-bit operands, the bias weights and the group-norm parameters must come out IDENTICAL bit for bit
(both forms call the same inline arithmetic, common.cuh: sgd_momentum_step / norm_param_step) in all three precision
types.  (Two whole training runs cannot be compared bit for bit: weight gradients and group-norm statistics are
accumulated with atomics, so two runs of the same code already differ in the last bit.)  The layer-by-layer kernels are
held to the reference fixtures by tests/test_gpu_network.py; the network-level test below checks that a network
trains to the same weights either way, within the FP32 tolerance of those fixtures.
"""
import ctypes

import numpy as np
import pytest

from oracle import ref_driver as rd
from tests import netdefs
from tests.common import HYPER, rel_err

pytestmark = pytest.mark.gpu

# (in_c, size, out_c, filter): 3x3 / 1x1 / 5x5 filters, channel counts off the multiples of 8 and 32, a 2-channel first layer
SHAPES = [(2, 12, 12, 3), (16, 8, 40, 3), (40, 8, 24, 1), (24, 6, 130, 3), (130, 6, 64, 1), (8, 10, 8, 5)]
NORMS = [(4, 12, 4, 0), (5, 40, 8, 1), (3, 130, 16, 0)]        # (batch, channels, group size, set_off)


class NormRef(ctypes.Structure):
    _fields_ = [("d_gamma", ctypes.c_void_p), ("d_beta", ctypes.c_void_p), ("gsum", ctypes.c_void_p),
                ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p), ("gamma_upd", ctypes.c_void_p), ("beta_upd", ctypes.c_void_p),
                ("batch", ctypes.c_int), ("nb_group", ctypes.c_int), ("set_off", ctypes.c_int), ("reduce", ctypes.c_int)]


@pytest.fixture(scope="module")
def cabi():
    from cianna_b200 import cabi as m
    m.init_device(0)
    L = m.lib()
    L.cb200_update_plan_create.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    L.cb200_update_plan_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.cb200_update_plan_destroy.argtypes = [ctypes.c_void_p]
    L.cb200_conv_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    L.cb200_norm_reduce_update.argtypes = [ctypes.c_void_p] * 10
    L.cb200_norm_update.argtypes = [ctypes.c_void_p] * 8
    L.cb200_norm_reduce_grads.argtypes = [ctypes.c_void_p] * 5
    return m


def _upload(cabi, buf, a):
    a = np.ascontiguousarray(a)
    cabi.check(cabi.lib().cb200_h2d(buf.ptr, a.ctypes.data, a.nbytes, None))
    cabi.check(cabi.lib().cb200_stream_sync(None))


def _conv_state(cabi, layer, es):
    L = cabi.lib()
    dp = ctypes.byref(layer.d)
    n_m, n_f, n_b = L.cb200_conv_master_elems(dp), L.cb200_conv_wfwd_elems(dp), L.cb200_conv_wbwd_elems(dp)
    raw = np.uint16 if es == 2 else np.uint32
    return [layer.bufs["master"].to_numpy(np.float32, (n_m,)), layer.bufs["moment"].to_numpy(np.float32, (n_m,)),
            layer.bufs["w_fwd"].to_numpy(raw, (n_f,)), layer.bufs["w_bwd"].to_numpy(raw, (n_b,)),
            layer.bufs["bias_w"].to_numpy(np.float32, (layer.d.out_c,))]


@pytest.mark.parametrize("reduce", [1, 0])
@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
def test_update_plan_is_bit_identical_to_layer_by_layer_updates(cabi, dtype_name, reduce):
    L = cabi.lib()
    dtype = getattr(cabi, dtype_name)
    es = L.cb200_dtype_size(dtype)
    rng = np.random.default_rng(3)
    hyper = cabi.DevBuf.from_numpy(np.array([0.02 / 4, 0.9, 0.0005 * 0.02, 64.0 if dtype_name == "FP16" else 1.0] + [0.0] * 12, np.float32))
    convs, init = [], []
    for in_c, size, out_c, f in SHAPES:
        c = cabi.ConvLayer(dtype, 4, in_c, size, size, out_c, f, pad=f // 2, bias_value=0.1)
        assert L.cb200_update_plan_accepts(ctypes.byref(c.d)) == 1
        dp = ctypes.byref(c.d)
        n_m, n_g = L.cb200_conv_master_elems(dp), L.cb200_conv_grad_elems(dp)
        st = dict(master=(rng.standard_normal(n_m) * 0.2).astype(np.float32), moment=(rng.standard_normal(n_m) * 0.01).astype(np.float32),
                  grad=rng.standard_normal(n_g).astype(np.float32), grad_b=rng.standard_normal(out_c).astype(np.float32))
        convs.append(c)
        init.append(st)
    norms, ninit = [], []
    for batch, ch, gs, set_off in NORMS:
        n = cabi.NormLayer(dtype, batch, ch, 4, 4, gs, set_off)
        G = n.nb_group
        st = dict(d_gamma=rng.standard_normal((batch, G)).astype(np.float32), d_beta=rng.standard_normal((batch, G)).astype(np.float32),
                  gamma=(1 + 0.2 * rng.standard_normal(G)).astype(np.float32), beta=(0.1 * rng.standard_normal(G)).astype(np.float32),
                  gamma_upd=(0.01 * rng.standard_normal(G)).astype(np.float32), beta_upd=(0.01 * rng.standard_normal(G)).astype(np.float32),
                  gsum=rng.standard_normal(2 * G).astype(np.float32))
        n.extra = dict(gamma_upd=cabi.DevBuf(G * 4), beta_upd=cabi.DevBuf(G * 4), gsum=cabi.DevBuf(2 * G * 4))
        norms.append(n)
        ninit.append(st)

    def reset():
        for c, st in zip(convs, init):
            for k, v in st.items():
                _upload(cabi, c.bufs[k], v)
            cabi.check(L.cb200_conv_prepare_weights(ctypes.byref(c.d), ctypes.byref(c.w), None))
        for n, st in zip(norms, ninit):
            _upload(cabi, n.d_gamma, st["d_gamma"]); _upload(cabi, n.d_beta, st["d_beta"])
            _upload(cabi, n.gamma, st["gamma"]); _upload(cabi, n.beta, st["beta"])
            for k in ("gamma_upd", "beta_upd", "gsum"):
                _upload(cabi, n.extra[k], st[k])

    def state():
        out = []
        for c in convs:
            out += _conv_state(cabi, c, es)
        for n in norms:
            G = n.nb_group
            out += [n.gamma.to_numpy(np.float32, (G,)), n.beta.to_numpy(np.float32, (G,)), n.extra["gamma_upd"].to_numpy(np.float32, (G,)),
                    n.extra["beta_upd"].to_numpy(np.float32, (G,)), n.extra["gsum"].to_numpy(np.float32, (2 * G,))]
        return out

    # layer by layer
    reset()
    for c in convs:
        cabi.check(L.cb200_conv_update(ctypes.byref(c.d), ctypes.byref(c.w), hyper.ptr, 0, None))
    for n in norms:
        if reduce:
            cabi.check(L.cb200_norm_reduce_update(ctypes.byref(n.d), n.d_gamma.ptr, n.d_beta.ptr, n.extra["gsum"].ptr, n.gamma.ptr, n.beta.ptr,
                                                  n.extra["gamma_upd"].ptr, n.extra["beta_upd"].ptr, hyper.ptr, None))
        else:       # data-parallel form: the sums arrive all-reduced in gsum
            cabi.check(L.cb200_norm_update(ctypes.byref(n.d), n.gamma.ptr, n.beta.ptr, n.extra["gamma_upd"].ptr, n.extra["beta_upd"].ptr,
                                           n.extra["gsum"].ptr, hyper.ptr, None))
    a = state()
    # the plan
    reset()
    descs = (ctypes.c_void_p * len(convs))(*[ctypes.addressof(c.d) for c in convs])
    ws = (ctypes.c_void_p * len(convs))(*[ctypes.addressof(c.w) for c in convs])
    refs = (NormRef * len(norms))(*[NormRef(n.d_gamma.ptr, n.d_beta.ptr, n.extra["gsum"].ptr, n.gamma.ptr, n.beta.ptr, n.extra["gamma_upd"].ptr,
                                            n.extra["beta_upd"].ptr, n.d.batch, n.nb_group, n.d.set_off, reduce) for n in norms])
    plan = ctypes.c_void_p()
    cabi.check(L.cb200_update_plan_create(ctypes.byref(plan), dtype, descs, ws, len(convs), refs, len(norms)))
    cabi.check(L.cb200_update_plan_run(plan, hyper.ptr, None))
    b = state()
    cabi.check(L.cb200_update_plan_destroy(plan))
    assert len(a) == len(b)
    for i, (u, v) in enumerate(zip(a, b)):
        assert np.array_equal(u, v), (i, dtype_name)
    # and the step did something
    assert not np.array_equal(a[0], init[0]["master"])


def _train(cnn, spec, mode, plan, batches, w0, freeze_at=None):
    cnn.set_update_plan(plan)
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)
    cnn.set_dropout_seed(99, network=0)                # same masks in both runs
    kinds = [k for k, _ in spec["layers"]]
    train = [i for i, k in enumerate(kinds) if k in ("conv", "dense", "norm")]
    if w0:
        for i in train:
            cnn.set_layer_weights(i, w0[i])
    else:
        w0 = {i: cnn.layer_weights(i).copy() for i in train}
    for s, (x, t) in enumerate(batches):
        if freeze_at is not None and s == freeze_at:
            cnn.set_frozen_layers(np.array([train[1], train[2]], dtype="int32"), network=0)
        cnn.load_batch(x, t)
        cnn.forward_batch()
        cnn.backward_batch(**HYPER)
    return w0, {i: cnn.layer_weights(i).copy() for i in train}


@pytest.mark.parametrize("freeze_at", [None, 1])
@pytest.mark.parametrize("net", ["tc_darknet", "dropout_net"])
def test_network_trains_to_the_same_weights_with_and_without_the_plan(net, freeze_at):
    """host side: which layers enter the plan, frozen layers (the plan is rebuilt when the set changes), dense layers and
    the patch-row first layer keeping their own launches - FP32, where two runs agree to accumulation order"""
    from cianna_b200 import CIANNA as cnn
    spec = getattr(netdefs, net)()
    rng = np.random.default_rng(11)
    flat = spec["in_ch"] * int(np.prod(spec["in_dim"]))
    batches = []
    for _ in range(3):
        x = np.empty((spec["batch"], flat + 1), np.float32)      # dataset rows: the bias slot comes last
        x[:, :flat] = rng.standard_normal((spec["batch"], flat))
        x[:, flat] = spec["bias"]
        t = np.zeros((spec["batch"], spec["out_dim"]), np.float32)
        t[np.arange(spec["batch"]), rng.integers(0, spec["out_dim"], spec["batch"])] = 1
        batches.append((x, t))
    try:
        w0, a = _train(cnn, spec, "off", 0, batches, None, freeze_at)
        _, b = _train(cnn, spec, "off", 1, batches, w0, freeze_at)
    finally:
        cnn.set_update_plan(1)
    moved = 0
    for i in a:
        assert rel_err(b[i], a[i]) < 1e-5, (net, i)
        moved += int(not np.array_equal(a[i], w0[i]))
    assert moved >= len(a) - 2 - (2 if freeze_at is not None else 0)
