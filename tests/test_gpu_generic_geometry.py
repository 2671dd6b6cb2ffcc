"""Generic convolution / pooling geometry against the live reference (C_BLAS CPU back-end, oracle/_ref): three-dimensional
filters and maps, 3-D pooling windows, internal padding (transposed convolution), and their combinations with stride
and padding - upstream's im2col address map with all its dimensions (src/cuda/cuda_conv_layer.cu:36-103, int_padding
:66-68) and max_pooling / avg_pooling over z, y, x (src/cuda/cuda_pool_layer.cu:31-277).  These layers run on the
generic CUDA-core kernels (conv_simt.cu, pool3d_* in pool.cu) in every precision mode.
Tolerances: FP32 1e-5 forward / 1e-4 backward and weight change; mixed 2e-2 forward."""
import numpy as np
import pytest

from oracle import ref_driver as rd
from tests.common import ref_available, rel_err

pytestmark = pytest.mark.gpu

HYPER = dict(lr=0.02, momentum=0.9, weight_decay=0.0005)


def net_3d(batch=3):
    return dict(in_dim=(8, 6, 5), in_ch=2, out_dim=4, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3, 3), nb_filters=6, padding=(1, 1, 1), activation="RELU")),
        ("pool", dict(p_size=(2, 2, 1), p_type="MAX")),
        ("conv", dict(f_size=(3, 3, 2), nb_filters=8, padding=(1, 1, 0), activation="RELU")),
        ("norm", dict(normalization="GN", group_size=4)),
        ("pool", dict(p_size=(2, 1, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1, 1), nb_filters=5, activation="LOGI")),
        ("pool", dict(p_size=(1, 3, 2), p_type="AVG")),
        ("dense", dict(nb_neurons=4, strict_size=1, activation="SMAX")),
    ])


def net_3d_strided(batch=2):
    return dict(in_dim=(9, 7, 6), in_ch=1, out_dim=3, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3, 2), nb_filters=4, stride=(2, 2, 2), padding=(0, 0, 0), activation="RELU")),
        ("conv", dict(f_size=(2, 2, 3), nb_filters=6, stride=(1, 1, 1), padding=(1, 0, 1), activation="RELU")),
        ("pool", dict(p_size=(3, 2, 2), stride=(2, 1, 1), padding=(1, 0, 0), p_type="MAX")),
        ("dense", dict(nb_neurons=3, strict_size=1, activation="SMAX")),
    ])


def net_transposed(batch=3):
    """an encoder / decoder pair: 2x2 stride-2 convolution down, internal padding 1 (zero-stuffing) + 3x3 filter up"""
    return dict(in_dim=(8, 8), in_ch=3, out_dim=5, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=6, padding=(1, 1), activation="RELU")),
        ("conv", dict(f_size=(2, 2), nb_filters=8, stride=(2, 2), activation="RELU")),
        ("conv", dict(f_size=(3, 3), nb_filters=6, padding=(1, 1), int_padding=(1, 1), activation="RELU")),      # 4x4 -> 7x7
        # (every dimension divides exactly: upstream warns that its results are "unstable" otherwise, src/conv_layer.c:34-39 -
        #  its backward pass then drops the last input rows)
        ("conv", dict(f_size=(4, 3), nb_filters=4, padding=(2, 1), int_padding=(2, 1), stride=(1, 2), activation="LIN")),     # 7x7 -> 20x7
        ("dense", dict(nb_neurons=5, strict_size=1, activation="SMAX")),
    ])


def net_3d_transposed(batch=2):
    return dict(in_dim=(4, 4, 3), in_ch=2, out_dim=3, bias=0.1, batch=batch, layers=[
        ("conv", dict(f_size=(3, 3, 3), nb_filters=4, padding=(1, 1, 1), int_padding=(1, 1, 1), activation="RELU")),     # 4,4,3 -> 7,7,5
        ("conv", dict(f_size=(3, 3, 3), nb_filters=5, padding=(0, 0, 0), activation="RELU")),
        ("dense", dict(nb_neurons=3, strict_size=1, activation="SMAX")),
    ])


SPECS = {"3d": net_3d, "3d_strided": net_3d_strided, "transposed": net_transposed, "3d_transposed": net_3d_transposed}


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    return m


def _inputs(spec, seed):
    rng = np.random.default_rng(seed)
    n = int(np.prod(spec["in_dim"])) * spec["in_ch"]
    x = np.empty((spec["batch"], n + 1), np.float32)
    x[:, :n] = rng.random((spec["batch"], n), dtype=np.float32) - 0.4
    x[:, n] = spec["bias"]
    t = np.zeros((spec["batch"], spec["out_dim"]), np.float32)
    t[np.arange(spec["batch"]), rng.integers(0, spec["out_dim"], spec["batch"])] = 1
    return x, t


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present on this box")
@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A", "BF16C_FP32A"])
@pytest.mark.parametrize("name", sorted(SPECS))
def test_training_step_matches_live_reference(cnn, name, mode):
    spec = SPECS[name]()
    kinds = [k for k, _ in spec["layers"]]
    ref = rd.RefNet(spec, "C_BLAS")
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)
    # a SEEDED Xavier-normal draw written into the reference (it seeds rand() with the time): the 16-bit comparisons of
    # these tiny networks sit on a handful of ReLU / max-pool decisions, so with its own draw every run would be another case
    w0 = {}
    rng = np.random.default_rng(2025)
    for l, k in enumerate(kinds):
        if k in ("conv", "dense"):
            w = ref.weights_view(l)
            draw = (rng.standard_normal(w.shape) * np.sqrt(2.0 / (w.shape[0] + w.shape[1]))).astype(np.float32)
            if k == "dense":
                w[:, :-1] = draw[:, :-1]      # the last column feeds the next layer's bias node: upstream's own values stay
            else:
                w[...] = draw
            w0[l] = w.copy()
            cnn.set_layer_weights(l, w0[l])
    x, t = _inputs(spec, 21)
    ref.forward(x)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    tol = 1e-5 if mode == "off" else 2e-2
    for l, k in enumerate(kinds):
        a, b = cnn.layer_output(l, network=0), ref.output(l)
        assert a.size == b.size, (l, k, a.shape, b.shape)
        assert rel_err(a, b) < tol, (l, k)
        if k == "pool" and spec["layers"][l][1].get("p_type") == "MAX" and mode == "off":
            assert np.array_equal(cnn.layer_pool_map(l, network=0).ravel(), ref.pool_map(l).ravel()), (l, "argmax map")
    ref.backward(t, **HYPER)
    cnn.backward_batch(HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"], network=0)
    if mode != "off":
        return
    for l, k in reversed(list(enumerate(kinds))):
        assert rel_err(cnn.layer_delta(l, network=0), ref.delta(l)) < 1e-4, (l, k)
        if k in ("conv", "dense"):
            assert rel_err(cnn.layer_weights(l, network=0) - w0[l].ravel(), (ref.weights_view(l) - w0[l]).ravel()) < 1e-4, (l, k)


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_checkpoint_round_trip_of_a_3d_network(cnn, mode, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    spec = net_3d()
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)
    x, t = _inputs(spec, 4)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    last = len(spec["layers"]) - 1
    before = cnn.layer_output(last, network=0)
    for use_bin in (1, 0):
        with rd._Quiet():
            cnn.save("net3d.dat", network=0, bin=use_bin)
            cnn.init(in_dim=rd.i_ar(spec["in_dim"]), in_nb_ch=spec["in_ch"], out_dim=spec["out_dim"], bias=0.1, b_size=spec["batch"],
                     comp_meth="C_CUDA", mixed_precision=mode, no_logo=1, network=0)
            cnn.load("net3d.dat", 0, network=0, bin=use_bin)
        cnn.load_batch(x, t, network=0)
        cnn.forward_batch(network=0)
        # (the text format keeps 6 significant digits: in 16-bit storage that can move an output by one unit in the last place)
        assert rel_err(cnn.layer_output(last, network=0), before) < (1e-6 if use_bin else (1e-4 if mode == "off" else 2e-3))
