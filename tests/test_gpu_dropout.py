"""GPU parity of dropout (conv / pool / dense layers).

The reference draws its masks at random (cuRAND / rand()), so parity is stated on everything that FOLLOWS from a mask:
the product reports the mask it used (it is a function of seed, layer, pass counter and position - nothing is stored),
the checker (oracle/oracle_net.py, pinned to a reference-made training step with dropout by tests/test_oracle_dropout.py)
takes the same mask, and outputs, deltas, momentum buffers and updated weights must agree to the usual tolerances; the
AVG_MODEL inference pass has no randomness and is compared with the reference fixture directly.  The masks themselves are
checked for what upstream specifies: 0/1, kept with probability 1 - rate, independent between layers / passes / seeds,
identical between a forward pass and its backward pass.
"""
import numpy as np
import pytest

from oracle import cianna_oracle as co
from oracle import ref_driver as rd
from tests import netdefs
from tests.common import HYPER, load_golden, oracle_from_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {"off": 1e-5, "FP16C_FP32A": 2e-2}


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


@pytest.mark.parametrize("dtype_name", ["FP32", "FP16", "BF16"])
@pytest.mark.parametrize("cfg", [(4, 16, 9, 0.25, "RELU", 3), (3, 13, 6, 0.5, "LIN", 3), (2, 24, 1, 0.2, "LOGI", 2)])
def test_dropout_forward_backward_follow_the_mask(cabi, cfg, dtype_name):
    B, C, S, rate, act_name, length = cfg
    dtype = getattr(cabi, dtype_name)
    act = {"RELU": cabi.activ(cabi.RELU), "LIN": cabi.activ(cabi.LINEAR), "LOGI": cabi.activ(cabi.LOGISTIC)}[act_name]
    rng = np.random.default_rng(5)
    x = _representable((rng.standard_normal((C, B, S * S)) * 2).astype(np.float32), dtype_name)
    dy = _representable(rng.standard_normal((C, B, S * S)).astype(np.float32), dtype_name)
    drop = cabi.Dropout(dtype, B, C, S, S, rate, act, seed=77, stream_id=3, draw=9, length=length)
    mask = drop.mask()
    assert set(np.unique(mask)) <= {0.0, 1.0}

    def activate(z):
        if act_name == "RELU":
            return co.relu_forward(z, length)
        if act_name == "LOGI":
            out = (1.0 / (1.0 + np.exp(np.minimum(-z.astype(np.float64), 6.0)))).astype(np.float32)
            out[:, length:, :] = 0
            return out
        return z

    tol = 1e-6 if dtype_name == "FP32" else 1e-2            # one rounding to the storage type (BF16: 2^-8)
    y = cabi.download_act(drop.forward(cabi.upload_act(x, dtype, B, C, S, S)), dtype, B, C, S, S)
    assert rel_err(y, activate(x * mask)) < tol
    y = cabi.download_act(drop.forward(cabi.upload_act(x, dtype, B, C, S, S), scale_only=True), dtype, B, C, S, S)
    assert rel_err(y, activate(x * np.float32(1.0 - rate))) < tol
    d = cabi.download_act(drop.backward(cabi.upload_act(dy, dtype, B, C, S, S)), dtype, B, C, S, S)
    assert np.array_equal(d, dy * mask)                     # the backward pass sees exactly the forward mask


def test_dropout_mask_statistics_and_streams(cabi):
    B, C, S = 8, 64, 32
    n = B * C * S * S
    for rate in (0.1, 0.25, 0.5, 0.8):
        m = cabi.Dropout(cabi.FP16, B, C, S, S, rate, seed=123, stream_id=1, draw=1).mask()
        kept = m.mean()
        sigma = np.sqrt(rate * (1 - rate) / n)
        assert abs(kept - (1 - rate)) < 5 * sigma + 1.0 / 65536, (rate, kept)
        # no structure along any axis of the tensor: per-channel / per-sample / per-pixel keep rates are all close to 1 - rate
        for axis in ((1, 2), (0, 2), (0, 1)):
            part = m.mean(axis=axis)
            assert np.abs(part - (1 - rate)).max() < 6 * np.sqrt(rate * (1 - rate) / (n / part.size))
    base = cabi.Dropout(cabi.FP32, B, C, S, S, 0.5, seed=123, stream_id=1, draw=1).mask()
    assert np.array_equal(base, cabi.Dropout(cabi.FP32, B, C, S, S, 0.5, seed=123, stream_id=1, draw=1).mask())
    # the mask does not depend on the storage type
    assert np.array_equal(base, cabi.Dropout(cabi.BF16, B, C, S, S, 0.5, seed=123, stream_id=1, draw=1).mask())
    for other in (dict(seed=124, stream_id=1, draw=1), dict(seed=123, stream_id=2, draw=1), dict(seed=123, stream_id=1, draw=2)):
        m = cabi.Dropout(cabi.FP32, B, C, S, S, 0.5, **other).mask()
        agree = (m == base).mean()
        assert abs(agree - 0.5) < 0.01, (other, agree)      # independent draws agree half of the time


def _build(cnn, spec, mode):
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_training_step_with_dropout_matches_oracle_on_the_same_masks(cnn, mode):
    g = load_golden("dropout_net_blas")
    spec = netdefs.dropout_net()
    kinds = [k for k, _ in spec["layers"]]
    tol = TOL[mode]
    _build(cnn, spec, mode)
    cnn.set_dropout_seed(2024, network=0)
    for i, k in enumerate(kinds):
        if k in ("conv", "dense"):
            cnn.set_layer_weights(i, g["w0_%d" % i], network=0)
    # inference (AVG_MODEL): no randomness -> the reference's own tensors
    cnn.load_batch(g["x"], g["t"], network=0)
    cnn.forward_batch(is_inference=1, network=0)
    for i, k in enumerate(kinds):
        assert rel_err(cnn.layer_output(i, network=0), g["inf_out_%d" % i]) < tol, ("inference output", i, k)
    # training step
    cnn.forward_batch(network=0)
    onet = oracle_from_golden(spec, g)
    for L in onet.layers:
        if L["drop"] > 0.01:
            m = cnn.layer_dropout_mask(L["idx"], network=0)
            assert set(np.unique(m)) <= {0.0, 1.0}
            if L["kind"] == "dense":
                m[:, -1] = 1.0                          # the bias node (not part of the product's tensor) is never dropped
            core = m[:, :-1] if L["kind"] == "dense" else m
            assert abs(core.mean() - (1 - L["drop"])) < 0.2, (L["idx"], core.mean())
            L["mask"] = m
    onet.forward(g["x"])
    for L in onet.layers:
        i = L["idx"]
        got = cnn.layer_output(i, network=0)
        if L["kind"] == "dense" and L["act"] != "SMAX":
            got, want = got[:, :-1], L["output"][:, :-1]
        else:
            want = L["output"]
        assert rel_err(got, want) < tol, ("output", i, L["kind"])
        if L["drop"] > 0.01 and L["kind"] != "dense":
            assert np.all(got[L["mask"] == 0] == 0)
    if mode != "off":
        return                                           # mixed precision: forward only (decision flips, see test_gpu_network)
    cnn.backward_batch(network=0, **HYPER)
    onet.backward(g["t"], **HYPER)
    for L in onet.layers:
        i = L["idx"]
        got, want = cnn.layer_delta(i, network=0), L["delta"]
        if L["kind"] == "dense":
            got, want = got[:, :-1], want[:, :-1]
        assert rel_err(got, want) < 5 * tol, ("delta", i, L["kind"])
        if L["drop"] > 0.01:
            mk = L["mask"][:, :-1] if L["kind"] == "dense" else L["mask"]
            assert np.all(got[mk == 0] == 0)
        if L["kind"] in ("conv", "dense"):
            assert rel_err(cnn.layer_weights(i, network=0), L["weights"]) < tol, ("weights", i)
            assert rel_err(cnn.layer_moment(i, network=0), L["update"]) < 5 * tol, ("moment", i)


def test_masks_change_every_pass_and_inference_modes(cnn):
    spec = netdefs.dropout_net()
    _build(cnn, spec, "off")
    cnn.set_dropout_seed(7, network=0)
    x, t = rd.make_inputs(spec, 4)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    m1, o1 = cnn.layer_dropout_mask(0, network=0), cnn.layer_output(6, network=0)
    cnn.forward_batch(network=0)
    m2, o2 = cnn.layer_dropout_mask(0, network=0), cnn.layer_output(6, network=0)
    assert 0.2 < (m1 != m2).mean() < 0.6 and not np.array_equal(o1, o2)
    # AVG_MODEL inference is deterministic, MC_MODEL inference draws masks like training
    cnn.forward_batch(is_inference=1, network=0)
    a1 = cnn.layer_output(6, network=0)
    cnn.forward_batch(is_inference=1, network=0)
    assert np.array_equal(a1, cnn.layer_output(6, network=0))
    cnn.set_inference_drop_mode("MC_MODEL", network=0)
    cnn.forward_batch(is_inference=1, network=0)
    c1 = cnn.layer_output(6, network=0)
    cnn.forward_batch(is_inference=1, network=0)
    assert not np.array_equal(c1, cnn.layer_output(6, network=0))
    cnn.set_inference_drop_mode("AVG_MODEL", network=0)
    # same seed, same pass counter -> same run
    _build(cnn, spec, "off")
    cnn.set_dropout_seed(7, network=0)
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(network=0)
    assert np.array_equal(cnn.layer_dropout_mask(0, network=0), m1)


def test_mnist_example_network_with_its_dropout_trains(cnn):
    """examples/MNIST/mnist_train.py:67-73 upstream as written (drop_rate 0.5 / 0.2 on the dense layers)"""
    spec = dict(in_dim=(28, 28), in_ch=1, out_dim=10, bias=0.1, batch=16, layers=[
        ("conv", dict(f_size=(5, 5), nb_filters=8, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(5, 5), nb_filters=16, padding=(2, 2), activation="RELU")),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("dense", dict(nb_neurons=256, activation="RELU", drop_rate=0.5)),
        ("dense", dict(nb_neurons=128, activation="RELU", drop_rate=0.2)),
        ("dense", dict(nb_neurons=10, strict_size=1, activation="SMAX")),
    ])
    _build(cnn, spec, "FP16C_FP32A")
    cnn.set_dropout_seed(1, network=0)
    rng = np.random.default_rng(0)
    # 10 class prototypes + noise: learnable in a few dozen steps
    protos = rng.standard_normal((10, 28 * 28)).astype(np.float32)
    def batch():
        lab = rng.integers(0, 10, 16)
        x = np.zeros((16, 28 * 28 + 1), np.float32)
        x[:, :-1] = protos[lab] + 0.5 * rng.standard_normal((16, 28 * 28)).astype(np.float32)
        t = np.zeros((16, 10), np.float32)
        t[np.arange(16), lab] = 1
        return x, t, lab
    losses = []
    for _ in range(60):
        x, t, _ = batch()
        cnn.load_batch(x, t, network=0)
        cnn.forward_batch(network=0)
        losses.append(cnn.batch_loss(network=0))
        cnn.backward_batch(0.02, 0.8, network=0)
    assert np.mean(losses[-10:]) < 0.5 * np.mean(losses[:5]), (losses[:5], losses[-10:])
    x, t, lab = batch()
    cnn.load_batch(x, t, network=0)
    cnn.forward_batch(is_inference=1, network=0)
    pred = cnn.layer_output(6, network=0)[:, :10].argmax(axis=1)
    assert (pred == lab).mean() > 0.8


@pytest.mark.skipif(not rd.ref_loader.available("serial"), reason="oracle/_ref not present on this box")
def test_forward_repeat_mc_dropout_file_matches_reference_layout(cnn, tmp_path, monkeypatch):
    """forward(repeat=N): every batch is forwarded N times from the first dropout layer on and every pass is written to
    fwd_res (src/auxil.c:1216-1226, 1346-1400).  AVG_MODEL has no randomness: the product's file must equal the one the
    compiled reference writes for the same network, weights and TEST set (10 samples, batch 6: a partial last batch), value
    for value and record for record.  MC_MODEL: same record layout, passes differ from each other, the part below the
    first dropout layer is not re-drawn (layer 0 has dropout here, so every pass is a full pass), and the mean over many
    passes approaches the AVG_MODEL prediction."""
    spec = netdefs.dropout_net(batch=6, size=12)
    kinds = [k for k, _ in spec["layers"]]
    n, rep = 10, 3
    rng = np.random.default_rng(21)
    dim = 12 * 12 * 2
    data = (rng.random((n, dim), dtype=np.float32) - 0.4).astype(np.float32)
    targ = np.zeros((n, 5), np.float32)
    targ[np.arange(n), rng.integers(0, 5, n)] = 1
    (tmp_path / "ref").mkdir()
    (tmp_path / "mine").mkdir()
    monkeypatch.chdir(tmp_path / "ref")
    ref = rd.RefNet(spec, "C_BLAS")
    w0 = {i: ref.weights_view(i).copy() for i, k in enumerate(kinds) if k in ("conv", "dense")}
    with rd._Quiet():
        ref.cnn.create_dataset("TEST", n, data, targ, network=0, silent=1)
        ref.cnn.forward(saving=2, drop_mode="AVG_MODEL", repeat=rep, network=0, silent=1)
    theirs = np.fromfile(tmp_path / "ref" / "fwd_res" / "net0_0000.dat", dtype=np.float32)
    monkeypatch.chdir(tmp_path / "mine")
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", "off", network=0)
    for i, w in w0.items():
        cnn.set_layer_weights(i, w)
    cnn.set_dropout_seed(5, network=0)
    with rd._Quiet():
        cnn.create_dataset("TEST", n, data, targ, network=0, silent=1)
        cnn.forward(saving=2, drop_mode="AVG_MODEL", repeat=rep, network=0, silent=1)
    mine = np.fromfile(tmp_path / "mine" / "fwd_res" / "net0_0000.dat", dtype=np.float32)
    assert mine.size == theirs.size == n * rep * 6     # nb_neurons + the bias node, as upstream writes dense outputs
    assert rel_err(mine, theirs) < 1e-5
    # record layout: batch after batch, `rep` consecutive blocks of the batch's samples
    assert np.all(mine.reshape(-1, 6)[:, 5] == 0.0)
    blocks = [mine[:6 * rep * 6].reshape(rep, 6, 6)[:, :, :5], mine[6 * rep * 6:].reshape(rep, 4, 6)[:, :, :5]]
    for blk in blocks:
        for r in range(1, rep):
            assert np.array_equal(blk[r], blk[0])
    avg = np.concatenate([blocks[0][0], blocks[1][0]])
    loss_avg = cnn.last_perf(network=0)[1]
    ref_loss = float(-(targ * np.log(np.maximum(avg, 1e-6))).sum() / n)
    assert abs(loss_avg - ref_loss) < 1e-4 * max(1.0, ref_loss)        # mean over samples AND repeats
    # MC_MODEL
    rep_mc = 64
    with rd._Quiet():
        cnn.forward(saving=2, drop_mode="MC_MODEL", repeat=rep_mc, network=0, silent=1)
    mc = np.fromfile(tmp_path / "mine" / "fwd_res" / "net0_0000.dat", dtype=np.float32)
    assert mc.size == n * rep_mc * 6
    b0 = mc[:6 * rep_mc * 6].reshape(rep_mc, 6, 6)[:, :, :5]
    assert not np.array_equal(b0[0], b0[1])
    assert np.allclose(b0.sum(axis=2), 1.0, atol=1e-4)
    assert np.abs(b0.mean(axis=0) - blocks[0][0]).max() < 0.15
    # training afterwards goes back to AVG_MODEL for its validation pass (src/auxil.c:1793)
    with rd._Quiet():
        cnn.create_dataset("TRAIN", n, data, targ, network=0, silent=1)
        cnn.create_dataset("VALID", n, data, targ, network=0, silent=1)
        cnn.train(nb_iter=1, learning_rate=0.0, control_interv=1, shuffle_every=0, silent=1, network=0)
    a = cnn.last_perf(network=0)[1]
    with rd._Quiet():
        cnn.train(nb_iter=1, learning_rate=0.0, control_interv=1, shuffle_every=0, silent=1, network=0)
    assert cnn.last_perf(network=0)[1] == a        # no masks drawn: two validation passes agree exactly
