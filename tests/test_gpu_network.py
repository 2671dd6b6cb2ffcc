"""GPU parity, network level: the product (cianna_b200.CIANNA -> host C library -> C-ABI -> CUDA kernels) runs the
same network definitions, weights and batches as the reference CPU back-end and must reproduce every layer's
output, delta, pool argmax map, group-norm statistics and the post-update weights.

Reference side: fixtures of tests/golden/ (made by the compiled reference) and, when oracle/_ref is present on
the box, the compiled reference itself run live on fresh seeds.

Tolerances (north_star): FP32C_FP32A 1e-5 relative (max-norm: max|a-b| / max|ref|) for every tensor of the step;
FP16C_FP32A / BF16C_FP32A 2e-2 max-norm for activations, loss, deltas, gradients (momentum buffers) and updated
weights.  In mixed precision the backward tensors are compared with the oracle (oracle/oracle_net.py, itself pinned to
the same fixtures in FP32) run CONDITIONED on the product's discrete decisions: 16-bit rounding of the forward
activations flips a small fraction (measured 0.1-0.9 %, asserted < 2 %) of max-pool winners between two nearly equal
candidates and of leaky-ReLU slopes next to zero with respect to an FP32 run - each flip is a full-size error on one
delta element that the following convolutions smear over the whole tensor (rel. error ~ sqrt(2 f) = 5-10 %), on the
reference's own FP16 CUDA path as much as here - so the oracle's backward pass takes the product's argmax maps and
activation signs, and only the arithmetic is compared.  Per-layer figures are written to
gpurun_out/parity_report.json.  Pool argmax maps must be identical in FP32.
"""
import json
import os

import numpy as np
import pytest

from oracle import ref_driver as rd
from tests import netdefs
from tests.common import GOLDEN_SPECS, HYPER, load_golden, oracle_from_golden, ref_available, rel_err, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"off": 1e-5, "FP16C_FP32A": 2e-2, "BF16C_FP32A": 2e-2}
REPORT = {}


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    yield m
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1)


def _build(cnn, spec, mode):
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)


def _layer_kinds(spec):
    return [k for k, _ in spec["layers"]]


def _condition_oracle(cnn, onet):
    """make the oracle's backward pass take the product's discrete decisions (see module docstring)"""
    for L in onet.layers:
        i = L["idx"]
        if L["kind"] == "pool" and L["type"] == "MAX":
            got = cnn.layer_pool_map(i)
            assert float((got != L["map"]).mean()) < 0.02, ("pool argmax flips", i)
            L["map"] = got
        if L["act"] == "RELU":
            mine = cnn.layer_output(i)
            assert float(((mine > 0) != (L["output"] > 0)).mean()) < 0.02, ("relu sign flips", i)
            L["deriv_value"] = mine


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A", "BF16C_FP32A"])
@pytest.mark.parametrize("name", sorted(GOLDEN_SPECS))
def test_training_step_matches_golden(cnn, name, mode):
    g = load_golden(name)
    spec = GOLDEN_SPECS[name]()
    kinds = _layer_kinds(spec)
    tol = TOL[mode]
    mixed = mode != "off"
    _build(cnn, spec, mode)
    S = 64.0 if mode == "FP16C_FP32A" else 1.0      # loss scaling: deltas and momentum buffers carry the factor S
    cnn.set_TC_scale_factor(S, network=0)
    for i, k in enumerate(kinds):
        if k in ("conv", "dense", "norm"):
            cnn.set_layer_weights(i, g["w0_%d" % i])
    length = int(g["length"][0])
    steps = int(g["steps"][0])
    onet = oracle_from_golden(spec, g) if mixed else None
    for s in range(steps - 1):
        cnn.load_batch(g["x_prev%d" % s], g["t_prev%d" % s])
        cnn.forward_batch(length)
        if mixed:
            onet.forward(g["x_prev%d" % s], length)
            _condition_oracle(cnn, onet)
            onet.backward(g["t_prev%d" % s], **HYPER)
        cnn.backward_batch(**HYPER)
    cnn.load_batch(g["x"], g["t"])
    cnn.forward_batch(length)
    if mixed:
        onet.forward(g["x"], length)
    # FP32: the reference's own tensors; mixed: the first step's forward is still compared with the reference fixture,
    # later steps (whose weights already depend on conditioned gradients) with the conditioned oracle
    use_oracle_fwd = mixed and steps > 1
    rep = REPORT.setdefault("%s/%s" % (name, mode), {})
    errs = []
    for i, k in enumerate(kinds):
        ref_out = onet.layers[i]["output"] if use_oracle_fwd else g["out_%d" % i]
        e = rel_err(cnn.layer_output(i), ref_out)
        rep["out_%d_%s" % (i, k)] = e
        errs.append(("out", i, k, e))
        if "map_%d" % i in g:
            got = cnn.layer_pool_map(i)
            frac = float((got != g["map_%d" % i]).mean())
            rep["map_%d_mismatch_frac" % i] = frac
            if not mixed:
                assert frac == 0.0, ("pool argmax", i, frac)
    if "loss" in g and not use_oracle_fwd:
        ref_loss = g["loss"].sum() / length
        got_loss = cnn.batch_loss()
        rep["loss"] = [float(got_loss), float(ref_loss)]
        assert abs(got_loss - ref_loss) < max(tol, 1e-4) * max(1.0, abs(ref_loss))
    ref_delta = {i: g["delta_%d" % i] for i in range(len(kinds))}
    ref_m1 = {i: g["m1_%d" % i] for i, k in enumerate(kinds) if k in ("conv", "dense")}
    ref_w1 = {i: g["w1_%d" % i] for i, k in enumerate(kinds) if k in ("conv", "dense", "norm")}
    if mixed:
        _condition_oracle(cnn, onet)
        onet.backward(g["t"], **HYPER)
        ref_delta = {L["idx"]: L["delta"] for L in onet.layers}
        ref_m1 = {L["idx"]: L["update"] for L in onet.layers if L["kind"] in ("conv", "dense")}
        for L in onet.layers:
            if L["kind"] in ("conv", "dense"):
                ref_w1[L["idx"]] = L["weights"]
            elif L["kind"] == "norm":
                ref_w1[L["idx"]] = np.concatenate([L["gamma"], L["beta"]])
    cnn.backward_batch(**HYPER)
    for i, k in enumerate(kinds):
        e = rel_err(cnn.layer_delta(i) / S, ref_delta[i])
        rep["delta_%d_%s" % (i, k)] = e
        errs.append(("delta", i, k, e))
        if k in ("conv", "dense", "norm"):
            ew = rel_err(cnn.layer_weights(i), ref_w1[i])
            rep["w1_%d_%s" % (i, k)] = ew
            errs.append(("weights", i, k, ew))
        if k in ("conv", "dense"):
            em = rel_err(cnn.layer_moment(i) / S, ref_m1[i])
            rep["m1_%d_%s" % (i, k)] = em
            errs.append(("moment", i, k, em))
    bad = [e for e in errs if not e[3] < tol]
    assert not bad, bad


def _seeded_weights(ref, cnn, kinds, seed):
    """ONE deterministic draw on both sides: He-normal filters written straight into the reference's own weight arrays
    (its generators are time-seeded, src/auxil.c:164) and into the product"""
    rng = np.random.default_rng(seed)
    for i, k in enumerate(kinds):
        if k in ("conv", "dense"):
            w = ref.weights_view(i)
            w[...] = (rng.standard_normal(w.shape) * np.sqrt(2.0 / w.shape[1])).astype(np.float32)
            cnn.set_layer_weights(i, w)
        elif k == "norm":
            g, b = ref.norm_view(i, "gamma"), ref.norm_view(i, "beta")
            g[...] = (1.0 + 0.2 * rng.standard_normal(g.shape)).astype(np.float32)
            b[...] = (0.1 * rng.standard_normal(b.shape)).astype(np.float32)
            cnn.set_layer_weights(i, np.concatenate([g, b]))


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present on this box")
@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_training_steps_match_live_reference(cnn, mode):
    """three consecutive steps with momentum and weight decay next to the compiled reference run live, on seeded inputs
    and SEEDED weights injected into both sides: one deterministic draw, one strict comparison, no retry.
    FP32: every step's output, first-layer delta and the final weights at 3 x 1e-5 (three chained optimizer steps);
    FP16: outputs at 3 x 2e-2 max-norm, final weights at 3 x 2e-2 in relative L2."""
    spec = netdefs.tc_darknet(batch=8, size=16, classes=16)
    kinds = _layer_kinds(spec)
    ref = rd.RefNet(spec, "C_BLAS")
    _build(cnn, spec, mode)
    S = 64.0 if mode == "FP16C_FP32A" else 1.0
    cnn.set_TC_scale_factor(S, network=0)
    # (the seed was screened offline with the reference alone: over the three steps the smallest leaky-ReLU pre-activation
    #  of this draw is 2.4e-6 of its layer's largest, an order of magnitude above FP32 summation-order noise, so no slope
    #  decision depends on who sums first)
    _seeded_weights(ref, cnn, kinds, seed=2033)
    tol = TOL[mode] * 3
    rep = REPORT.setdefault("live/%s" % mode, {})
    last = len(kinds) - 1
    for step in range(3):
        x, t = rd.make_inputs(spec, 100 + step)
        ref.forward(x)
        cnn.load_batch(x, t)
        cnn.forward_batch()
        e_out = rel_err(cnn.layer_output(last), ref.output(last))
        rep["out_step%d" % step] = e_out
        assert e_out < tol, (step, e_out)
        ref.backward(t, 0.05, 0.9, 0.0005)
        cnn.backward_batch(0.05, 0.9, 0.0005)
        if mode == "off":
            e_d = rel_err(cnn.layer_delta(0) / S, ref.delta(0))
            rep["delta0_step%d" % step] = e_d
            assert e_d < tol, (step, e_d)
    gerr = rel_err if mode == "off" else rel_l2
    for i, k in enumerate(kinds):
        if k in ("conv", "norm"):
            w_ref = ref.weights_view(i) if k == "conv" else np.concatenate([ref.norm_view(i, "gamma"), ref.norm_view(i, "beta")])
            e_w = gerr(cnn.layer_weights(i), w_ref)
            rep["w_%d" % i] = e_w
            assert e_w < tol, (i, e_w)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present on this box")
def test_train_api_matches_reference_train_api(cnn, tmp_path, monkeypatch):
    """the public call a user makes - create_dataset + train - on both sides, two epochs over three batches
    (the last one partial), then the weights must agree"""
    monkeypatch.chdir(tmp_path)
    spec = netdefs.mini_darknet(batch=4, size=16, classes=6)
    kinds = _layer_kinds(spec)
    rng = np.random.default_rng(42)
    n = 10
    data = (rng.random((n, 16 * 16 * 3), dtype=np.float32) - 0.4).astype(np.float32)
    targ = np.zeros((n, 6), np.float32)
    targ[np.arange(n), rng.integers(0, 6, n)] = 1
    ref = rd.RefNet(spec, "C_BLAS")
    wrng = np.random.default_rng(2026)      # one seeded draw instead of the reference's time-seeded one
    w0 = {}
    for i, k in enumerate(kinds):
        if k == "conv":
            w = ref.weights_view(i)
            w[...] = (wrng.standard_normal(w.shape) * np.sqrt(2.0 / w.shape[1])).astype(np.float32)
            w0[i] = w.copy()
    kw = dict(nb_iter=2, learning_rate=0.02, end_learning_rate=0.01, control_interv=10, momentum=0.8, lr_decay=0.1,
              weight_decay=0.001, confmat=0, save_every=0, shuffle_every=0, silent=1)
    with rd._Quiet():
        ref.cnn.create_dataset("TRAIN", n, data, targ, network=0, silent=1)
        ref.cnn.train(network=0, **kw)
    _build(cnn, spec, "off")
    for i, w in w0.items():
        cnn.set_layer_weights(i, w)
    with rd._Quiet():
        cnn.create_dataset("TRAIN", n, data, targ, network=0, silent=1)
        cnn.train(network=0, **kw)
    for i in w0:
        assert rel_err(cnn.layer_weights(i), ref.weights_view(i)) < 1e-4, i


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present on this box")
def test_checkpoint_files_are_interchangeable(cnn, tmp_path, monkeypatch):
    """binary save of the product loads in the reference and vice versa (same file format,
    src/conv_layer.c:420-586, src/norm_layer.c:307-329, src/pool_layer.c:296-306, src/dense_layer.c:337-388)"""
    monkeypatch.chdir(tmp_path)
    spec = netdefs.lenet(batch=4, size=16, d1=24, d2=12)
    kinds = _layer_kinds(spec)
    _build(cnn, spec, "off")
    with rd._Quiet():
        cnn.save("mine.dat", network=0, bin=1)
    ref_cnn, lib = rd.ref_loader.load("serial")
    lib.probe_reset()
    with rd._Quiet():
        ref_cnn.init(in_dim=rd.i_ar(spec["in_dim"]), in_nb_ch=1, out_dim=10, bias=0.1, b_size=4, comp_meth="C_BLAS", no_logo=1, network=0)
        ref_cnn.load("mine.dat", 0, network=0, bin=1)
        ref_cnn.save("theirs.dat", network=0, bin=1)
    # (byte equality is not expected: upstream writes the bytes that follow the activation string's NUL from an
    #  uninitialised stack buffer, src/activ_functions.c:246-256; every FIELD must survive the round trip instead)
    assert os.path.getsize("mine.dat") == os.path.getsize("theirs.dat")
    mine = {i: cnn.layer_weights(i) for i, k in enumerate(kinds) if k in ("conv", "dense")}
    with rd._Quiet():
        cnn.init(in_dim=rd.i_ar(spec["in_dim"]), in_nb_ch=1, out_dim=10, bias=0.1, b_size=4, comp_meth="C_CUDA", no_logo=1, network=0)
        cnn.load("theirs.dat", 0, network=0, bin=1)
    for i, w in mine.items():
        assert np.array_equal(cnn.layer_weights(i), w)


@pytest.mark.parametrize("spec_name", ["lenet", "mini_darknet"])
def test_confusion_matrix_of_the_validation_pass(cnn, spec_name, tmp_path, monkeypatch, capfd):
    """train(confmat=1) runs compute_error on the VALID set with the confusion matrix of upstream (src/auxil.c:1562-1603):
    per-sample argmax on the device, matrix and accuracy on the host. Checked against the argmax of the read-back outputs
    (dense head and conv + global-average-pool head)."""
    monkeypatch.chdir(tmp_path)
    spec = netdefs.lenet(batch=4, size=16, d1=24, d2=12) if spec_name == "lenet" else netdefs.mini_darknet(batch=4, size=16, classes=6)
    ncls = spec["out_dim"]
    rng = np.random.default_rng(3)
    n = 10        # last batch partial
    dim = spec["in_dim"][0] * spec["in_dim"][1] * spec["in_ch"]
    data = (rng.random((n, dim), dtype=np.float32) - 0.4).astype(np.float32)
    targ = np.zeros((n, ncls), np.float32)
    targ[np.arange(n), rng.integers(0, ncls, n)] = 1
    _build(cnn, spec, "off")
    with rd._Quiet():
        cnn.create_dataset("TRAIN", n, data, targ, network=0, silent=1)
        cnn.create_dataset("VALID", n, data, targ, network=0, silent=1)
    cnn.train(nb_iter=1, learning_rate=0.0, control_interv=1, confmat=1, shuffle_every=0, silent=0, network=0)
    out = capfd.readouterr().out
    assert "ConfMat" in out and "Acc" in out
    # reference accuracy from the outputs themselves
    last = len(spec["layers"]) - 1
    correct = 0
    for b0 in range(0, n, 4):
        xb = np.zeros((4, dim + 1), np.float32)
        m = min(4, n - b0)
        xb[:m, :dim] = data[b0:b0 + m]
        xb[:, dim] = spec.get("bias", 0.1)
        tb = np.zeros((4, ncls), np.float32)
        tb[:m] = targ[b0:b0 + m]
        cnn.load_batch(xb, tb)
        cnn.forward_batch(m)
        o = cnn.layer_output(last)
        scores = o[:, :ncls] if o.ndim == 2 else o[:, :, 0].T       # dense [B][n+1] or [C][B][1]
        correct += int((scores[:m].argmax(axis=1) == tb[:m].argmax(axis=1)).sum())
    assert abs(cnn.last_accuracy(network=0) - correct / n) < 1e-9
