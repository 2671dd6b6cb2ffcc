"""CPU-only: pins the NumPy restatement (oracle/) against the reference's own output.

  - against the fixtures in tests/golden/ (always; they were produced by the compiled reference), and
  - against the live compiled reference (oracle/_ref) when it is present.
Tolerance: the restatement accumulates in float64 like the NAIV back-end; the BLAS back-end (sgemm, float32
accumulation) differs from it by a few 1e-6 relative, hence 2e-5.
"""
import numpy as np
import pytest

from tests.common import GOLDEN_SPECS, HYPER, load_golden, oracle_from_golden, ref_available, rel_err

TOL = 2e-5


@pytest.mark.parametrize("name", sorted(GOLDEN_SPECS))
def test_oracle_matches_golden(name):
    g = load_golden(name)
    spec = GOLDEN_SPECS[name]()
    net = oracle_from_golden(spec, g)
    length = int(g["length"][0])
    steps = int(g["steps"][0])
    for s in range(steps - 1):
        net.forward(g["x_prev%d" % s], length)
        net.backward(g["t_prev%d" % s], **HYPER)
    net.forward(g["x"], length)
    for L in net.layers:
        i = L["idx"]
        assert rel_err(L["output"], g["out_%d" % i]) < TOL, ("output", i, L["kind"])
        if "map_%d" % i in g:
            assert np.array_equal(L["map"], g["map_%d" % i]), ("pool argmax", i)
        if L["kind"] == "norm":
            assert rel_err(L["mean"], g["mean_%d" % i]) < TOL
            assert rel_err(L["var"], g["var_%d" % i]) < TOL
    if "loss" in g:
        assert rel_err(net.loss(g["t"]), g["loss"]) < TOL
    net.backward(g["t"], **HYPER)
    for L in net.layers:
        i = L["idx"]
        assert rel_err(L["delta"], g["delta_%d" % i]) < 5 * TOL, ("delta", i, L["kind"])
        if L["kind"] in ("conv", "dense"):
            assert rel_err(L["weights"], g["w1_%d" % i]) < TOL, ("weights", i)
            assert rel_err(L["update"], g["m1_%d" % i]) < 5 * TOL, ("moment", i)
        if L["kind"] == "norm":
            got = np.concatenate([L["gamma"], L["beta"]])
            assert rel_err(got, g["w1_%d" % i]) < TOL
            assert rel_err(L["d_gamma"], g["dgamma_%d" % i]) < 5 * TOL
            assert rel_err(L["d_beta"], g["dbeta_%d" % i]) < 5 * TOL


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("comp", ["C_NAIV", "C_BLAS"])
def test_oracle_matches_live_reference(comp):
    from oracle import ref_driver as rd
    from oracle.oracle_net import OracleNet
    from tests import netdefs
    spec = netdefs.mini_darknet(batch=3, size=12, classes=5)
    ref = rd.RefNet(spec, comp)
    net = OracleNet(spec)
    for L in net.layers:
        if L["kind"] == "conv":
            L["weights"] = ref.weights_view(L["idx"]).copy()
    x, t = rd.make_inputs(spec, 3)
    ref.forward(x, 2)
    net.forward(x, 2)
    for L in net.layers:
        assert rel_err(L["output"], ref.output(L["idx"])) < TOL, L["idx"]
    ref.backward(t, 0.05, 0.5, 0.001)
    net.backward(t, 0.05, 0.5, 0.001)
    for L in net.layers:
        assert rel_err(L["delta"], ref.delta(L["idx"])) < 5 * TOL, L["idx"]
        if L["kind"] == "conv":
            assert rel_err(L["weights"], ref.weights_view(L["idx"])) < TOL


def test_golden_fixture_integrity():
    """edge cases present in the fixtures: a tail batch (length < batch) zeroes the extra samples"""
    g = load_golden("mini_darknet_naiv_tail")
    assert int(g["length"][0]) == 3
    assert np.all(g["out_0"][:, 3:, :] == 0)          # RELU conv output of the padded sample
    assert np.all(g["delta_%d" % 11][:, 3:, :] == 0)


# ---- YOLO output head: the restatement (oracle/yolo_oracle.py) against fixtures made by the compiled reference
from oracle import yolo_oracle as yo  # noqa: E402
from tests import netdefs  # noqa: E402


def _yolo_case(name):
    spec = netdefs.yolo_head(name)
    g = dict(np.load("%s/golden/yolo_%s.npz" % (__file__.rsplit("/", 1)[0], name)))
    grid = (spec["in_dim"][0] // spec["layers"][0][1]["stride"][0], spec["in_dim"][1] // spec["layers"][0][1]["stride"][1])
    return spec, g, yo.YoloSetup(spec["yolo"], spec["in_dim"], grid)


@pytest.mark.parametrize("name", netdefs.YOLO_HEAD_CASES)
def test_yolo_oracle_matches_golden(name):
    spec, g, s = _yolo_case(name)
    assert rel_err(yo.activation(s, g["x"]), g["a"]) < 1e-6
    for sfx in ("", "_h"):
        a = g["a"] if sfx == "" else g["a"].astype(np.float16).astype(np.float32)
        t = g["t"] if sfx == "" else _rz16(g["t"])
        d, st = yo.run(s, a, t, "delta")
        assert np.array_equal(st, g["state" + sfx]), "box association differs"
        assert rel_err(d, g["delta" + sfx]) < 1e-6
        e, mon = yo.run(s, a, t, "loss")
        assert np.array_equal(mon[..., 0] > -0.98, g["monitor" + sfx][..., 0] > -0.98)
        assert rel_err(e, g["loss" + sfx]) < 1e-6
        assert rel_err(mon, g["monitor" + sfx]) < 1e-6


def _rz16(a):
    a = np.asarray(a, dtype=np.float32)
    h = a.astype(np.float16)
    over = np.abs(h.astype(np.float32)) > np.abs(a)
    return np.where(over, np.nextafter(h, np.float16(0)), h).astype(np.float32)


def test_yolo_fixtures_exercise_every_association_branch():
    """the fixtures are only worth something if they reach the branches: associated / good-but-not-best / background boxes,
    strict prior sets, low-IoU re-association, difficult targets that are skipped, a class-only image, crowded cells"""
    seen = {}
    for name in netdefs.YOLO_HEAD_CASES:
        spec, g, s = _yolo_case(name)
        st = g["state"]
        seen[name] = [int((st == v).sum()) for v in (0, 1, 2)]
        assert seen[name][0] > 0 and seen[name][2] > 0, (name, seen[name])
    assert sum(v[1] for v in seen.values()) > 5
    spec, g, s = _yolo_case("giou_default")
    assert g["t"][0, 0] == -1.0
    spec, g, s = _yolo_case("single_box")
    assert (g["state"] == 2).sum() < g["t"][:, 0].sum()          # more targets than boxes somewhere
    spec, g, s = _yolo_case("diou2_difficult")
    per = 7 + 1 + 1
    nobj = g["t"][:, 0].astype(int)
    flags = np.concatenate([g["t"][b, 1: 1 + nobj[b] * per].reshape(-1, per)[:, -1] for b in range(len(nobj))])
    assert (flags > 0).any() and (g["state"] == 2).sum() < nobj.sum()


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not present on this box")
@pytest.mark.parametrize("name", ["giou_default", "iou_strict", "custom_tables"])
def test_yolo_oracle_matches_live_reference(name):
    """fresh seeds against the compiled reference, incl. its end-to-end forward of the raw values"""
    from oracle import ref_driver as rd
    spec, _, s = _yolo_case(name)
    ref = rd.RefNet(spec, "C_BLAS")
    ref.set_iter(1, spec["batch"])
    last = ref.n_layers - 1
    for seed in (7, 8):
        rng = np.random.default_rng(seed)
        x = (1.5 * rng.standard_normal(ref.out_shape(last))).astype(np.float32)
        t = rd.make_yolo_targets(spec, seed + 50)
        ref.set_last_output(x)
        ref.last_activation()
        a = ref.output(last)
        assert rel_err(yo.activation(s, x), a) < 1e-6
        ref.last_deriv_error(t)
        d, st = yo.run(s, a, t, "delta")
        assert np.array_equal(st, ref.yolo_box_state(s.nb_box))
        assert rel_err(d, ref.delta(last)) < 1e-6
        ref.set_last_output(a)
        e, mon = yo.run(s, a, t, "loss")
        assert rel_err(e, ref.loss(t)) < 1e-6
        assert rel_err(mon, ref.yolo_monitor(s.nb_box)) < 1e-6
