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
