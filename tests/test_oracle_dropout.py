"""CPU-only: pins the dropout branches of the NumPy restatement (oracle/oracle_net.py: OracleNet._dropout and the
delta masking of backward) against a training step of the compiled reference (tests/golden/dropout_net_blas.npz,
tests/golden/make_golden_dropout.py).  The masks are the reference's own random draws, stored in the fixture."""
import numpy as np

from tests import netdefs
from tests.common import HYPER, load_golden, oracle_from_golden, rel_err

TOL = 2e-5


def _net_with_masks(g):
    net = oracle_from_golden(netdefs.dropout_net(), g)
    for L in net.layers:
        if L["drop"] > 0.01:
            L["mask"] = g["mask_%d" % L["idx"]]
    return net


def test_fixture_masks_are_what_upstream_documents():
    g = load_golden("dropout_net_blas")
    spec = netdefs.dropout_net()
    for i, (kind, a) in enumerate(spec["layers"]):
        rate = a.get("drop_rate", 0.0)
        assert ("mask_%d" % i in g) == (rate > 0.01)
        if rate > 0.01:
            m = g["mask_%d" % i]
            assert set(np.unique(m)) <= {0.0, 1.0}
            if kind == "dense":
                assert np.all(m[:, -1] == 1.0)          # the bias node is never dropped (naiv_dense_layer.c:87)
                m = m[:, :-1]
            assert abs(m.mean() - (1 - rate)) < 0.1
            # dropped pre-activations give act(0) = 0 for RELU / LIN layers
            out = g["out_%d" % i] if kind != "dense" else g["out_%d" % i][:, :-1]
            assert np.all(out[m == 0] == 0)


def test_oracle_training_step_with_given_masks_matches_reference():
    g = load_golden("dropout_net_blas")
    net = _net_with_masks(g)
    net.forward(g["x"])
    for L in net.layers:
        assert rel_err(L["output"], g["out_%d" % L["idx"]]) < TOL, ("output", L["idx"], L["kind"])
        if L["kind"] == "pool":
            assert np.array_equal(L["map"], g["map_%d" % L["idx"]])
    net.backward(g["t"], **HYPER)
    for L in net.layers:
        i = L["idx"]
        assert rel_err(L["delta"], g["delta_%d" % i]) < 5 * TOL, ("delta", i, L["kind"])
        if L["kind"] in ("conv", "dense"):
            assert rel_err(L["weights"], g["w1_%d" % i]) < TOL, ("weights", i)
            assert rel_err(L["update"], g["m1_%d" % i]) < 5 * TOL, ("moment", i)


def test_oracle_inference_scales_instead_of_masking():
    g = load_golden("dropout_net_blas")
    net = _net_with_masks(g)
    net.forward(g["x"], inference=True)
    for L in net.layers:
        assert rel_err(L["output"], g["inf_out_%d" % L["idx"]]) < TOL, ("inference output", L["idx"], L["kind"])
