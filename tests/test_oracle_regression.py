"""CPU-only: pins the logistic activation and the LIN / RELU / LOGI output-layer branches of the NumPy restatement
against training steps of the compiled reference (tests/golden/regress_*.npz, tests/golden/make_golden_regression.py)."""
import numpy as np
import pytest

from tests import netdefs
from tests.common import HYPER, load_golden, oracle_from_golden, rel_err

TOL = 2e-5


def per_sample_loss(net, target):
    e = net.loss(target)
    return e.sum(axis=1) if net.layers[-1]["kind"] == "dense" else e.sum(axis=(0, 2))


@pytest.mark.parametrize("case", netdefs.REGRESSION_CASES, ids=lambda c: "%s_%s" % c)
def test_oracle_regression_step_matches_reference(case):
    act, head = case
    g = load_golden("regress_%s_%s" % (act.lower(), head))
    net = oracle_from_golden(netdefs.regression_net(act, head), g)
    length = int(g["length"][0])
    net.forward(g["x"], length)
    for L in net.layers:
        assert rel_err(L["output"], g["out_%d" % L["idx"]]) < TOL, ("output", L["idx"], L["kind"])
    assert rel_err(per_sample_loss(net, g["t"]), g["loss_per_sample"]) < TOL
    assert g["loss_per_sample"][length:].sum() == 0
    net.backward(g["t"], **HYPER)
    for L in net.layers:
        i = L["idx"]
        assert rel_err(L["delta"], g["delta_%d" % i]) < 5 * TOL, ("delta", i, L["kind"])
        if L["kind"] in ("conv", "dense"):
            assert rel_err(L["weights"], g["w1_%d" % i]) < TOL, ("weights", i)
            assert rel_err(L["update"], g["m1_%d" % i]) < 5 * TOL, ("moment", i)


def test_logistic_saturates_the_exponent_argument():
    from oracle import cianna_oracle as co
    x = np.array([-100.0, -6.0, 0.0, 6.0, 100.0], np.float32).reshape(1, 1, 5)
    y = co.logistic_forward(x, 1).ravel()
    assert np.isclose(y[0], 1 / (1 + np.exp(6.0))) and y[0] == y[1]       # clamped: never below 1/(1+e^6)
    assert y[2] == 0.5 and np.isclose(y[4], 1.0)
