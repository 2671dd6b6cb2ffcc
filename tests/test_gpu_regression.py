"""GPU parity of quadratic-loss networks: logistic hidden layers, LIN / RELU / LOGI output layers (dense and conv heads),
against training steps of the compiled reference (tests/golden/regress_*.npz).  A non-linear output layer multiplies
(o - t) by its own derivative (src/cuda/cuda_activ_functions.cu:2221-2233,2273-2285): cb200_output_delta_activ."""
import numpy as np
import pytest

from oracle import ref_driver as rd
from tests import netdefs
from tests.common import HYPER, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {"off": 1e-5, "FP16C_FP32A": 2e-2, "BF16C_FP32A": 2e-2}


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    return m


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A", "BF16C_FP32A"])
@pytest.mark.parametrize("case", netdefs.REGRESSION_CASES, ids=lambda c: "%s_%s" % c)
def test_regression_step_matches_golden(cnn, case, mode):
    act, head = case
    g = load_golden("regress_%s_%s" % (act.lower(), head))
    spec = netdefs.regression_net(act, head)
    kinds = [k for k, _ in spec["layers"]]
    tol = TOL[mode]
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0)
    S = 64.0 if mode == "FP16C_FP32A" else 1.0
    cnn.set_TC_scale_factor(S, network=0)
    for i, k in enumerate(kinds):
        if k in ("conv", "dense"):
            cnn.set_layer_weights(i, g["w0_%d" % i], network=0)
    length = int(g["length"][0])
    cnn.load_batch(g["x"], g["t"], network=0)
    cnn.forward_batch(length, network=0)
    for i, k in enumerate(kinds):
        assert rel_err(cnn.layer_output(i, network=0), g["out_%d" % i]) < tol, ("output", i, k)
    want_loss = g["loss_per_sample"].sum() / length
    assert abs(cnn.batch_loss(network=0) - want_loss) < max(tol, 1e-4) * max(1.0, abs(want_loss))
    cnn.backward_batch(network=0, **HYPER)
    for i, k in enumerate(kinds):
        assert rel_err(cnn.layer_delta(i, network=0) / S, g["delta_%d" % i]) < tol, ("delta", i, k)
        if k in ("conv", "dense"):
            assert rel_err(cnn.layer_weights(i, network=0), g["w1_%d" % i]) < tol, ("weights", i)
            assert rel_err(cnn.layer_moment(i, network=0) / S, g["m1_%d" % i]) < tol, ("moment", i)
