"""perf_eval(): the per-layer timing table of upstream (src/auxil.c:802-870), filled from event-bracketed samples of the
training loop (first mini-batch of the first and of every 16th epoch) instead of a device synchronisation per layer."""
import numpy as np
import pytest

from oracle import ref_driver as rd
from tests import netdefs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    return m


def test_perf_eval_table(cnn, tmp_path, monkeypatch, capfd):
    monkeypatch.chdir(tmp_path)
    spec = netdefs.mini_darknet(batch=8, size=16, classes=6)
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", "FP16C_FP32A", network=0)
    n_layers = len(spec["layers"])
    fwd, back, samples = cnn.perf_eval_table(network=0)
    assert samples == 0
    rng = np.random.default_rng(0)
    x = rng.standard_normal((24, 16 * 16 * 3)).astype(np.float32)
    t = np.zeros((24, 6), np.float32)
    t[np.arange(24), rng.integers(0, 6, 24)] = 1
    with rd._Quiet():
        cnn.create_dataset("TRAIN", 24, x, t, network=0, silent=1)
        cnn.train(nb_iter=3, learning_rate=0.001, control_interv=100, silent=1, network=0)
    fwd, back, samples = cnn.perf_eval_table(network=0)
    assert samples == 1                                   # first epoch only (then every 16th)
    kinds = [k for k, _ in spec["layers"]]
    for i, k in enumerate(kinds):
        assert fwd[i] >= 0 and back[i] >= 0
        if k == "conv":
            assert 1.0 < fwd[i] < 5e4 and 1.0 < back[i] < 5e4, (i, fwd[i], back[i])     # microseconds of real kernels
    with rd._Quiet():
        cnn.train(nb_iter=14, learning_rate=0.001, control_interv=100, silent=1, network=0)    # epochs 4..17 contain the 16th
    assert cnn.perf_eval_table(network=0)[2] == 2
    capfd.readouterr()
    cnn.perf_eval(network=0)
    out = capfd.readouterr().out
    rows = [l for l in out.splitlines() if l.strip() and l.split()[0].isdigit()]
    assert len(rows) == n_layers
    assert [r.split()[1] for r in rows] == [{"conv": "C", "pool": "P", "norm": "N", "dense": "D"}[k] for k in kinds]
    assert "Forward" in out and "Backprop" in out and "Total" in out
