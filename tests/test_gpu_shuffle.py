"""train() shuffles the TRAIN set every shuffle_every epochs like upstream (src/auxil.c:1768-1789): a uniform permutation
of whole samples across batches, input and target rows together, host copy and device-resident copy alike."""
import numpy as np
import pytest

from oracle import ref_driver as rd
from tests import netdefs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cnn():
    from cianna_b200 import CIANNA as m
    return m


def _setup(cnn, mode, dynamic_load, n=37):
    spec = netdefs.lenet(batch=8, size=8, d1=16, d2=8)
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0, dynamic_load=dynamic_load)
    dim = 8 * 8
    x = np.repeat(np.arange(n, dtype=np.float32)[:, None], dim, axis=1)      # sample i is filled with the value i
    t = np.zeros((n, 10), np.float32)
    t[np.arange(n), np.arange(n) % 10] = 1
    with rd._Quiet():
        cnn.create_dataset("TRAIN", n, x, t, network=0, silent=1)
    return n, dim


def _ids(cnn, n, dim, device=False):
    xs, ts = cnn.dataset_rows("TRAIN", range(n), network=0, device=device)
    assert np.all(xs[:, :dim] == xs[:, :1])                                   # rows stay whole
    ids = xs[:, 0].astype(int)
    assert np.array_equal(ts.argmax(axis=1), ids % 10)                       # targets travel with their inputs
    return ids


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_shuffle_is_a_permutation_of_whole_samples(cnn, mode):
    n, dim = _setup(cnn, mode, dynamic_load=1)
    assert np.array_equal(_ids(cnn, n, dim), np.arange(n))
    seen = []
    for _ in range(3):
        cnn.shuffle_dataset("TRAIN", network=0)
        ids = _ids(cnn, n, dim)
        assert np.array_equal(np.sort(ids), np.arange(n))
        seen.append(ids)
    assert not np.array_equal(seen[0], np.arange(n)) and not np.array_equal(seen[0], seen[1])
    # samples cross batch boundaries (batch 8): some sample of the first batch came from elsewhere
    assert (seen[0][:8] >= 8).any()


@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_device_shuffle_of_a_resident_set(cnn, mode):
    """train(shuffle_gpu=1) path: cb200_rows_permute scatters whole rows on the device (upstream: shfl_kern +
    get_back_shuffle, src/cuda/cuda_main.cu:590-666); the host copy follows when it is read"""
    n, dim = _setup(cnn, mode, dynamic_load=0)
    cnn.upload_dataset("TRAIN", network=0)
    seen = []
    for _ in range(3):
        cnn.shuffle_dataset("TRAIN", network=0, device=True)
        ids = _ids(cnn, n, dim, device=True)
        assert np.array_equal(np.sort(ids), np.arange(n))
        assert np.array_equal(_ids(cnn, n, dim, device=False), ids)        # host copy refreshed from the device
        seen.append(ids)
    assert not np.array_equal(seen[0], np.arange(n)) and not np.array_equal(seen[0], seen[1])
    assert (seen[0][:8] >= 8).any()
    # a host shuffle afterwards starts from the permuted state and stays coherent with the resident copy
    cnn.shuffle_dataset("TRAIN", network=0)
    assert np.array_equal(_ids(cnn, n, dim, device=True), _ids(cnn, n, dim, device=False))


def test_train_with_shuffle_gpu_permutes_the_resident_set(cnn, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    n, dim = _setup(cnn, "FP16C_FP32A", dynamic_load=0)
    with rd._Quiet():
        cnn.train(nb_iter=2, learning_rate=0.001, shuffle_gpu=1, shuffle_every=1, control_interv=10, silent=1, network=0)
    ids = _ids(cnn, n, dim, device=True)
    assert np.array_equal(np.sort(ids), np.arange(n)) and not np.array_equal(ids, np.arange(n))


@pytest.mark.parametrize("dynamic_load", [1, 0])
def test_train_shuffles_every_shuffle_every_epochs(cnn, dynamic_load, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    n, dim = _setup(cnn, "FP16C_FP32A", dynamic_load)
    with rd._Quiet():
        cnn.train(nb_iter=2, learning_rate=0.001, shuffle_every=0, control_interv=10, silent=1, network=0)
    assert np.array_equal(_ids(cnn, n, dim), np.arange(n))                  # shuffle_every = 0: order untouched
    with rd._Quiet():
        cnn.train(nb_iter=2, learning_rate=0.001, shuffle_every=1, control_interv=10, silent=1, network=0)
    ids = _ids(cnn, n, dim)
    assert np.array_equal(np.sort(ids), np.arange(n)) and not np.array_equal(ids, np.arange(n))
    if not dynamic_load:
        assert np.array_equal(_ids(cnn, n, dim, device=True), ids)           # the device-resident copy follows


@pytest.mark.parametrize("mode,dt", [("FP16C_FP32A", 1), ("BF16C_FP32A", 2)])
def test_large_dataset_is_converted_on_the_device_bit_exact(cnn, mode, dt):
    """create_dataset on >= 2^18 values goes through cb200_dataset_pack (FP32 rows staged on the device, cast there with
    round-toward-zero, bias slot appended): value for value the host conversion loop upstream uses
    (copy_to_FP16 / copy_to_BF16, src/cuda/cuda_main.cu:108-113,790,813), including a partial last batch"""
    import ctypes
    from cianna_b200 import cabi
    spec = netdefs.lenet(batch=8, size=64, d1=16, d2=8)
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_CUDA", mode, network=0, dynamic_load=1)
    n, dim = 67, 64 * 64
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((n, dim)) * np.exp(rng.uniform(-12, 12, (n, dim)))).astype(np.float32)   # normal, subnormal and overflow range of FP16
    t = rng.random((n, 10)).astype(np.float32)
    with rd._Quiet():
        cnn.create_dataset("TRAIN", n, x, t, network=0, silent=1)
    xs, ts = cnn.dataset_rows("TRAIN", range(n), network=0)
    lib = cabi.lib()

    def host_rz(a):
        a = np.ascontiguousarray(a, np.float32)
        out = np.zeros(a.shape, np.uint16)
        assert lib.cb200_host_cast_from_f32(out.ctypes.data_as(ctypes.c_void_p), dt, a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size)) == 0
        return out.view(np.float16).astype(np.float32) if dt == 1 else (out.astype(np.uint32) << 16).view(np.float32)
    assert np.array_equal(xs[:, :dim], host_rz(x))
    assert np.array_equal(xs[:, dim], host_rz(np.full(n, 0.1, np.float32)))
    assert np.array_equal(ts, host_rz(t))
