"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cianna_b200.h declares, refuses to compute without a device (no CPU fallback), and the host
library exports the reference's C API names."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORE = os.path.join(ROOT, "cianna_b200", "libcianna_b200.so")
HOST = os.path.join(ROOT, "cianna_b200", "libcianna_host.so")


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cianna_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_\w+)\s*\(", txt)))


def test_header_declares_a_real_surface():
    syms = declared_symbols()
    assert len(syms) >= 50
    for must in ("cb200_init", "cb200_conv_forward", "cb200_conv_backward_data", "cb200_conv_backward_weights",
                 "cb200_pool_forward", "cb200_norm_backward", "cb200_dp_allreduce"):
        assert must in syms


def test_core_exports_every_declared_symbol():
    lib = ctypes.CDLL(CORE)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert missing == []


def test_host_exports_reference_api():
    ctypes.CDLL(CORE, mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(HOST)
    for name in ("init_network", "create_dataset", "conv_create", "pool_create", "norm_create", "dense_create",
                 "train_network", "forward_testset", "compute_error", "save_network", "load_network", "set_frozen_layers",
                 "conv_save", "conv_load", "pool_save", "norm_save", "dense_save", "nb_area_comp"):
        assert hasattr(lib, name), name


def test_python_surface_matches_reference_method_table():
    from cianna_b200 import CIANNA as cnn
    # the 27 methods of CIANNAMethods, src/python_module.c:1016-1045
    names = ["init", "create_dataset", "delete_dataset", "swap_data_buffers", "linear", "relu", "logistic", "softmax", "yolo",
             "dense", "conv", "pool", "norm", "lrn", "set_frozen_layers", "set_IoU_limits", "set_fit_parts", "set_error_scales",
             "set_sm_single", "set_slopes_and_maxes", "set_yolo_params", "perf_eval", "load", "save", "train", "forward",
             "print_arch_tex"]
    for n in names:
        assert callable(getattr(cnn, n)), n
    assert cnn.relu(saturation=100.0, leaking=0.1) == "RELU_S100.00_L0.10"
    assert cnn.logistic(beta=2.0) == "LOGI_B2.00"


@pytest.mark.skipif(ctypes.CDLL(CORE).cb200_device_count() > 0, reason="a device is present")
def test_no_cpu_fallback_without_device():
    lib = ctypes.CDLL(CORE)
    lib.cb200_last_error.restype = ctypes.c_char_p
    assert lib.cb200_init(0) != 0
    p = ctypes.c_void_p()
    assert lib.cb200_malloc(ctypes.byref(p), 64) == 1          # CB200_ERR_NO_DEVICE
    assert lib.cb200_conv_forward(None, None, None, None, None) == 1
    assert b"no CPU fallback" in lib.cb200_last_error() or b"no CUDA device" in lib.cb200_last_error()


def test_host_cast_round_toward_zero():
    """dataset conversion uses round-toward-zero like upstream (__float2half_rz / __float2bfloat16_rz)"""
    lib = ctypes.CDLL(CORE)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 300, np.array([0.0, -0.0, 1e-6, -1e-6, 65504.0, 1e6, -1e6, 6e-8], np.float32)])
    out = np.empty(x.size, dtype=np.uint16)
    assert lib.cb200_host_cast_from_f32(out.ctypes.data_as(ctypes.c_void_p), 1, x.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size)) == 0
    h = out.view(np.float16).astype(np.float32)
    rn = x.astype(np.float16).astype(np.float32)
    finite = np.abs(x) < 65504
    # RZ result never exceeds |x| and is within one half-ulp step of round-to-nearest
    assert np.all(np.abs(h[finite]) <= np.abs(x[finite]))
    up = np.nextafter(rn.astype(np.float16), np.float16(0)).astype(np.float32)
    assert np.all((h[finite] == rn[finite]) | (h[finite] == up[finite]))
    assert np.all(np.abs(h[~finite]) == 65504.0)
    out_b = np.empty(x.size, dtype=np.uint16)
    assert lib.cb200_host_cast_from_f32(out_b.ctypes.data_as(ctypes.c_void_p), 2, x.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size)) == 0
    assert np.array_equal(out_b, (x.view(np.uint32) >> 16).astype(np.uint16))
