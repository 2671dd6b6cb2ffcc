"""Loader for the reference CPU build (oracle/_ref/<variant>/CIANNA.so).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.

`load(variant)` returns (cnn, lib): the reference's own Python module (its public API,
src/python_module.c:1016-1045) and a ctypes handle on the same shared object that also
exposes the reference C API (src/prototypes.h) and oracle/ref_probe.c.
"""
import ctypes
import glob
import importlib.util
import os
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_loaded = {}


def _preload_blas_deps():
    sp = sysconfig.get_paths()["purelib"]
    for pat in ("opencv_python_headless.libs/libgfortran-*.so*", "scipy.libs/libgfortran-*.so*"):
        for p in sorted(glob.glob(os.path.join(sp, pat))):
            try:
                ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                pass


def available(variant="serial"):
    return os.path.exists(os.path.join(_HERE, "_ref", variant, "CIANNA.so"))


def load(variant="serial"):
    """variant: 'serial' (no OpenMP, float abs: parity runs) or 'omp' (upstream flags: timing)."""
    if variant in _loaded:
        return _loaded[variant]
    path = os.path.join(_HERE, "_ref", variant, "CIANNA.so")
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run oracle/build_ref.sh where /root/reference exists)")
    _preload_blas_deps()
    spec = importlib.util.spec_from_file_location("CIANNA", path)
    cnn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cnn)
    lib = ctypes.CDLL(path)  # same object the import mapped: shares networks[] with cnn
    lib.probe_ptr.restype = ctypes.c_void_p
    lib.probe_layer_bias.restype = ctypes.c_float
    _loaded[variant] = (cnn, lib)
    return cnn, lib
