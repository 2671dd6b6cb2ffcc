"""Drives a CIANNA.so whose compute method is C_CUDA one mini-batch at a time through oracle/ref_probe_cuda.c and the
back-end boundary of the reference (src/prototypes.h:217-295).  Two such libraries exist:

  oracle/_ref/cuda/CIANNA.so           the UNMODIFIED reference incl. its own src/cuda/*.cu + cuBLAS (second oracle:
                                       what upstream's FP32 / FP16 / BF16 GPU path itself computes; parity only, never timed)
  oracle/_ref/dropin/CIANNA.so         the UNMODIFIED reference host code linked against cianna_b200/shim/cuda_b200_shim.c
                                       (the product as a drop-in back-end; built by oracle/build_ref.sh as well)

TEST INFRASTRUCTURE ONLY.  Needs a GPU.  The class mirrors oracle.ref_driver.RefNet (same method names and layouts) so
tests can swap one for the other.
"""
import ctypes
import importlib.util
import os

import numpy as np

from .ref_driver import CONV, DENSE, LRN, NORM, POOL, _Quiet, build_network

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
PATHS = {
    "cuda": os.path.join(_HERE, "_ref", "cuda", "CIANNA.so"),
    "dropin": os.path.join(_HERE, "_ref", "dropin", "CIANNA.so"),
}
_loaded = {}


def available(which="cuda"):
    if not os.path.exists(PATHS[which]):
        return False
    if which == "cuda":
        try:
            ctypes.CDLL("libcublas.so.12")
        except OSError:
            return os.path.exists("/usr/local/cuda/lib64/libcublas.so.12")
    return True


def load(which="cuda"):
    if which in _loaded:
        return _loaded[which]
    path = PATHS[which]
    spec = importlib.util.spec_from_file_location("CIANNA", path)
    cnn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cnn)
    lib = ctypes.CDLL(path)
    _loaded[which] = (cnn, lib)
    return cnn, lib


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class CudaBackendNet:
    def __init__(self, spec, mode="off", which="cuda", quiet=True, inference_only=0, dynamic_load=1):
        self.cnn, self.lib = load(which)
        self.spec, self.B, self.mode = spec, spec["batch"], mode
        self.lib.probe_cuda_reset()
        if quiet:
            with _Quiet():
                build_network(self.cnn, spec, "C_CUDA", mode, network=0, inference_only=inference_only, dynamic_load=dynamic_load)
        else:
            build_network(self.cnn, spec, "C_CUDA", mode, network=0, inference_only=inference_only, dynamic_load=dynamic_load)
        self.n_layers = self.lib.probe_cuda_nb_layers(0)
        self.in_dim = int(np.prod(spec["in_dim"])) * spec["in_ch"]      # (w, h) or (w, h, d)
        self._keep = []

    def geom(self, l):
        g = (ctypes.c_int * 16)()
        self.lib.probe_cuda_layer_geom(0, l, g)
        return list(g)

    def layer_type(self, l):
        return self.lib.probe_cuda_layer_type(0, l)

    def out_shape(self, l):
        t, g = self.layer_type(l), self.geom(l)
        if t in (CONV, POOL):
            return (g[0], self.B, g[7] * g[8] * g[9])
        if t == DENSE:
            return (self.B, g[0] + 1)
        if t in (NORM, LRN):
            return (g[0], self.B, g[4])
        raise ValueError(t)

    def _act(self, l, what):
        shape = self.out_shape(l)
        a = np.empty(int(np.prod(shape)), dtype=np.float32)
        self.lib.probe_cuda_read_act(0, l, what, _fp(a), ctypes.c_size_t(a.size))
        return a.reshape(shape)

    def output(self, l):
        return self._act(l, 0)

    def delta(self, l):
        return self._act(l, 1)

    def _wshape(self, l):
        """(stored shape, shape without upstream's TC padding columns)"""
        t, g = self.layer_type(l), self.geom(l)
        if t == CONV:
            return (g[0], g[1] + g[2]), (g[0], g[1])
        if t == DENSE:
            return (g[1], g[0] + 1), (g[1], g[0] + 1)
        return None, None

    def weights(self, l):
        stored, real = self._wshape(l)
        a = np.empty(int(np.prod(stored)), dtype=np.float32)
        self.lib.probe_cuda_read_f32(0, l, 2, _fp(a), ctypes.c_size_t(a.size))
        return a.reshape(stored)[:, :real[1]].copy()

    def set_weights(self, l, w):
        stored, real = self._wshape(l)
        a = np.zeros(stored, dtype=np.float32)
        a[:, :real[1]] = np.asarray(w, dtype=np.float32).reshape(real)
        self.lib.probe_cuda_write_weights(0, l, _fp(a), ctypes.c_size_t(a.size))

    def moment(self, l):
        stored, real = self._wshape(l)
        a = np.empty(int(np.prod(stored)), dtype=np.float32)
        self.lib.probe_cuda_read_update(0, l, _fp(a), ctypes.c_size_t(a.size))
        return a.reshape(stored)[:, :real[1]].copy()

    def norm(self, l, what):
        g = self.geom(l)
        names = {"gamma": 5, "beta": 6, "mean": 7, "var": 8, "d_gamma": 9, "d_beta": 10}
        shape = (g[2],) if what in ("gamma", "beta") else (self.B, g[2])
        a = np.empty(int(np.prod(shape)), dtype=np.float32)
        self.lib.probe_cuda_read_f32(0, l, names[what], _fp(a), ctypes.c_size_t(a.size))
        return a.reshape(shape)

    def set_norm(self, l, gamma, beta):
        ga, be = np.ascontiguousarray(gamma, np.float32), np.ascontiguousarray(beta, np.float32)
        self.lib.probe_cuda_write_norm(0, l, _fp(ga), _fp(be), int(ga.size))

    def forward(self, inputs, length=None, is_inference=0):
        x = np.ascontiguousarray(inputs, dtype=np.float32)
        assert x.shape == (self.B, self.in_dim + 1), x.shape
        self._keep = [x]
        self.lib.probe_cuda_forward(0, _fp(x), self.B if length is None else int(length), int(is_inference))

    def backward(self, targets, lr, momentum=0.0, weight_decay=0.0, TC_scale=1.0):
        t = np.ascontiguousarray(targets, dtype=np.float32)
        self._keep.append(t)
        self.lib.probe_cuda_backward.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_float] * 4
        self.lib.probe_cuda_backward(0, _fp(t), lr, momentum, weight_decay, TC_scale)

    def loss(self, targets):
        t = np.ascontiguousarray(targets, dtype=np.float32)
        shape = self.out_shape(self.n_layers - 1)
        out_size = int(np.prod(shape)) // self.B
        err = np.zeros(self.B * out_size, dtype=np.float32)
        self.lib.probe_cuda_loss(0, _fp(t), _fp(err), out_size)
        return err.reshape(shape)
