"""Drives the UNMODIFIED reference CPU back-end (oracle/_ref/<variant>/CIANNA.so) one mini-batch at a
time and reads its tensors back as NumPy arrays, through the reference's own Python API for network
construction and oracle/ref_probe.c for the per-batch step.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py, __graft_entry__.smoke(), bench.py's
reference / cpu_baseline legs).

A network is described by a `spec` dict that is fed unchanged to BOTH sides - the reference module and
cianna_b200.CIANNA expose the same construction API, which is what makes the product a drop-in:

    spec = dict(in_dim=(8, 8), in_ch=3, out_dim=4, bias=0.1, batch=4, layers=[
        ("conv", dict(f_size=(3, 3), nb_filters=8, padding=(1, 1), activation="RELU")),
        ("norm", dict(group_size=4)),
        ("pool", dict(p_size=(2, 2), p_type="MAX")),
        ("conv", dict(f_size=(1, 1), nb_filters=4, activation="LIN")),
        ("pool", dict(p_type="AVG", p_global=1, activation="SMAX")),
    ])
"""
import ctypes
import os
import sys

import numpy as np

from . import ref_loader

CONV, POOL, DENSE, NORM, LRN = 0, 1, 2, 3, 4


def i_ar(v):
    return np.array(v, dtype="int32")


def build_network(cnn, spec, comp_meth, mixed_precision="off", network=None, dynamic_load=1, inference_only=0):
    """Same call sequence on the reference module and on cianna_b200.CIANNA."""
    kw = {} if network is None else {"network": network}
    cnn.init(in_dim=i_ar(spec["in_dim"]), in_nb_ch=spec["in_ch"], out_dim=spec["out_dim"], bias=spec.get("bias", 0.1),
             b_size=spec["batch"], comp_meth=comp_meth, dynamic_load=dynamic_load, mixed_precision=mixed_precision,
             inference_only=inference_only, no_logo=1, **kw)
    if "yolo" in spec:
        # YOLO head set-up goes between init and the layers on both sides (upstream ex. scripts do the same)
        y = dict(spec["yolo"])
        if "prior_size" in y:
            y["prior_size"] = np.ascontiguousarray(y["prior_size"], dtype=np.float32)
        for key in ("prior_noobj_prob", "error_scales", "slopes_and_maxes", "param_ind_scales", "IoU_limits"):
            if key in y:
                y[key] = np.ascontiguousarray(y[key], dtype=np.float32)
        if "fit_parts" in y:
            y["fit_parts"] = np.ascontiguousarray(y["fit_parts"], dtype=np.int32)
        cnn.set_yolo_params(network=0 if network is None else network, **y)
    for kind, a in spec["layers"]:
        a = dict(a)
        for key in ("f_size", "stride", "padding", "int_padding", "p_size"):
            if key in a:
                a[key] = i_ar(a[key])
        a.update(kw)
        if kind == "conv":
            cnn.conv(**a)
        elif kind == "pool":
            cnn.pool(**a)
        elif kind == "norm":
            cnn.norm(**a)
        elif kind == "dense":
            cnn.dense(**a)
        elif kind == "lrn":
            cnn.lrn(**a)
        else:
            raise ValueError(kind)


class _Quiet:
    """silences the C-level stdout of the reference (it prints a lot on every layer creation)"""

    def __enter__(self):
        sys.stdout.flush()
        self._fd = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *a):
        os.dup2(self._fd, 1)
        os.close(self._null)
        os.close(self._fd)


class RefNet:
    def __init__(self, spec, comp_meth="C_BLAS", variant="serial", quiet=True):
        self.cnn, self.lib = ref_loader.load(variant)
        self.spec = spec
        self.B = spec["batch"]
        self.lib.probe_reset()
        if quiet:
            with _Quiet():
                build_network(self.cnn, spec, comp_meth, "off", network=0)
        else:
            build_network(self.cnn, spec, comp_meth, "off", network=0)
        self.n_layers = self.lib.probe_nb_layers(0)
        self.in_dim = int(np.prod(spec["in_dim"])) * spec["in_ch"]      # (w, h) or (w, h, d)
        self._keep = []

    # ---- geometry
    def geom(self, l):
        g = (ctypes.c_int * 16)()
        self.lib.probe_layer_geom(0, l, g)
        return list(g)

    def layer_type(self, l):
        return self.lib.probe_layer_type(0, l)

    def out_shape(self, l):
        """shape of the layer output in the reference layout"""
        t, g = self.layer_type(l), self.geom(l)
        if t == CONV:
            return (g[0], self.B, g[7] * g[8] * g[9])
        if t == POOL:
            return (g[0], self.B, g[7] * g[8] * g[9])
        if t == DENSE:
            return (self.B, g[0] + 1)
        if t in (NORM, LRN):
            return (g[0], self.B, g[4])
        raise ValueError(t)

    def _array(self, l, what, shape, dtype=np.float32):
        p = self.lib.probe_ptr(0, l, what)
        if not p:
            return None
        n = int(np.prod(shape))
        ct = ctypes.c_float if dtype == np.float32 else ctypes.c_int
        buf = (ct * n).from_address(p)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    # ---- tensors (views on the reference's own memory: copy before the next step)
    def output(self, l):
        return self._array(l, 0, self.out_shape(l)).copy()

    def delta(self, l):
        return self._array(l, 1, self.out_shape(l)).copy()

    def weights_view(self, l):
        t, g = self.layer_type(l), self.geom(l)
        if t == CONV:
            return self._array(l, 2, (g[0], g[1]))
        if t == DENSE:
            return self._array(l, 2, (g[1], g[0] + 1))
        return None

    def moment_view(self, l):
        t, g = self.layer_type(l), self.geom(l)
        if t == CONV:
            return self._array(l, 3, (g[0], g[1]))
        if t == DENSE:
            return self._array(l, 3, (g[1], g[0] + 1))
        return None

    def pool_map(self, l):
        return self._array(l, 4, self.out_shape(l), np.int32).copy()

    def dropout_mask(self, l):
        """0/1 mask of the last forward pass that drew one (layout of output(l)); None without dropout"""
        a = self._array(l, 16, self.out_shape(l))
        return None if a is None else a.copy()

    def norm_view(self, l, what):
        g = self.geom(l)
        names = {"gamma": 5, "beta": 6, "mean": 7, "var": 8, "d_gamma": 9, "d_beta": 10, "gamma_update": 11, "beta_update": 12}
        shape = (g[2],) if what in ("gamma", "beta", "gamma_update", "beta_update") else (self.B, g[2])
        return self._array(l, names[what], shape)

    # ---- one mini-batch
    def forward(self, inputs, length=None, is_inference=0):
        x = np.ascontiguousarray(inputs, dtype=np.float32)
        assert x.shape == (self.B, self.in_dim + 1), x.shape
        self._keep = [x]
        self.lib.probe_forward(0, x.ctypes.data_as(ctypes.c_void_p), self.B if length is None else int(length), int(is_inference))

    def backward(self, targets, lr, momentum=0.0, weight_decay=0.0):
        t = np.ascontiguousarray(targets, dtype=np.float32)
        self._keep.append(t)
        self.lib.probe_backward.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float]
        self.lib.probe_backward(0, t.ctypes.data_as(ctypes.c_void_p), lr, momentum, weight_decay)

    def loss(self, targets):
        """per-element loss of the last layer, reference layout of the last layer's output"""
        t = np.ascontiguousarray(targets, dtype=np.float32)
        shape = self.out_shape(self.n_layers - 1)
        out_size = int(np.prod(shape)) // self.B
        err = np.zeros(self.B * out_size, dtype=np.float32)
        self.lib.probe_loss(0, t.ctypes.data_as(ctypes.c_void_p), err.ctypes.data_as(ctypes.c_void_p), out_size)
        return err.reshape(shape)


    # ---- YOLO output layer
    def set_iter(self, it, train_size):
        self.lib.probe_set_iter(0, int(it), int(train_size))

    def _yolo_array(self, what, shape, dtype):
        self.lib.probe_yolo_ptr.restype = ctypes.c_void_p
        p = self.lib.probe_yolo_ptr(0, what)
        ct = ctypes.c_float if dtype == np.float32 else ctypes.c_int
        return np.frombuffer((ct * int(np.prod(shape))).from_address(p), dtype=dtype).reshape(shape).copy()

    def yolo_monitor(self, nb_box):
        """[B][cells][nb_box][2] (objectness, IoU) of the associated boxes after loss(), -1 elsewhere"""
        cells = self.out_shape(self.n_layers - 1)[2]
        return self._yolo_array(0, (self.B, cells, nb_box, 2), np.float32)

    def yolo_box_state(self, nb_box):
        """[B][cells][nb_box] upstream box_locked after the last association pass"""
        cells = self.out_shape(self.n_layers - 1)[2]
        return self._yolo_array(1, (self.B, cells, nb_box), np.int32)

    def set_last_output(self, x):
        l = self.n_layers - 1
        self._array(l, 0, self.out_shape(l))[...] = x

    def last_activation(self, length=None):
        self.lib.probe_last_activation(0, self.B if length is None else int(length))

    def last_deriv_error(self, targets, length=None):
        t = np.ascontiguousarray(targets, dtype=np.float32)
        self._keep.append(t)
        self.lib.probe_last_deriv_error.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        self.lib.probe_last_deriv_error(0, t.ctypes.data_as(ctypes.c_void_p), self.B if length is None else int(length))


def make_yolo_targets(spec, seed, n_obj=None, fill=0.0):
    """seeded target rows [B][1 + max_nb_obj*(7+nb_param+diff_flag)]: boxes inside the image whose sizes scatter around the
    priors, classes in 1..nb_class, params in [0,1), difficult flags in 0..3 when enabled"""
    rng = np.random.default_rng(seed)
    y = spec["yolo"]
    B = spec["batch"]
    W, H = spec["in_dim"]
    nb_param, diff = y.get("nb_param", 0), y.get("diff_flag", 0)
    per = 7 + nb_param + diff
    max_obj = y["max_nb_obj_per_image"]
    prior = np.asarray(y["prior_size"], dtype=np.float32)       # [dims][nb_box]
    t = np.full((B, 1 + max_obj * per), fill, dtype=np.float32)
    for b in range(B):
        n = int(rng.integers(0, max_obj + 1)) if n_obj is None else int(n_obj)
        t[b, 0] = n
        for j in range(n):
            k = int(rng.integers(0, prior.shape[1]))
            w = float(prior[0, k]) * float(np.exp(rng.normal(0, 0.35)))
            h = float(prior[1, k]) * float(np.exp(rng.normal(0, 0.35)))
            w, h = min(w, W - 1.0), min(h, H - 1.0)
            cx = float(rng.uniform(w / 2, W - w / 2))
            cy = float(rng.uniform(h / 2, H - h / 2))
            row = t[b, 1 + j * per: 1 + (j + 1) * per]
            row[0] = rng.integers(1, max(1, y.get("nb_class", 0)) + 1)
            row[1:7] = (cx - w / 2, cy - h / 2, 0.0, cx + w / 2, cy + h / 2, 1.0)
            if nb_param:
                row[7:7 + nb_param] = rng.random(nb_param)
            if diff:
                row[7 + nb_param] = rng.integers(0, 4)
    return t


def make_inputs(spec, seed, scale=1.0):
    """seeded dataset-layout batch [B][C*H*W + 1] (bias slot = spec bias) and one-hot-ish targets"""
    rng = np.random.default_rng(seed)
    B = spec["batch"]
    n = spec["in_dim"][0] * spec["in_dim"][1] * spec["in_ch"]
    x = np.empty((B, n + 1), dtype=np.float32)
    x[:, :n] = (rng.random((B, n), dtype=np.float32) - 0.4) * scale
    x[:, n] = spec.get("bias", 0.1)
    t = np.zeros((B, spec["out_dim"]), dtype=np.float32)
    t[np.arange(B), rng.integers(0, spec["out_dim"], B)] = 1.0
    return x, t
