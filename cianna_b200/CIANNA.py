"""Python surface of the B200-native CIANNA core.

Same method names, keyword arguments and defaults as the reference's C extension
(src/python_module.c:1016-1045; kwlists at :48,78,194,301,325,376,420,483,531,585,906,921,941,966),
bound with ctypes onto cianna_b200/libcianna_host.so (host C library) which itself drives
libcianna_b200.so (CUDA core, C-ABI of include/cianna_b200.h).  Usage is unchanged:

    from cianna_b200 import CIANNA as cnn
    cnn.init(in_dim=i_ar([448,448]), in_nb_ch=3, out_dim=1000, bias=0.1, b_size=16,
             comp_meth="C_CUDA", dynamic_load=1, mixed_precision="FP16C_FP32A")
    cnn.conv(f_size=i_ar([3,3]), nb_filters=32, padding=i_ar([1,1]), activation="RELU")
    ...
    cnn.train(nb_iter=1, learning_rate=0.003, momentum=0.9, ...)

There is no CPU path: comp_meth must be "C_CUDA" and a missing CUDA extension / device is an error.
Methods that belong to parts of upstream outside this round's scope (yolo parameters, lrn,
print_arch_tex) raise NotImplementedError rather than silently doing nothing.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST_SO = os.path.join(_HERE, "libcianna_host.so")
_CORE_SO = os.path.join(_HERE, "libcianna_b200.so")
_lib = None
_core = None


def _load():
    global _lib, _core
    if _lib is not None:
        return _lib
    if not (os.path.exists(_HOST_SO) and os.path.exists(_CORE_SO)):
        raise ImportError(
            "cianna_b200 native libraries are missing (%s, %s): run `python -m cianna_b200.build`; "
            "there is no Python/CPU fallback" % (_CORE_SO, _HOST_SO))
    _core = ctypes.CDLL(_CORE_SO, mode=ctypes.RTLD_GLOBAL)
    _lib = ctypes.CDLL(_HOST_SO)
    L = _lib
    vp, ci, cf, cd, cs = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_char_p
    ip, fp = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
    L.init_network.argtypes = [ci, ip, ci, cf, ci, cs, ci, cs, ci, ci, ci]
    L.cb_get_network.restype = vp; L.cb_get_network.argtypes = [ci]
    L.cb_net_layer.restype = vp; L.cb_net_layer.argtypes = [vp, ci]
    L.cb_net_nb_layers.argtypes = [vp]
    L.cb_net_batch_size.argtypes = [vp]
    L.cb_net_last_items_per_s.restype = cf; L.cb_net_last_items_per_s.argtypes = [vp]
    L.cb_net_last_epoch_loss.restype = cd; L.cb_net_last_epoch_loss.argtypes = [vp]
    L.cb_net_set_no_error.argtypes = [vp, ci]
    L.cb_set_dataset.argtypes = [vp, cs, ci, vp, vp]
    L.cb_swap_data_buffers.argtypes = [vp, cs]
    L.cb_net_dataset.restype = vp; L.cb_net_dataset.argtypes = [vp, cs]
    L.free_dataset.argtypes = [vp]
    L.conv_create.argtypes = [vp, vp, ip, ci, ip, ip, ip, ip, cs, fp, cf, cs, cf, vp, ci]
    L.pool_create.argtypes = [vp, vp, ip, ip, ip, cs, cs, ci, cf]
    L.norm_create.argtypes = [vp, vp, cs, cs, ci, ci, vp, ci]
    L.dense_create.argtypes = [vp, vp, ci, cs, fp, cf, ci, cs, cf, vp, ci]
    L.train_network.argtypes = [vp, ci, ci, cf, cf, cf, cf, cf, ci, ci, ci, ci, ci, cf, ci]
    L.forward_testset.argtypes = [vp, ci, ci, ci, ci]
    L.save_network.argtypes = [vp, cs, ci]
    L.load_network.argtypes = [vp, cs, ci, ci, ci]
    L.set_frozen_layers.argtypes = [vp, ip, ci]
    L.perf_eval_display.argtypes = [vp]
    L.cb_load_batch.argtypes = [vp, vp, vp]
    L.cb_forward.argtypes = [vp, ci, ci]
    L.cb_backward.argtypes = [vp, cf, cf, cf]
    L.cb_batch_loss.restype = cf; L.cb_batch_loss.argtypes = [vp]
    L.cb_train_step.argtypes = [vp, cf, cf, cf]
    L.cb_layer_export_output.argtypes = [vp, ci, vp]
    L.cb_layer_export_delta.argtypes = [vp, ci, vp]
    L.cb_layer_export_pool_map.argtypes = [vp, ci, vp]
    L.cb_layer_weight_count.restype = ctypes.c_size_t; L.cb_layer_weight_count.argtypes = [vp, ci]
    L.cb_layer_get_weights.argtypes = [vp, ci, vp]
    L.cb_layer_set_weights.argtypes = [vp, ci, vp]
    L.cb_layer_get_moment.argtypes = [vp, ci, vp]
    L.cb_layer_get_norm_stats.argtypes = [vp, ci, vp, vp, vp, vp]
    L.cb_layer_shape.argtypes = [vp, ci, ip]
    L.cb_dp_unique_id.argtypes = [vp]
    L.cb_dp_init.argtypes = [vp, vp, ci, ci]
    _core.cb200_last_error.restype = cs
    _core.cb200_last_conv_impl.restype = cs
    _core.cb200_version.restype = cs
    _core.cb200_launch_count.restype = ctypes.c_longlong
    _core.cb200_launch_count.argtypes = [ci]
    return _lib


def core():
    """ctypes handle on libcianna_b200.so (the C-ABI), for tests and the benchmark."""
    _load()
    return _core


def host():
    _load()
    return _lib


def _net(network):
    L = _load()
    nid = L.cb_nb_networks() - 1 if network is None else network
    p = L.cb_get_network(nid)
    if not p:
        raise RuntimeError("network %d is not initialised (call init first)" % nid)
    return p


def _i3(arr, default):
    out = (ctypes.c_int * 3)(*default)
    if arr is not None:
        a = np.asarray(arr).astype(np.int64).ravel()
        for i in range(min(3, a.size)):
            out[i] = int(a[i])
    return out


def _s(x):
    return x.encode() if isinstance(x, str) else x


# ------------------------------------------------------------------ reference API
def init(in_dim, in_nb_ch, out_dim, bias=0.1, b_size=8, comp_meth="C_CUDA", network=None, dynamic_load=1,
         mixed_precision="off", inference_only=0, no_logo=0, adv_size=0):
    L = _load()
    dims = (ctypes.c_int * 4)(1, 1, 1, 1)
    a = np.asarray(in_dim).astype(np.int64).ravel()
    for i in range(min(3, a.size)):
        dims[i] = int(a[i])
    dims[3] = int(in_nb_ch)
    nid = L.cb_nb_networks() if network is None else int(network)
    L.init_network(nid, dims, int(out_dim), float(bias), int(b_size), _s(comp_meth), int(dynamic_load),
                   _s(mixed_precision), int(inference_only), int(no_logo), int(adv_size))


def create_dataset(dataset, size, input, target, network=None, silent=0):
    L = _load()
    net = _net(network)
    x = np.ascontiguousarray(input, dtype=np.float32) if input is not None else None
    t = np.ascontiguousarray(target, dtype=np.float32) if target is not None else None
    if not silent:
        print("Setting %s set (size %d)" % (dataset, size))
    L.cb_set_dataset(net, _s(dataset), int(size), x.ctypes.data if x is not None else None,
                     t.ctypes.data if t is not None else None)


def delete_dataset(dataset, network=None, silent=0):
    L = _load()
    L.free_dataset(L.cb_net_dataset(_net(network), _s(dataset)))


def swap_data_buffers(dataset, network=None):
    _load().cb_swap_data_buffers(_net(network), _s(dataset))


def linear():
    return "LIN"


def relu(saturation=float("nan"), leaking=float("nan")):
    s = "RELU"
    if not math.isnan(saturation):
        s += "_S%0.2f" % saturation
    if not math.isnan(leaking):
        s += "_L%0.2f" % leaking
    return s


def logistic(saturation=float("nan"), beta=float("nan")):
    s = "LOGI"
    if not math.isnan(saturation):
        s += "_S%0.2f" % saturation
    if not math.isnan(beta):
        s += "_B%0.2f" % beta
    return s


def softmax():
    return "SMAX"


def yolo():
    return "YOLO"


def _prev(L, net, prev_layer):
    if prev_layer == -1:
        prev_layer = L.cb_net_nb_layers(net) - 1
    return L.cb_net_layer(net, prev_layer) if prev_layer >= 0 else None


def _check_init(init_fct, init_scaling):
    if init_fct in ("normal", "uniform") and init_scaling < 0.0:
        raise SystemExit("ERROR: init_scaling keywork is mandatory when using custom normal or uniform weight initialisation.")


def dense(nb_neurons, activation="RELU", bias=float("nan"), prev_layer=-1, drop_rate=0.0, strict_size=0,
          init_fct="xavier", init_scaling=-1.0, network=None):
    L = _load()
    net = _net(network)
    _check_init(init_fct, init_scaling)
    b = None if math.isnan(bias) else ctypes.pointer(ctypes.c_float(bias))
    return L.dense_create(net, _prev(L, net, prev_layer), int(nb_neurons), _s(activation), b, float(drop_rate),
                          int(strict_size), _s(init_fct), float(init_scaling), None, 0)


def conv(f_size, nb_filters, stride=None, padding=None, int_padding=None, activation="RELU", bias=float("nan"),
         prev_layer=-1, input_shape=None, drop_rate=0.0, init_fct="xavier", init_scaling=-1.0, network=None):
    L = _load()
    net = _net(network)
    _check_init(init_fct, init_scaling)
    b = None if math.isnan(bias) else ctypes.pointer(ctypes.c_float(bias))
    shp = None
    if input_shape is not None:
        shp = (ctypes.c_int * 4)(*[int(v) for v in np.asarray(input_shape).ravel()[:4]])
    return L.conv_create(net, _prev(L, net, prev_layer), _i3(f_size, (1, 1, 1)), int(nb_filters), _i3(stride, (1, 1, 1)),
                         _i3(padding, (0, 0, 0)), _i3(int_padding, (0, 0, 0)), shp, _s(activation), b, float(drop_rate),
                         _s(init_fct), float(init_scaling), None, 0)


def pool(p_size=None, stride=None, padding=None, prev_layer=-1, drop_rate=0.0, p_type="MAX", activation="LIN",
         p_global=0, network=None):
    L = _load()
    net = _net(network)
    # default pool size is 2 along every input dimension larger than 1 (src/python_module.c:497-500)
    in_dims = _net_in_dims(net)
    size = _i3(p_size, tuple(2 if in_dims[i] > 1 else 1 for i in range(3)))
    strd = _i3(stride, tuple(size[i] for i in range(3)))
    return L.pool_create(net, _prev(L, net, prev_layer), size, strd, _i3(padding, (0, 0, 0)), _s(p_type), _s(activation),
                         int(p_global), float(drop_rate))


def _net_in_dims(net):
    # in_dims sits behind the layer table in struct network; ask the library instead of mirroring the struct
    L = _load()
    if not hasattr(L, "_in_dims_ready"):
        L.cb_net_in_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L._in_dims_ready = True
    out = (ctypes.c_int * 4)()
    L.cb_net_in_dims(net, out)
    return list(out)


def norm(normalization="GN", activation="LIN", prev_layer=-1, group_size=8, set_off=0, network=None):
    L = _load()
    net = _net(network)
    return L.norm_create(net, _prev(L, net, prev_layer), _s(normalization), _s(activation), int(group_size), int(set_off), None, 0)


def lrn(*args, **kwargs):
    raise NotImplementedError("lrn layers are not part of this round's hot-path scope (SURVEY.md 8a16)")


def set_frozen_layers(froz_array, network=None):
    L = _load()
    a = np.asarray(froz_array).astype(np.int32).ravel()
    arr = (ctypes.c_int * a.size)(*[int(v) for v in a])
    L.set_frozen_layers(_net(network), arr, int(a.size))


def _yolo_na(*args, **kwargs):
    raise NotImplementedError("YOLO output-layer configuration is not built yet in the B200 core (SURVEY.md 8a21-25)")


set_IoU_limits = set_fit_parts = set_error_scales = set_sm_single = set_slopes_and_maxes = set_yolo_params = _yolo_na


def perf_eval(network=None):
    _load().perf_eval_display(_net(network))


def load(file, iteration, network=None, nb_layers=0, bin=0):
    _load().load_network(_net(network), _s(file), int(iteration), int(nb_layers), int(bin))


def save(file, network=None, bin=0):
    _load().save_network(_net(network), _s(file), int(bin))


def train(nb_iter, learning_rate, end_learning_rate=0.0, control_interv=1, momentum=0.0, lr_decay=0.0, weight_decay=0.0,
          confmat=0, save_every=0, save_bin=0, network=None, shuffle_gpu=1, shuffle_every=1, TC_scale_factor=1.0, silent=0):
    _load().train_network(_net(network), int(nb_iter), int(control_interv), float(learning_rate), float(end_learning_rate),
                          float(momentum), float(lr_decay), float(weight_decay), int(confmat), int(save_every), int(save_bin),
                          int(shuffle_gpu), int(shuffle_every), float(TC_scale_factor), int(silent))


def forward(saving=1, drop_mode="AVG_MODEL", no_error=0, repeat=1, network=None, silent=0):
    L = _load()
    net = _net(network)
    L.cb_net_set_no_error(net, int(no_error))
    L.forward_testset(net, int(saving), int(repeat), 1 if drop_mode == "MC_MODEL" else 0, int(silent))


def print_arch_tex(*args, **kwargs):
    raise NotImplementedError("print_arch_tex (LaTeX export) is outside the hot-path scope (SURVEY.md 8f rank 4)")


# ------------------------------------------------------------------ additions: explicit mini-batch control / read-back
def load_batch(inputs, targets, network=None):
    """inputs: [batch][input_dim+1] FP32 rows in the dataset layout (bias slot last); targets: [batch][out_dim]."""
    x = np.ascontiguousarray(inputs, dtype=np.float32)
    t = np.ascontiguousarray(targets, dtype=np.float32) if targets is not None else None
    _load().cb_load_batch(_net(network), x.ctypes.data, t.ctypes.data if t is not None else None)


def forward_batch(length=None, is_inference=0, network=None):
    L = _load()
    net = _net(network)
    L.cb_forward(net, L.cb_net_batch_size(net) if length is None else int(length), int(is_inference))


def backward_batch(lr, momentum=0.0, weight_decay=0.0, network=None):
    _load().cb_backward(_net(network), float(lr), float(momentum), float(weight_decay))


def batch_loss(network=None):
    return float(_load().cb_batch_loss(_net(network)))


def layer_shape(l, network=None):
    out = (ctypes.c_int * 4)()
    _load().cb_layer_shape(_net(network), int(l), out)
    return tuple(out)


def _act_array(l, network):
    L = _load()
    net = _net(network)
    c, h, w, typ = layer_shape(l, network)
    b = L.cb_net_batch_size(net)
    if typ == 2:  # DENSE: reference layout [B][n+1]
        return np.zeros((b, c + 1), dtype=np.float32)
    return np.zeros((c, b, h * w), dtype=np.float32)


def layer_output(l, network=None):
    """Layer output in the REFERENCE layout: [C][B][H*W] (conv/pool/norm) or [B][n+1] (dense)."""
    a = _act_array(l, network)
    _load().cb_layer_export_output(_net(network), int(l), a.ctypes.data)
    return a


def layer_delta(l, network=None):
    a = _act_array(l, network)
    _load().cb_layer_export_delta(_net(network), int(l), a.ctypes.data)
    return a


def layer_pool_map(l, network=None):
    L = _load()
    net = _net(network)
    c, h, w, _ = layer_shape(l, network)
    a = np.zeros((c, L.cb_net_batch_size(net), h * w), dtype=np.int32)
    L.cb_layer_export_pool_map(net, int(l), a.ctypes.data)
    return a


def layer_weights(l, network=None):
    L = _load()
    net = _net(network)
    a = np.zeros(L.cb_layer_weight_count(net, int(l)), dtype=np.float32)
    if a.size:
        L.cb_layer_get_weights(net, int(l), a.ctypes.data)
    return a


def layer_moment(l, network=None):
    L = _load()
    net = _net(network)
    a = np.zeros(L.cb_layer_weight_count(net, int(l)), dtype=np.float32)
    if a.size:
        L.cb_layer_get_moment(net, int(l), a.ctypes.data)
    return a


def set_layer_weights(l, values, network=None):
    L = _load()
    net = _net(network)
    a = np.ascontiguousarray(values, dtype=np.float32).ravel()
    assert a.size == L.cb_layer_weight_count(net, int(l)), (a.size, L.cb_layer_weight_count(net, int(l)))
    L.cb_layer_set_weights(net, int(l), a.ctypes.data)


def norm_stats(l, nb_group, network=None):
    L = _load()
    net = _net(network)
    b = L.cb_net_batch_size(net)
    arrs = [np.zeros((b, nb_group), dtype=np.float32) for _ in range(4)]
    L.cb_layer_get_norm_stats(net, int(l), *[a.ctypes.data for a in arrs])
    return arrs


def set_TC_scale_factor(value, network=None):
    L = _load()
    L.cb_set_TC_scale_factor.argtypes = [ctypes.c_void_p, ctypes.c_float]
    L.cb_set_TC_scale_factor(_net(network), float(value))


def last_conv_impl():
    return core().cb200_last_conv_impl().decode()


def force_simt(on):
    core().cb200_force_simt(int(on))


def last_perf(network=None):
    L = _load()
    net = _net(network)
    return float(L.cb_net_last_items_per_s(net)), float(L.cb_net_last_epoch_loss(net))


def reset():
    """Forget all networks (device memory of earlier networks is not reclaimed, as upstream)."""
    L = _load()
    ctypes.c_int.in_dll(L, "nb_networks").value = 0
