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
Options of upstream that the B200 core does not cover yet ("not covered") stop with an explicit
error rather than silently doing nothing.
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
    L.lrn_create.argtypes = [vp, vp, cs, ci, cf, cf, cf, vp, ci]
    L.print_architecture_tex.argtypes = [vp, cs, cs] + [ci] * 11
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


def shuffle_dataset(dataset="TRAIN", network=None, device=False):
    """the permutation train() applies every shuffle_every epochs, on demand; device=True: the on-device variant of
    train(shuffle_gpu=1) for a resident set (dynamic_load=0)"""
    L = _load()
    fn = L.shuffle_dataset_device if device else L.shuffle_dataset
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    fn(_net(network), L.cb_net_dataset(_net(network), _s(dataset)))


def upload_dataset(dataset="TRAIN", network=None):
    """make the set device-resident now (train() does it itself when dynamic_load=0)"""
    L = _load()
    L.dataset_upload.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.dataset_upload(_net(network), L.cb_net_dataset(_net(network), _s(dataset)))


def dataset_rows(dataset, indices, network=None, device=False):
    """(inputs [n][input_dim+1], targets [n][output_dim]) of the given samples as stored (FP32 view of the storage type)"""
    L = _load()
    net = _net(network)
    L.cb_dataset_read_row.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.cb_net_io_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    dims = (ctypes.c_longlong * 3)()
    L.cb_net_io_dims(net, dims)
    in_dim, out_dim, dtype = int(dims[0]), int(dims[1]), int(dims[2])
    np_t = np.float32 if dtype == 0 else np.uint16
    data = L.cb_net_dataset(net, _s(dataset))
    xs = np.zeros((len(indices), in_dim + 1), np_t)
    ts = np.zeros((len(indices), max(out_dim, 1)), np_t)
    for k, i in enumerate(indices):
        L.cb_dataset_read_row(net, data, int(i), 0, int(device), xs[k].ctypes.data)
        if out_dim:
            L.cb_dataset_read_row(net, data, int(i), 1, int(device), ts[k].ctypes.data)

    def widen(a):
        if dtype == 0:
            return a
        if dtype == 1:
            return a.view(np.float16).astype(np.float32)
        return (a.astype(np.uint32) << 16).view(np.float32)
    return widen(xs), widen(ts[:, :out_dim])


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


def lrn(activation="LIN", prev_layer=-1, range=5, k=1.0, alpha=1.0, beta=0.5, network=None):
    """Local response normalisation across channels; keywords and defaults of src/python_module.c:551-575."""
    L = _load()
    net = _net(network)
    return L.lrn_create(net, _prev(L, net, prev_layer), _s(activation), int(range), float(k), float(alpha), float(beta), None, 0)


def set_frozen_layers(froz_array, network=None):
    L = _load()
    a = np.asarray(froz_array).astype(np.int32).ravel()
    arr = (ctypes.c_int * a.size)(*[int(v) for v in a])
    L.set_frozen_layers(_net(network), arr, int(a.size))


# ---- YOLO output layer configuration (src/python_module.c:583-884): the five small helpers only build arrays whose
# "unset" entries carry the markers set_yolo_params looks for
def set_IoU_limits(good_IoU_lim=-2.0, low_IoU_best_box_assoc=-2.0, min_prob_IoU_lim=-2.0, min_obj_IoU_lim=-2.0,
                   min_class_IoU_lim=-2.0, min_param_IoU_lim=-2.0, diff_IoU_lim=-2.0, diff_obj_lim=-2.0):
    return np.array([good_IoU_lim, low_IoU_best_box_assoc, min_prob_IoU_lim, min_obj_IoU_lim, min_class_IoU_lim,
                     min_param_IoU_lim, diff_IoU_lim, diff_obj_lim], dtype=np.float32)


def set_fit_parts(position=-2, size=-2, probability=-2, objectness=-2, classes=-2, parameters=-2):
    return np.array([position, size, probability, objectness, classes, parameters], dtype=np.int32)


def set_error_scales(position=-1.0, size=-1.0, probability=-1.0, objectness=-1.0, classes=-1.0, parameters=-1.0):
    return np.array([position, size, probability, objectness, classes, parameters], dtype=np.float32)


def set_sm_single(slope=-1.0, fmin=-100000.0, fmax=100000.0):
    return np.array([slope, fmax, fmin], dtype=np.float32)


def set_slopes_and_maxes(position=None, size=None, probability=None, objectness=None, classes=None, parameters=None):
    out = np.empty((6, 3), dtype=np.float32)
    for i, v in enumerate((position, size, probability, objectness, classes, parameters)):
        out[i] = (-1.0, 100000.0, -100000.0) if v is None else np.asarray(v, dtype=np.float32).ravel()[:3]
    return out


def _fptr(a, n=None):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    if n is not None and a.size != n:
        raise SystemExit("ERROR: YOLO parameter array has %d elements, expected %d" % (a.size, n))
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def set_yolo_params(nb_box=0, nb_class=0, nb_param=0, max_nb_obj_per_image=0, prior_size=None, prior_noobj_prob=None,
                    error_scales=None, slopes_and_maxes=None, param_ind_scales=None, IoU_limits=None, fit_parts=None,
                    IoU_type="empty", prior_dist_type="empty", strict_box_size=0, fit_dim=0, rand_startup=-1,
                    rand_prob_best_box_assoc=0.0, rand_prob=0.0, min_prior_forced_scaling=0.0, class_softmax=0, diff_flag=0,
                    network=0, error_type="empty", no_override=0, raw_output=0):
    """Same keywords and defaults as upstream (src/python_module.c:764-884); returns the number of filters the YOLO
    layer must have. prior_size is [dims][nb_box] like upstream."""
    L = _load()
    net = _net(network)
    keep = []
    c_prior = None
    if prior_size is not None:
        ps = np.asarray(prior_size, dtype=np.float32)
        if ps.ndim != 2 or ps.shape[1] != nb_box:
            raise SystemExit("ERROR: The prior_size array must have nb_box elements!")
        if fit_dim <= 0:
            fit_dim = ps.shape[0]
        elif fit_dim > ps.shape[0]:
            raise SystemExit("ERROR: fit_dim parameter cannot be larger than the number of dimensions of prior_size!")
        flat = np.zeros((nb_box, 3), dtype=np.float32)
        flat[:, :fit_dim] = ps[:fit_dim].T
        keep.append(flat)
        c_prior = flat.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    a_noobj, c_noobj = _fptr(prior_noobj_prob, nb_box)
    a_scales, c_scales = _fptr(error_scales, 6)
    a_pis, c_pis = _fptr(param_ind_scales, nb_param)
    a_lim, c_lim = _fptr(IoU_limits, 8)
    c_sm = None
    if slopes_and_maxes is not None:
        sm = np.ascontiguousarray(slopes_and_maxes, dtype=np.float32).reshape(6, 3)
        rows = (ctypes.POINTER(ctypes.c_float) * 6)(*[sm[i].ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for i in range(6)])
        keep += [sm, rows]
        c_sm = rows
    c_fit = None
    if fit_parts is not None:
        fp = np.ascontiguousarray(fit_parts, dtype=np.int32).ravel()
        keep.append(fp)
        c_fit = fp.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    L.set_yolo_params.restype = ctypes.c_int
    L.set_yolo_params.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                  ctypes.c_char_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.POINTER(ctypes.c_float)),
                                  ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int),
                                  ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    return L.set_yolo_params(net, int(nb_box), int(nb_class), int(nb_param), int(max_nb_obj_per_image), _s(IoU_type),
                             _s(prior_dist_type), c_prior, c_noobj, int(fit_dim), int(strict_box_size), int(rand_startup),
                             float(rand_prob_best_box_assoc), float(rand_prob), float(min_prior_forced_scaling), c_scales, c_sm,
                             c_pis, c_lim, c_fit, int(class_softmax), int(diff_flag), _s(error_type), int(no_override),
                             int(raw_output))


def perf_eval(network=None):
    _load().perf_eval_display(_net(network))


def perf_eval_table(network=None):
    """(forward_us[nb_layers], backprop_us[nb_layers], number of sampled mini-batches) behind perf_eval()'s table"""
    L = _load()
    net = _net(network)
    n = L.cb_net_nb_layers(net)
    fwd, back = np.zeros(n, np.float64), np.zeros(n, np.float64)
    L.cb_perf_eval_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    samples = L.cb_perf_eval_read(net, fwd.ctypes.data, back.ctypes.data)
    return fwd, back, int(samples)


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


def print_arch_tex(path, file_name, size=1, in_size=1, f_size=1, out_size=1, stride=1, padding=1, in_padding=0,
                   activation=0, bias=0, dropout=0, param_count=0, network=None):
    """Architecture table as path/file_name.tex (+ .pdf when pdflatex is installed); keywords and defaults of
    src/python_module.c:991-1009."""
    _load().print_architecture_tex(_net(network), _s(path), _s(file_name), int(size), int(in_size), int(f_size),
                                   int(out_size), int(stride), int(padding), int(in_padding), int(activation),
                                   int(bias), int(dropout), int(param_count))


# ------------------------------------------------------------------ additions: explicit mini-batch control / read-back
def load_batch(inputs, targets, network=None):
    """inputs: [batch][input_dim+1] FP32 rows in the dataset layout (bias slot last); targets: [batch][out_dim]."""
    x = np.ascontiguousarray(inputs, dtype=np.float32)
    t = np.ascontiguousarray(targets, dtype=np.float32) if targets is not None else None
    L = _load()
    net = _net(network)
    dims = (ctypes.c_longlong * 3)()
    L.cb_net_io_dims.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    L.cb_net_io_dims(net, dims)
    b = L.cb_net_batch_size(net)
    if x.size != b * (dims[0] + 1):
        raise ValueError("load_batch: inputs hold %d values, the network takes %d rows of input_dim + 1 = %d (bias slot last)"
                         % (x.size, b, dims[0] + 1))
    if t is not None and dims[1] > 0 and t.size != b * dims[1]:
        raise ValueError("load_batch: targets hold %d values, the network takes %d rows of %d" % (t.size, b, dims[1]))
    L.cb_load_batch(net, x.ctypes.data, t.ctypes.data if t is not None else None)


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


def set_dropout_seed(seed, network=None):
    """Dropout masks are a function of (seed, layer, forward-pass counter, position); the default seed is time based."""
    L = _load()
    L.cb_set_dropout_seed.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
    L.cb_set_dropout_seed(_net(network), int(seed))


def set_inference_drop_mode(drop_mode="AVG_MODEL", network=None):
    """for forward_batch(is_inference=1); forward() takes drop_mode itself like upstream"""
    L = _load()
    L.cb_set_inference_drop_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.cb_set_inference_drop_mode(_net(network), 1 if drop_mode == "MC_MODEL" else 0)


def layer_dropout_mask(l, network=None):
    """0/1 mask layer l used in its last forward pass, same layout as layer_output (a dense layer's bias node reads 0)."""
    L = _load()
    L.cb_layer_export_dropout_mask.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    a = _act_array(l, network)
    L.cb_layer_export_dropout_mask(_net(network), int(l), a.ctypes.data)
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


# ---- YOLO read-backs (parity tests): association state, loss split, decoded boxes
def _yolo_geom(network):
    L = _load()
    net = _net(network)
    last = L.cb_net_nb_layers(net) - 1
    c, h, w, _ = layer_shape(last, network)
    return L, net, L.cb_net_batch_size(net), c, h * w


def yolo_set_seed(seed, network=None):
    L = _load()
    L.cb_yolo_set_seed.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong]
    L.cb_yolo_set_seed(_net(network), int(seed))


def set_iter(iteration, train_size=0, network=None):
    """epoch counter of the network (drives the YOLO random start-up phase: iter * train.size <= rand_startup);
    train_size stands in for the TRAIN dataset's size when batches are fed with load_batch"""
    L = _load()
    L.cb_net_set_iter.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.cb_net_set_iter(_net(network), int(iteration), int(train_size))


def yolo_box_state(nb_box, network=None):
    """[B][cells][nb_box] int32 after backward_batch: 0 background, 1 good-but-not-best, 2 associated to a target"""
    L, net, b, c, cells = _yolo_geom(network)
    a = np.zeros((b, cells, nb_box), dtype=np.int32)
    L.cb_yolo_box_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.cb_yolo_box_state(net, a.ctypes.data)
    return a


def yolo_loss_parts(nb_box, network=None):
    """after batch_loss: (parts[6] averaged over the batch, monitor [B][cells][nb_box][2] = objectness, IoU or -1)"""
    L, net, b, c, cells = _yolo_geom(network)
    parts = np.zeros(6, dtype=np.float32)
    mon = np.zeros((b, cells, nb_box, 2), dtype=np.float32)
    L.cb_yolo_loss_parts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.cb_yolo_loss_parts(net, parts.ctypes.data, mon.ctypes.data)
    return parts, mon


def yolo_boxes(network=None):
    """decoded forward output, reference layout [C][B][cells]: box corners in pixels, then prob / obj / classes / params"""
    L, net, b, c, cells = _yolo_geom(network)
    a = np.zeros((c, b, cells), dtype=np.float32)
    L.cb_yolo_export_boxes.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.cb_yolo_export_boxes(net, a.ctypes.data)
    return a


def set_fusion(on):
    """group-norm + 2x2 max-pool fusion for the layers created from now on (default on)"""
    _load().cb_set_fusion(int(on))


def set_update_plan(on):
    """optimizer sweep of the whole network in three launches (default) or layer by layer; same results bit for bit"""
    _load().cb_set_update_plan(int(on))


def last_conv_impl():
    return core().cb200_last_conv_impl().decode()


def force_simt(on):
    core().cb200_force_simt(int(on))


def last_perf(network=None):
    L = _load()
    net = _net(network)
    return float(L.cb_net_last_items_per_s(net)), float(L.cb_net_last_epoch_loss(net))


def last_accuracy(network=None):
    """fraction of correct argmax predictions of the last validation pass run with confmat > 0"""
    L = _load()
    L.cb_net_last_accuracy.restype = ctypes.c_double
    L.cb_net_last_accuracy.argtypes = [ctypes.c_void_p]
    return float(L.cb_net_last_accuracy(_net(network)))


def reset():
    """Forget all networks (device memory of earlier networks is not reclaimed, as upstream)."""
    L = _load()
    ctypes.c_int.in_dll(L, "nb_networks").value = 0
