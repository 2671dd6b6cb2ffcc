"""ctypes view of the C-ABI (include/cianna_b200.h): structure mirrors + a tiny device-buffer helper.

Used by the parity tests and bench.py to call the kernels exactly the way a C host would, with host
buffers in and out.  No torch types cross this boundary.
"""
import ctypes

import numpy as np

from . import CIANNA

FP32, FP16, BF16 = 0, 1, 2
RELU, LOGISTIC, SOFTMAX, YOLO, LINEAR = 0, 1, 2, 3, 4
POOL_MAX, POOL_AVG = 0, 1


class Activ(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int), ("leak", ctypes.c_float), ("saturation", ctypes.c_float), ("beta", ctypes.c_float)]


def activ(kind=LINEAR, leak=0.05, saturation=800.0, beta=1.0):
    if kind == LINEAR:
        return Activ(LINEAR, 0.0, 0.0, 0.0)
    if kind == LOGISTIC:
        return Activ(LOGISTIC, 0.0, 6.0 if saturation == 800.0 else saturation, beta)
    return Activ(kind, leak, saturation, 0.0)


class ConvDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("length", ctypes.c_int),
                ("in_c", ctypes.c_int), ("in_h", ctypes.c_int), ("in_w", ctypes.c_int),
                ("out_c", ctypes.c_int), ("out_h", ctypes.c_int), ("out_w", ctypes.c_int),
                ("f_h", ctypes.c_int), ("f_w", ctypes.c_int), ("stride_h", ctypes.c_int), ("stride_w", ctypes.c_int),
                ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int), ("bias_value", ctypes.c_float), ("activ", Activ),
                ("input_is_patches", ctypes.c_int),
                ("in_d", ctypes.c_int), ("out_d", ctypes.c_int), ("f_d", ctypes.c_int), ("stride_d", ctypes.c_int), ("pad_d", ctypes.c_int),
                ("ipad_w", ctypes.c_int), ("ipad_h", ctypes.c_int), ("ipad_d", ctypes.c_int)]


class ConvWeights(ctypes.Structure):
    _fields_ = [("master", ctypes.c_void_p), ("moment", ctypes.c_void_p), ("w_fwd", ctypes.c_void_p), ("w_bwd", ctypes.c_void_p),
                ("bias_w", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("grad_b", ctypes.c_void_p)]


class PoolDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("c", ctypes.c_int), ("in_h", ctypes.c_int), ("in_w", ctypes.c_int),
                ("out_h", ctypes.c_int), ("out_w", ctypes.c_int), ("p_h", ctypes.c_int), ("p_w", ctypes.c_int),
                ("stride_h", ctypes.c_int), ("stride_w", ctypes.c_int), ("pad_h", ctypes.c_int), ("pad_w", ctypes.c_int),
                ("pool_type", ctypes.c_int), ("length", ctypes.c_int), ("activ", Activ),
                ("in_d", ctypes.c_int), ("out_d", ctypes.c_int), ("p_d", ctypes.c_int), ("stride_d", ctypes.c_int), ("pad_d", ctypes.c_int)]


class NormDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("length", ctypes.c_int), ("c", ctypes.c_int), ("h", ctypes.c_int),
                ("w", ctypes.c_int), ("group_size", ctypes.c_int), ("nb_group", ctypes.c_int), ("set_off", ctypes.c_int), ("eps", ctypes.c_float)]


def lib():
    L = CIANNA.core()
    if not getattr(L, "_cabi_ready", False):
        L.cb200_conv_wfwd_elems.restype = ctypes.c_size_t
        L.cb200_conv_wbwd_elems.restype = ctypes.c_size_t
        L.cb200_conv_grad_elems.restype = ctypes.c_size_t
        L.cb200_conv_master_elems.restype = ctypes.c_size_t
        L.cb200_norm_workspace_bytes.restype = ctypes.c_size_t
        L.cb200_dtype_size.restype = ctypes.c_size_t
        L.cb200_malloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
        L.cb200_free.argtypes = [ctypes.c_void_p]
        L.cb200_h2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.cb200_d2h.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.cb200_memset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
        L.cb200_stream_sync.argtypes = [ctypes.c_void_p]
        L.cb200_cast_from_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.cb200_cast_to_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
        L.cb200_import_cbhw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
        L.cb200_export_cbhw.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
        L.cb200_conv_prepare_weights.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.cb200_conv_forward.argtypes = [ctypes.c_void_p] * 5
        L.cb200_conv_backward_data.argtypes = [ctypes.c_void_p] * 7
        L.cb200_conv_backward_weights.argtypes = [ctypes.c_void_p] * 5
        L.cb200_conv_update.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.cb200_event_create.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
        L.cb200_event_record.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.cb200_event_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        L.cb200_host_cast_from_f32.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L._cabi_ready = True
    return L


def init_device(device=0):
    check(lib().cb200_init(int(device)))


def check(rc):
    if rc != 0:
        raise RuntimeError("cb200 call failed (%d): %s" % (rc, lib().cb200_last_error().decode()))


class DevBuf:
    """a device allocation owned through the C-ABI"""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = ctypes.c_void_p()
        check(lib().cb200_malloc(ctypes.byref(self.ptr), self.nbytes))

    @classmethod
    def from_numpy(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        check(lib().cb200_h2d(b.ptr, a.ctypes.data, a.nbytes, None))
        check(lib().cb200_stream_sync(None))
        return b

    def to_numpy(self, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        check(lib().cb200_d2h(out.ctypes.data, self.ptr, out.nbytes, None))
        check(lib().cb200_stream_sync(None))
        return out

    def free(self):
        if self.ptr:
            lib().cb200_free(self.ptr)
            self.ptr = ctypes.c_void_p()


def round8(c):
    return (c + 7) & ~7


def upload_act(x_cbhw, dtype, batch, c, h, w):
    """reference-layout FP32 [C][B][H*W] host array -> device tensor in the core's layout / dtype"""
    L = lib()
    src = DevBuf.from_numpy(np.ascontiguousarray(x_cbhw, dtype=np.float32))
    dst = DevBuf(batch * h * w * round8(c) * L.cb200_dtype_size(dtype))
    check(L.cb200_import_cbhw(dst.ptr, dtype, src.ptr, batch, c, h, w, None))
    check(L.cb200_stream_sync(None))
    src.free()
    return dst


def download_act(buf, dtype, batch, c, h, w):
    L = lib()
    tmp = DevBuf(c * batch * h * w * 4)
    check(L.cb200_export_cbhw(tmp.ptr, buf.ptr, dtype, batch, c, h, w, None))
    out = tmp.to_numpy(np.float32, (c, batch, h * w))
    tmp.free()
    return out


class ConvLayer:
    """one conv layer driven through the C-ABI only (weights in the reference layout [N][k*k*C+1])"""

    def __init__(self, dtype, batch, in_c, in_h, in_w, out_c, f, stride=1, pad=0, bias_value=0.1, act=None, length=None):
        L = lib()
        out_h = (in_h + 2 * pad - f) // stride + 1
        out_w = (in_w + 2 * pad - f) // stride + 1
        self.d = ConvDesc(dtype, batch, batch if length is None else length, in_c, in_h, in_w, out_c, out_h, out_w,
                          f, f, stride, stride, pad, pad, bias_value, act if act is not None else activ(LINEAR), 0)
        self.dtype = dtype
        es = L.cb200_dtype_size(dtype)
        dp = ctypes.byref(self.d)
        self.bufs = dict(
            master=DevBuf(L.cb200_conv_master_elems(dp) * 4), moment=DevBuf(L.cb200_conv_master_elems(dp) * 4),
            w_fwd=DevBuf(L.cb200_conv_wfwd_elems(dp) * es), w_bwd=DevBuf(L.cb200_conv_wbwd_elems(dp) * es),
            bias_w=DevBuf(out_c * 4), grad=DevBuf(L.cb200_conv_grad_elems(dp) * 4), grad_b=DevBuf(out_c * 4))
        self.w = ConvWeights(*[self.bufs[k].ptr for k in ("master", "moment", "w_fwd", "w_bwd", "bias_w", "grad", "grad_b")])
        self.y = DevBuf(batch * out_h * out_w * round8(out_c) * es)
        self.dx = DevBuf(batch * in_h * in_w * round8(in_c) * es)

    def set_weights(self, w_ref):
        L = lib()
        a = np.ascontiguousarray(w_ref, dtype=np.float32)
        check(L.cb200_h2d(self.bufs["master"].ptr, a.ctypes.data, a.nbytes, None))
        check(L.cb200_stream_sync(None))
        check(L.cb200_conv_prepare_weights(ctypes.byref(self.d), ctypes.byref(self.w), None))

    def forward(self, x_buf):
        check(lib().cb200_conv_forward(ctypes.byref(self.d), ctypes.byref(self.w), x_buf.ptr, self.y.ptr, None))
        return self.y

    def forward_stats(self, x_buf, norm):
        """forward + the following group-norm layer's (sum, sum of squares) from the epilogue (cb200_conv_forward_stats);
        returns 1 when norm.ws now holds them"""
        L = lib()
        L.cb200_conv_forward_stats.argtypes = [ctypes.c_void_p] * 8
        done = ctypes.c_int(0)
        check(L.cb200_conv_forward_stats(ctypes.byref(self.d), ctypes.byref(self.w), x_buf.ptr, self.y.ptr, ctypes.byref(norm.d), norm.ws.ptr,
                                         ctypes.byref(done), None))
        return done.value

    def backward_data(self, dy_buf, prev_act=None, prev_out=None):
        pa = ctypes.byref(prev_act) if prev_act is not None else None
        po = prev_out.ptr if prev_out is not None else None
        check(lib().cb200_conv_backward_data(ctypes.byref(self.d), ctypes.byref(self.w), dy_buf.ptr, self.dx.ptr, pa, po, None))
        return self.dx

    def backward_weights(self, x_buf, dy_buf):
        check(lib().cb200_conv_backward_weights(ctypes.byref(self.d), ctypes.byref(self.w), x_buf.ptr, dy_buf.ptr, None))

    def grad_ref_layout(self):
        """raw gradient re-ordered to the reference filter layout [N][k*k*C + 1] (bias column = bias_value * sum dy)"""
        d = self.d
        taps, cp = d.f_h * d.f_w, round8(d.in_c)
        g = self.bufs["grad"].to_numpy(np.float32, (d.out_c, taps, cp))[:, :, : d.in_c]
        gb = self.bufs["grad_b"].to_numpy(np.float32, (d.out_c,))
        out = np.empty((d.out_c, taps * d.in_c + 1), dtype=np.float32)
        out[:, :-1] = g.transpose(0, 2, 1).reshape(d.out_c, d.in_c * taps)
        out[:, -1] = gb * d.bias_value
        return out

    def free(self):
        for b in self.bufs.values():
            b.free()
        self.y.free()
        self.dx.free()


def _ensure_pool_norm_sigs():
    L = lib()
    if getattr(L, "_pn_ready", False):
        return L
    L.cb200_pool_forward.argtypes = [ctypes.c_void_p] * 5
    L.cb200_pool_backward.argtypes = [ctypes.c_void_p] * 7
    L.cb200_export_pool_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]
    L.cb200_norm_forward.argtypes = [ctypes.c_void_p] * 9
    L.cb200_norm_backward.argtypes = [ctypes.c_void_p] * 13
    L._pn_ready = True
    return L


class PoolLayer:
    def __init__(self, dtype, batch, c, in_h, in_w, p, stride=None, pad=0, ptype=POOL_MAX, act=None, length=None):
        L = _ensure_pool_norm_sigs()
        stride = p if stride is None else stride
        out_h = (in_h + 2 * pad - p) // stride + 1
        out_w = (in_w + 2 * pad - p) // stride + 1
        self.d = PoolDesc(dtype, batch, c, in_h, in_w, out_h, out_w, p, p, stride, stride, pad, pad, ptype,
                          batch if length is None else length, act if act is not None else activ(LINEAR))
        es = L.cb200_dtype_size(dtype)
        self.y = DevBuf(batch * out_h * out_w * round8(c) * es)
        self.map = DevBuf(batch * out_h * out_w * round8(c))
        self.dx = DevBuf(batch * in_h * in_w * round8(c) * es)

    def forward(self, x_buf):
        check(lib().cb200_pool_forward(ctypes.byref(self.d), x_buf.ptr, self.y.ptr, self.map.ptr, None))
        return self.y

    def map_ref_layout(self):
        d = self.d
        tmp = DevBuf(d.c * d.batch * d.out_h * d.out_w * 4)
        check(lib().cb200_export_pool_map(tmp.ptr, self.map.ptr, d.batch, d.c, d.out_h, d.out_w, None))
        out = tmp.to_numpy(np.int32, (d.c, d.batch, d.out_h * d.out_w))
        tmp.free()
        return out

    def backward(self, dy_buf, prev_act=None, prev_out=None):
        pa = ctypes.byref(prev_act) if prev_act is not None else None
        po = prev_out.ptr if prev_out is not None else None
        check(lib().cb200_pool_backward(ctypes.byref(self.d), dy_buf.ptr, self.map.ptr, self.dx.ptr, pa, po, None))
        return self.dx


class NormLayer:
    def __init__(self, dtype, batch, c, h, w, group_size, set_off=0, length=None):
        L = _ensure_pool_norm_sigs()
        nb_group = (c + group_size - 1) // group_size
        self.d = NormDesc(dtype, batch, batch if length is None else length, c, h, w, group_size, nb_group, set_off, 0.001)
        es = L.cb200_dtype_size(dtype)
        n = batch * h * w * round8(c) * es
        self.y, self.dx = DevBuf(n), DevBuf(n)
        self.gamma = DevBuf.from_numpy(np.ones(nb_group, np.float32))
        self.beta = DevBuf.from_numpy(np.zeros(nb_group, np.float32))
        self.mean, self.var = DevBuf(batch * nb_group * 4), DevBuf(batch * nb_group * 4)
        self.d_gamma, self.d_beta = DevBuf(batch * nb_group * 4), DevBuf(batch * nb_group * 4)
        self.ws = DevBuf(L.cb200_norm_workspace_bytes(ctypes.byref(self.d)))
        self.colsum = DevBuf(c * 4)
        self.nb_group = nb_group

    def set_params(self, gamma, beta):
        self.gamma.free(); self.beta.free()
        self.gamma = DevBuf.from_numpy(np.ascontiguousarray(gamma, np.float32))
        self.beta = DevBuf.from_numpy(np.ascontiguousarray(beta, np.float32))

    def forward(self, x_buf, stats_ready=0):
        L = lib()
        L.cb200_norm_forward_ex.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int, ctypes.c_void_p]
        check(L.cb200_norm_forward_ex(ctypes.byref(self.d), x_buf.ptr, self.y.ptr, self.gamma.ptr, self.beta.ptr,
                                      self.mean.ptr, self.var.ptr, self.ws.ptr, int(stats_ready), None))
        return self.y

    def backward(self, x_buf, dy_buf, prev_act=None):
        pa = ctypes.byref(prev_act) if prev_act is not None else None
        check(lib().cb200_norm_backward(ctypes.byref(self.d), x_buf.ptr, dy_buf.ptr, self.dx.ptr, self.gamma.ptr, self.mean.ptr,
                                        self.var.ptr, self.d_gamma.ptr, self.d_beta.ptr, pa, self.colsum.ptr, self.ws.ptr, None))
        return self.dx

    def forward_pool(self, x_buf, pool, stats_ready=0):
        """fused group-norm + 2x2 max-pool: writes pool.y / pool.map only"""
        L = lib()
        L.cb200_norm_pool_forward_ex.argtypes = [ctypes.c_void_p] * 10 + [ctypes.c_int, ctypes.c_void_p]
        check(L.cb200_norm_pool_forward_ex(ctypes.byref(self.d), ctypes.byref(pool.d), x_buf.ptr, pool.y.ptr, pool.map.ptr, self.gamma.ptr,
                                           self.beta.ptr, self.mean.ptr, self.var.ptr, self.ws.ptr, int(stats_ready), None))
        return pool.y

    def backward_pool(self, x_buf, dpool_buf, pool, prev_act=None, from_pooled_output=False):
        """from_pooled_output: the backward reductions read the pooled delta and pool.y (cb200_norm_pool_backward_ex)
        instead of the input-sized tensor"""
        L = lib()
        L.cb200_norm_pool_backward.argtypes = [ctypes.c_void_p] * 15
        L.cb200_norm_pool_backward_ex.argtypes = [ctypes.c_void_p] * 17
        pa = ctypes.byref(prev_act) if prev_act is not None else None
        if from_pooled_output:
            check(L.cb200_norm_pool_backward_ex(ctypes.byref(self.d), ctypes.byref(pool.d), x_buf.ptr, dpool_buf.ptr, pool.map.ptr, self.dx.ptr,
                                                self.gamma.ptr, self.mean.ptr, self.var.ptr, self.d_gamma.ptr, self.d_beta.ptr, pa,
                                                self.colsum.ptr, self.ws.ptr, pool.y.ptr, self.beta.ptr, None))
        else:
            check(L.cb200_norm_pool_backward(ctypes.byref(self.d), ctypes.byref(pool.d), x_buf.ptr, dpool_buf.ptr, pool.map.ptr, self.dx.ptr,
                                             self.gamma.ptr, self.mean.ptr, self.var.ptr, self.d_gamma.ptr, self.d_beta.ptr, pa,
                                             self.colsum.ptr, self.ws.ptr, None))
        return self.dx

    def stats(self):
        shp = (self.d.batch, self.nb_group)
        return (self.mean.to_numpy(np.float32, shp), self.var.to_numpy(np.float32, shp),
                self.d_gamma.to_numpy(np.float32, shp), self.d_beta.to_numpy(np.float32, shp))


class LrnDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("length", ctypes.c_int), ("c", ctypes.c_int),
                ("h", ctypes.c_int), ("w", ctypes.c_int), ("range", ctypes.c_int), ("k", ctypes.c_float),
                ("alpha", ctypes.c_float), ("beta", ctypes.c_float)]


class LrnLayer:
    def __init__(self, dtype, batch, c, h, w, rng=5, k=1.0, alpha=1.0, beta=0.5, length=None):
        L = lib()
        L.cb200_lrn_forward.argtypes = [ctypes.c_void_p] * 5
        L.cb200_lrn_backward.argtypes = [ctypes.c_void_p] * 9
        self.d = LrnDesc(dtype, batch, batch if length is None else length, c, h, w, rng, k, alpha, beta)
        n = batch * h * w * round8(c)
        self.n = n
        self.y, self.dx = DevBuf(n * L.cb200_dtype_size(dtype)), DevBuf(n * L.cb200_dtype_size(dtype))
        self.scale = DevBuf(n * 4)

    def forward(self, x_buf, keep_scale=True):
        check(lib().cb200_lrn_forward(ctypes.byref(self.d), x_buf.ptr, self.y.ptr, self.scale.ptr if keep_scale else None, None))
        return self.y

    def backward(self, x_buf, dy_buf, prev_act=None, prev_out=None):
        pa = ctypes.byref(prev_act) if prev_act is not None else None
        po = prev_out.ptr if prev_out is not None else None
        check(lib().cb200_lrn_backward(ctypes.byref(self.d), x_buf.ptr, self.y.ptr, dy_buf.ptr, self.dx.ptr, self.scale.ptr, pa, po, None))
        return self.dx

    def scale_ref_layout(self):
        """local_scale as [C][B][H*W]"""
        d = self.d
        s = self.scale.to_numpy(np.float32, (d.batch, d.h * d.w, round8(d.c)))
        return np.ascontiguousarray(s[:, :, :d.c].transpose(2, 0, 1))


class DropoutDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("length", ctypes.c_int), ("c", ctypes.c_int),
                ("h", ctypes.c_int), ("w", ctypes.c_int), ("drop_rate", ctypes.c_float), ("stream_id", ctypes.c_uint),
                ("seed", ctypes.c_ulonglong), ("draw", ctypes.c_ulonglong), ("activ", Activ)]


class Dropout:
    """cb200_dropout_forward / _backward on caller-owned tensors (in place)"""

    def __init__(self, dtype, batch, c, h, w, rate, act=None, seed=1, stream_id=0, draw=1, length=None):
        L = lib()
        L.cb200_dropout_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.cb200_dropout_backward.argtypes = [ctypes.c_void_p] * 3
        self.d = DropoutDesc(dtype, batch, batch if length is None else length, c, h, w, rate, stream_id, seed, draw,
                             act if act is not None else activ(LINEAR))

    def forward(self, y_buf, scale_only=False):
        check(lib().cb200_dropout_forward(ctypes.byref(self.d), y_buf.ptr, 1 if scale_only else 0, None))
        return y_buf

    def backward(self, dy_buf):
        check(lib().cb200_dropout_backward(ctypes.byref(self.d), dy_buf.ptr, None))
        return dy_buf

    def mask(self):
        """the 0/1 mask in the reference layout [C][B][H*W] (the mask function evaluated on a tensor of ones)"""
        d = self.d
        ones = upload_act(np.ones((d.c, d.batch, d.h * d.w), np.float32), d.dtype, d.batch, d.c, d.h, d.w)
        lin = Dropout(d.dtype, d.batch, d.c, d.h, d.w, d.drop_rate, None, d.seed, d.stream_id, d.draw)
        lin.forward(ones)
        m = download_act(ones, d.dtype, d.batch, d.c, d.h, d.w)
        ones.free()
        return m


class YoloDesc(ctypes.Structure):
    _fields_ = [("dtype", ctypes.c_int), ("batch", ctypes.c_int), ("length", ctypes.c_int), ("grid_h", ctypes.c_int), ("grid_w", ctypes.c_int),
                ("nb_box", ctypes.c_int), ("nb_class", ctypes.c_int), ("nb_param", ctypes.c_int), ("max_nb_obj", ctypes.c_int),
                ("target_stride", ctypes.c_int), ("fit_dim", ctypes.c_int), ("IoU_type", ctypes.c_int), ("prior_dist_type", ctypes.c_int),
                ("error_type", ctypes.c_int), ("class_softmax", ctypes.c_int), ("diff_flag", ctypes.c_int),
                ("strict_box_size_association", ctypes.c_int), ("rand_startup", ctypes.c_int),
                ("rand_prob_best_box_assoc", ctypes.c_float), ("rand_prob", ctypes.c_float), ("min_prior_forced_scaling", ctypes.c_float),
                ("cell_size", ctypes.c_int * 3), ("scale_tab", ctypes.c_float * 6), ("slopes_and_maxes", ctypes.c_float * 18),
                ("IoU_limits", ctypes.c_float * 8), ("fit_parts", ctypes.c_int * 6),
                ("prior_size", ctypes.c_void_p), ("noobj_prob_prior", ctypes.c_void_p), ("param_ind_scale", ctypes.c_void_p)]


IOU_TYPES = {"IoU": 0, "GIoU": 1, "DIoU": 2, "DIoU2": 3}
DIST_TYPES = {"IoU": 0, "IOU": 0, "SIZE": 1, "OFFSET": 2}
_IOU_LIMITS = {0: (0.5, 0.1, 0.0, 0.0, 0.2, 0.2, 0.5, 0.3), 1: (0.4, -0.5, -1.0, -1.0, -0.3, -0.3, 0.4, 0.2),
               2: (0.3, -0.6, -1.0, -1.0, -0.5, -0.5, 0.3, 0.1), 3: (0.3, -0.5, -1.0, -1.0, -0.4, -0.4, 0.3, 0.1)}


class YoloHead:
    """YOLO output head driven through the C-ABI only. `y` takes the keywords of set_yolo_params (prior_size [dims][nb_box]);
    the defaults applied here are upstream's (src/activ_functions.c:1129-1380)."""

    def __init__(self, dtype, batch, grid_h, grid_w, in_w, in_h, y, length=None):
        L = lib()
        L.cb200_yolo_workspace_bytes.restype = ctypes.c_size_t
        L.cb200_yolo_activation.argtypes = [ctypes.c_void_p] * 3
        L.cb200_yolo_delta.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_float, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_ulonglong] + [ctypes.c_void_p] * 3
        L.cb200_yolo_loss.argtypes = [ctypes.c_void_p] * 8
        L.cb200_yolo_export_boxes.argtypes = [ctypes.c_void_p] * 4
        nb_box, nc, npar = y["nb_box"], y.get("nb_class", 0), y.get("nb_param", 0)
        diff = y.get("diff_flag", 0)
        d = YoloDesc()
        d.dtype, d.batch, d.length = dtype, batch, batch if length is None else length
        d.grid_h, d.grid_w = grid_h, grid_w
        d.nb_box, d.nb_class, d.nb_param, d.max_nb_obj = nb_box, nc, npar, y["max_nb_obj_per_image"]
        d.target_stride = 1 + d.max_nb_obj * (7 + npar + diff)
        ps = np.asarray(y["prior_size"], dtype=np.float32)
        d.fit_dim = y.get("fit_dim", 0) or ps.shape[0]
        d.IoU_type = IOU_TYPES.get(y.get("IoU_type", "empty"), 1)
        d.prior_dist_type = DIST_TYPES.get(y.get("prior_dist_type", "empty"), 1)
        d.error_type = 0 if y.get("error_type", "empty") == "complete" else 1
        d.class_softmax, d.diff_flag = y.get("class_softmax", 0), diff
        d.strict_box_size_association = y.get("strict_box_size", 0)
        rs = y.get("rand_startup", -1)
        d.rand_startup = 64000 if rs < 0 else rs
        d.rand_prob_best_box_assoc = max(0.0, y.get("rand_prob_best_box_assoc", 0.0))
        d.rand_prob = max(0.0, y.get("rand_prob", 0.0))
        d.min_prior_forced_scaling = max(0.0, y.get("min_prior_forced_scaling", 0.0))
        d.cell_size[0], d.cell_size[1], d.cell_size[2] = in_w // grid_w, in_h // grid_h, 1
        scales = [2.0, 2.0, 1.0, 2.0, 1.0, 1.0]
        for i, v in enumerate(y.get("error_scales", [-1.0] * 6)):
            if v > 0.0:
                scales[i] = float(v)
        sm = np.array([[1, 6, -6], [1, 1.6, -1.6], [1, 6, -6], [1, 6, -6], [1, 6, -6], [1, 1.2, -0.2]], dtype=np.float32)
        if "slopes_and_maxes" in y:
            u = np.asarray(y["slopes_and_maxes"], dtype=np.float32).reshape(6, 3)
            sm[:, 0] = np.where(u[:, 0] > 0.0, u[:, 0], sm[:, 0])
            sm[:, 1] = np.where(u[:, 1] < 100000.0, u[:, 1], sm[:, 1])
            sm[:, 2] = np.where(u[:, 2] > -100000.0, u[:, 2], sm[:, 2])
        lim = list(_IOU_LIMITS[d.IoU_type])
        for i, v in enumerate(y.get("IoU_limits", [-2.0] * 8)):
            if v > -1.99:
                lim[i] = float(v)
        fit = [1, 1, 1, 1, 1 if nc > 0 else -1, 1 if npar > 0 else -1]
        for i, v in enumerate(y.get("fit_parts", [-2] * 6)):
            if v > -2:
                fit[i] = int(v)
        for i in range(6):
            d.scale_tab[i] = scales[i]
            d.fit_parts[i] = fit[i]
        for i in range(18):
            d.slopes_and_maxes[i] = float(sm.ravel()[i])
        for i in range(8):
            d.IoU_limits[i] = lim[i]
        prior = np.zeros((nb_box, 3), dtype=np.float32)
        prior[:, : d.fit_dim] = ps[: d.fit_dim].T
        prior = np.maximum(prior, 1.0)
        noobj = np.asarray(y.get("prior_noobj_prob", [0.2] * nb_box), dtype=np.float32)
        pis = np.asarray(y.get("param_ind_scales", [1.0] * max(npar, 1)), dtype=np.float32)
        self.tables = [DevBuf.from_numpy(prior), DevBuf.from_numpy(noobj), DevBuf.from_numpy(pis)]
        d.prior_size, d.noobj_prob_prior, d.param_ind_scale = [t.ptr.value for t in self.tables]
        self.d, self.dtype = d, dtype
        self.C = nb_box * (8 + nc + npar)
        self.cells = grid_h * grid_w
        self.ws = DevBuf(L.cb200_yolo_workspace_bytes(ctypes.byref(d)))
        es = L.cb200_dtype_size(dtype)
        self.delta = DevBuf(batch * self.cells * round8(self.C) * es)
        self.state = DevBuf(batch * self.cells * nb_box * 4)
        self.loss_buf = DevBuf(batch * 4)
        self.parts = DevBuf(batch * 6 * 4)
        self.monitor = DevBuf(batch * self.cells * nb_box * 2 * 4)

    def upload_targets(self, t):
        """FP32 rows [B][target_stride] -> device buffer in the head's dtype (round toward zero like the dataset path)"""
        L = lib()
        t = np.ascontiguousarray(t, dtype=np.float32)
        host = np.empty(t.size * L.cb200_dtype_size(self.dtype), dtype=np.uint8)
        L.cb200_host_cast_from_f32(host.ctypes.data, self.dtype, t.ctypes.data, t.size)
        return DevBuf.from_numpy(host)

    def activation(self, y_buf):
        check(lib().cb200_yolo_activation(ctypes.byref(self.d), y_buf.ptr, None))

    def deriv_error(self, y_buf, t_buf, tc_scale=1.0, nb_im_iter=10**9, seed=1, step=0):
        d = self.d
        check(lib().cb200_yolo_delta(ctypes.byref(d), self.delta.ptr, y_buf.ptr, t_buf.ptr, tc_scale, nb_im_iter, seed, step,
                                     self.state.ptr, self.ws.ptr, None))
        return (download_act(self.delta, self.dtype, d.batch, self.C, d.grid_h, d.grid_w),
                self.state.to_numpy(np.int32, (d.batch, self.cells, d.nb_box)))

    def loss(self, y_buf, t_buf):
        d = self.d
        check(lib().cb200_yolo_loss(ctypes.byref(d), self.loss_buf.ptr, self.parts.ptr, self.monitor.ptr, y_buf.ptr, t_buf.ptr, self.ws.ptr, None))
        return (self.loss_buf.to_numpy(np.float32, (d.batch,)), self.parts.to_numpy(np.float32, (d.batch, 6)),
                self.monitor.to_numpy(np.float32, (d.batch, self.cells, d.nb_box, 2)))

    def boxes(self, y_buf):
        d = self.d
        tmp = DevBuf(self.C * d.batch * self.cells * 4)
        check(lib().cb200_yolo_export_boxes(ctypes.byref(d), tmp.ptr, y_buf.ptr, None))
        out = tmp.to_numpy(np.float32, (self.C, d.batch, self.cells))
        tmp.free()
        return out
