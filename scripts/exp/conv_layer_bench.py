"""one convolution layer at batch B (FP16): forward and data gradient, us per call (CUDA events, median of 5)
    python scripts/exp/conv_layer_bench.py C S N f [B]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cianna_b200 import cabi
cabi.init_device(0)
L = cabi.lib()
C, S, N, f = (int(v) for v in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 128
conv = cabi.ConvLayer(cabi.FP16, B, C, S, S, N, f, 1, f // 2, bias_value=0.1, act=cabi.activ(cabi.RELU))
rng = np.random.default_rng(0)
conv.set_weights((rng.standard_normal((N, f * f * C + 1)) / np.sqrt(f * f * C)).astype(np.float32))
xb = cabi.DevBuf(B * S * S * cabi.round8(C) * 2)
cabi.check(L.cb200_memset(xb.ptr, 0x2c, xb.nbytes, None))
dyb = cabi.DevBuf(B * S * S * cabi.round8(N) * 2)
cabi.check(L.cb200_memset(dyb.ptr, 0x1c, dyb.nbytes, None))
def ev():
    e = ctypes.c_void_p(); cabi.check(L.cb200_event_create(ctypes.byref(e))); return e
def timeit(fn, n=5):
    ts = []
    fn()
    for _ in range(n):
        a, b = ev(), ev()
        cabi.check(L.cb200_event_record(a, None)); fn(); cabi.check(L.cb200_event_record(b, None))
        ms = ctypes.c_float(); cabi.check(L.cb200_event_elapsed_ms(a, b, ctypes.byref(ms))); ts.append(ms.value)
    return float(np.median(ts)) * 1e3
tf = timeit(lambda: conv.forward(xb)); impl_f = L.cb200_last_conv_impl().decode()
td = timeit(lambda: conv.backward_data(dyb)); impl_d = L.cb200_last_conv_impl().decode()
tw = timeit(lambda: conv.backward_weights(xb, dyb)); impl_w = L.cb200_last_conv_impl().decode()
print("%d->%d %dx%d @%d B=%d: forward %.1f us (%s), data gradient %.1f us (%s), weight gradient %.1f us (%s)" % (C, N, f, f, S, B, tf, impl_f, td, impl_d, tw, impl_w))
