"""per-layer cost of the group-norm statistics in the conv epilogue: Darknet19 forward convolutions at batch 128,
cb200_conv_forward vs cb200_conv_forward_stats (+ the statistics pass it replaces), CUDA events, median of 7"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cianna_b200 import cabi
cabi.init_device(0)
L = cabi.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
LAYERS = [(32, 224, 64, 3, 8), (64, 112, 128, 3, 8), (128, 112, 64, 1, 8), (128, 56, 256, 3, 16), (256, 56, 128, 1, 16),
          (256, 28, 512, 3, 16), (512, 28, 256, 1, 16), (512, 14, 1024, 3, 32), (1024, 14, 512, 1, 16)]
def ev():
    e = ctypes.c_void_p(); cabi.check(L.cb200_event_create(ctypes.byref(e))); return e
def timeit(fn, n=7):
    ts = []
    for _ in range(n):
        a, b = ev(), ev()
        cabi.check(L.cb200_event_record(a, None)); fn(); cabi.check(L.cb200_event_record(b, None))
        ms = ctypes.c_float(); cabi.check(L.cb200_event_elapsed_ms(a, b, ctypes.byref(ms))); ts.append(ms.value)
    return float(np.median(ts)) * 1e3
rng = np.random.default_rng(0)
tot = [0.0, 0.0, 0.0]
for C, S, N, f, gs in LAYERS:
    conv = cabi.ConvLayer(cabi.FP16, B, C, S, S, N, f, 1, f // 2, bias_value=0.1, act=cabi.activ(cabi.RELU))
    conv.set_weights((rng.standard_normal((N, f * f * C + 1)) / np.sqrt(f * f * C)).astype(np.float32))
    xb = cabi.DevBuf(B * S * S * C * 2)
    cabi.check(L.cb200_memset(xb.ptr, 0x11, xb.nbytes, None))
    norm = cabi.NormLayer(cabi.FP16, B, N, S, S, gs)
    t0 = timeit(lambda: conv.forward(xb))
    impl = L.cb200_last_conv_impl().decode()
    done = [0]
    def fs():
        done[0] = conv.forward_stats(xb, norm)
    t1 = timeit(fs)
    t2 = timeit(lambda: norm.forward(conv.y)) - timeit(lambda: norm.forward(conv.y, stats_ready=1))
    print("%4d->%4d %dx%d @%3d gs %2d %-13s fwd %7.1f us  fwd+stats %7.1f us (fused=%d)  statistics pass %6.1f us" % (C, N, f, f, S, gs, impl, t0, t1, done[0], t2), flush=True)
    tot[0] += t0; tot[1] += t1; tot[2] += t2
    conv.free(); xb.free()
    for b in (norm.y, norm.dx, norm.ws): b.free()
print("total fwd %.1f  fwd+stats %.1f  statistics passes %.1f us" % tuple(tot))
