"""first Darknet19 layer (3 -> 32, 3x3, 448 px, batch 128, FP16) straight from the dataset batch: forward and weight
gradient, us per call (CUDA events, median of 7)"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cianna_b200 import cabi
cabi.init_device(0)
L = cabi.lib()
B, C, S, N, f, pad = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 3, 448, 32, 3, 1
layer = cabi.ConvLayer(cabi.FP16, B, C, S, S, N, f, 1, pad, bias_value=0.1, act=cabi.activ(cabi.RELU))
layer.d.input_is_patches = 2
rng = np.random.default_rng(0)
layer.set_weights((rng.standard_normal((N, f * f * C + 1)) * 0.2).astype(np.float32))
n_in = B * (C * S * S + 1)
src = cabi.DevBuf(n_in * 2)
cabi.check(L.cb200_memset(src.ptr, 0x2c, src.nbytes, None))      # 0x2c2c = 0.0652 in FP16
dy = cabi.DevBuf(B * S * S * N * 2)
cabi.check(L.cb200_memset(dy.ptr, 0x1c, dy.nbytes, None))
def ev():
    e = ctypes.c_void_p(); cabi.check(L.cb200_event_create(ctypes.byref(e))); return e
def timeit(fn, n=7):
    ts = []
    fn()
    for _ in range(n):
        a, b = ev(), ev()
        cabi.check(L.cb200_event_record(a, None)); fn(); cabi.check(L.cb200_event_record(b, None))
        ms = ctypes.c_float(); cabi.check(L.cb200_event_elapsed_ms(a, b, ctypes.byref(ms))); ts.append(ms.value)
    return float(np.median(ts)) * 1e3
tf = timeit(lambda: layer.forward(src))
tw = timeit(lambda: layer.backward_weights(src, dy))
out_gb = B * S * S * N * 2 / 1e9
print("first layer B=%d: forward %.1f us (%.2f TB/s of output), weight gradient %.1f us (%.2f TB/s of dy)" % (B, tf, out_gb / tf * 1e6 / 1e3, tw, out_gb / tw * 1e6 / 1e3))

# group-norm sums out of the forward epilogue (conv_first_fwd_kernel<GN>) against the statistics pass they replace:
# the chain the network runs - forward, then group-norm + 2x2 max-pool
gs = int(os.environ.get("FIRST_GS", "4"))
norm = cabi.NormLayer(cabi.FP16, B, N, S, S, gs, 0, B)
norm.set_params(np.ones((N + gs - 1) // gs, np.float32), np.zeros((N + gs - 1) // gs, np.float32))
pool = cabi.PoolLayer(cabi.FP16, B, N, S, S, 2, 2, 0, cabi.POOL_MAX, length=B)
flag = layer.forward_stats(src, norm)
t_fs = timeit(lambda: layer.forward_stats(src, norm))
t_np0 = timeit(lambda: norm.forward_pool(layer.y, pool, stats_ready=0))
layer.forward_stats(src, norm)
def chain1():
    f = layer.forward_stats(src, norm)
    norm.forward_pool(layer.y, pool, stats_ready=f)
def chain0():
    layer.forward(src)
    norm.forward_pool(layer.y, pool, stats_ready=0)
t_c1, t_c0 = timeit(chain1), timeit(chain0)
print("first layer + group-norm(gs %d) + max-pool: forward %.1f us, forward with the sums %.1f us (fused=%d); norm+pool with its statistics pass %.1f us;"
      " chain %.1f us -> %.1f us" % (gs, tf, t_fs, flag, t_np0, t_c0, t_c1))
