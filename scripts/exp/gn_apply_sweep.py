"""Grid sizing of the group-norm kernels (default two-pass path) on the Darknet19-448 layer shapes at batch 128 (FP16):
blocks wanted per SM x most pixels per block.  us per call, mean of REPS after warm-up, CUDA events.

    python scripts/exp/gn_apply_sweep.py [batch]
"""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cianna_b200 import cabi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
REPS = 5
SHAPES = [(32, 448, 4, True), (64, 224, 8, True), (128, 112, 8, False), (128, 112, 8, True), (64, 112, 8, False), (256, 56, 16, False),
          (256, 56, 16, True), (128, 56, 16, False), (512, 28, 16, False), (512, 28, 16, True), (256, 28, 16, False), (1024, 14, 32, False), (512, 14, 16, False)]
COUNT = [1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 3, 2]      # how often the shape occurs in Darknet19
KNOBS = [(8, 4096)]


def timed(L, fn):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    L.cb200_event_create(ctypes.byref(e0)); L.cb200_event_create(ctypes.byref(e1))
    fn(); fn()
    L.cb200_event_record(e0, None)
    for _ in range(REPS):
        fn()
    L.cb200_event_record(e1, None)
    ms = ctypes.c_float()
    L.cb200_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    return ms.value / REPS * 1e3


def main():
    cabi.init_device(0)
    L = cabi.lib()
    L.cb200_event_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.cb200_d2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.cb200_norm_set_tuning.argtypes = [ctypes.c_int, ctypes.c_int]
    rng = np.random.default_rng(1)
    seedbuf = cabi.DevBuf.from_numpy(rng.standard_normal(1 << 20).astype(np.float16))
    totals = {k: [0.0, 0.0] for k in KNOBS}
    for (C, S, gs, pooled), cnt in zip(SHAPES, COUNT):
        n = B * S * S * C
        xb = cabi.DevBuf(n * 2)
        off = 0
        while off < n * 2:
            m = min(1 << 21, n * 2 - off)
            cabi.check(L.cb200_d2d(ctypes.c_void_p(xb.ptr.value + off), seedbuf.ptr, m, None))
            off += m
        pa = cabi.activ(cabi.RELU)
        line = "%4d ch @%3d gs %2d %s |" % (C, S, gs, "pool" if pooled else "    ")
        for knob in KNOBS:
            L.cb200_norm_set_tuning(*knob)
            nl = cabi.NormLayer(cabi.FP16, B, C, S, S, gs)
            if pooled:
                pool = cabi.PoolLayer(cabi.FP16, B, C, S, S, 2, 2, 0, cabi.POOL_MAX)
                dpb = cabi.DevBuf(n // 4 * 2)
                cabi.check(L.cb200_d2d(dpb.ptr, xb.ptr, n // 4 * 2, None))
                f = timed(L, lambda: nl.forward_pool(xb, pool))
                b = timed(L, lambda: nl.backward_pool(xb, dpb, pool, pa, from_pooled_output=True))
                for buf in (pool.y, pool.map, pool.dx, dpb):
                    buf.free()
            else:
                dyb = cabi.DevBuf(n * 2)
                cabi.check(L.cb200_d2d(dyb.ptr, xb.ptr, n * 2, None))
                f = timed(L, lambda: nl.forward(xb))
                b = timed(L, lambda: nl.backward(xb, dyb, pa))
                dyb.free()
            for buf in (nl.y, nl.dx, nl.ws):
                buf.free()
            totals[knob][0] += f * cnt
            totals[knob][1] += b * cnt
            line += " %5.0f/%5.0f" % (f, b)
        xb.free()
        print(line, flush=True)
    print("knobs (blocks per SM, max pixels per block):", KNOBS)
    print("Darknet19 totals fwd/bwd us:", " ".join("%5.0f/%5.0f" % tuple(totals[k]) for k in KNOBS))


if __name__ == "__main__":
    main()
