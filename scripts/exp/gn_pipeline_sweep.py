"""Group-norm statistics -> apply: two passes over the whole batch against the chunked launches, on the Darknet19-448 layer
shapes at batch 128 (FP16).  Prints ms per call (mean of REPS after warm-up, CUDA events on the compute stream) for the
forward and the backward pass; tensors of the large layers exceed L2, the small ones run back to back as in the step.

    python scripts/exp/gn_pipeline_sweep.py [batch]
"""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cianna_b200 import cabi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
REPS = int(os.environ.get("GN_SWEEP_REPS", "5"))
# (channels, map size, group size, followed by a 2x2 max-pool)
SHAPES = [(32, 448, 4, True), (64, 224, 8, True), (128, 112, 8, False), (128, 112, 8, True), (256, 56, 16, False),
          (256, 56, 16, True), (512, 28, 16, False), (1024, 14, 32, False)]
VARIANTS = [("two-pass", 0, 0, 0)] + [("chunk %2d MB x%d" % (mb, c), 1, mb * 1024, c) for c in (2, 1, 3) for mb in (24, 32, 48, 64)]
# GN_SWEEP_SHAPES=0,1  GN_SWEEP_VARIANTS=0:0,48:2 (chunk MB : blocks per SM and role; 0:0 = two-pass) restrict the run (ncu captures)
if os.environ.get("GN_SWEEP_SHAPES"):
    SHAPES = [SHAPES[int(i)] for i in os.environ["GN_SWEEP_SHAPES"].split(",")]
if os.environ.get("GN_SWEEP_VARIANTS"):
    VARIANTS = []
    for v in os.environ["GN_SWEEP_VARIANTS"].split(","):
        mb, c = (int(t) for t in v.split(":"))
        VARIANTS.append(("two-pass", 0, 0, 0) if mb == 0 else ("chunk %2d MB x%d" % (mb, c), 1, mb * 1024, c))


def timed(L, fn):
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    L.cb200_event_create(ctypes.byref(e0)); L.cb200_event_create(ctypes.byref(e1))
    for _ in range(2):
        fn()
    L.cb200_event_record(e0, None)
    for _ in range(REPS):
        fn()
    L.cb200_event_record(e1, None)
    ms = ctypes.c_float()
    L.cb200_event_elapsed_ms(e0, e1, ctypes.byref(ms))
    return ms.value / REPS


def main():
    cabi.init_device(0)
    L = cabi.lib()
    L.cb200_event_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    L.cb200_d2d.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.cb200_norm_set_pipeline.argtypes = [ctypes.c_int] * 3
    out = []
    for (C, S, gs, pooled) in SHAPES:
        rng = np.random.default_rng(1)
        n = B * S * S * C
        x = cabi.DevBuf.from_numpy((rng.standard_normal(1 << 20).astype(np.float16)))
        xb = cabi.DevBuf(n * 2)
        # fill by repeating a 2 MB block (values only need to be finite)
        off = 0
        while off < n * 2:
            m = min(1 << 21, n * 2 - off)
            cabi.check(L.cb200_d2d(ctypes.c_void_p(xb.ptr.value + off), x.ptr, m, None))
            off += m
        nl = cabi.NormLayer(cabi.FP16, B, C, S, S, gs)
        pa = cabi.activ(cabi.RELU)
        if pooled:
            pool = cabi.PoolLayer(cabi.FP16, B, C, S, S, 2, 2, 0, cabi.POOL_MAX)
            dpb = cabi.DevBuf(n // 4 * 2)
            cabi.check(L.cb200_d2d(dpb.ptr, xb.ptr, n // 4 * 2, None))
            fwd = lambda: nl.forward_pool(xb, pool)
            bwd = lambda: nl.backward_pool(xb, dpb, pool, pa)
            E = n * 2
            bytes_f, bytes_b = 2.0 * E + 0.25 * n * 3, 3.0 * E + 0.5 * n * 3
        else:
            dyb = cabi.DevBuf(n * 2)
            cabi.check(L.cb200_d2d(dyb.ptr, xb.ptr, n * 2, None))
            fwd = lambda: nl.forward(xb)
            bwd = lambda: nl.backward(xb, dyb, pa)
            bytes_f, bytes_b = 3.0 * n * 2, 5.0 * n * 2
        row = {"shape": "%dch %dpx gs%d%s" % (C, S, gs, " +pool" if pooled else ""), "variants": {}}
        for name, on, kb, ctas in VARIANTS:
            L.cb200_norm_set_pipeline(on, kb, ctas)
            tf, tb = timed(L, fwd), timed(L, bwd)
            row["variants"][name] = {"fwd_ms": round(tf, 4), "bwd_ms": round(tb, 4),
                                     "fwd_alg_TBs": round(bytes_f / tf / 1e9, 2), "bwd_alg_TBs": round(bytes_b / tb / 1e9, 2)}
            print("%-22s %-14s fwd %7.3f ms (%5.2f TB/s alg)  bwd %7.3f ms (%5.2f TB/s alg)" % (
                row["shape"], name, tf, bytes_f / tf / 1e9, tb, bytes_b / tb / 1e9), flush=True)
        out.append(row)
        bufs = [xb, x, nl.y, nl.dx, dpb if pooled else dyb]
        if pooled:
            bufs += [getattr(pool, a) for a in ("y", "dx", "map") if hasattr(pool, a)]
        for b in bufs:
            b.free()
    L.cb200_norm_set_pipeline(0, 24 * 1024, 6)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
