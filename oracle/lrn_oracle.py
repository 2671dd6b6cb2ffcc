"""CPU restatement (NumPy) of the reference's local response normalisation layer.

TEST INFRASTRUCTURE ONLY (same rule as oracle/cianna_oracle.py): used by tests/ as the checker of cb200_lrn_*.

Parity status: PINNED to upstream's own CUDA kernels.  Upstream has this layer in its CUDA back-end only (lrn_create
exits for C_NAIV / C_BLAS, src/lrn_layer.c:192-197), so the CPU reference builds cannot produce outputs for it; this
file restates the two CUDA kernels (src/cuda/cuda_lrn_layer.cu:35-101) line by line in the reference layout
[C][B][H*W].  It is checked against tests/golden/lrn_refcuda.npz - tensors around two LRN layers produced on a B200 by
the UNMODIFIED src/cuda/*.cu compiled for sm_100 (oracle/_ref/cuda, oracle/build_ref.sh) - in tests/test_oracle_lrn.py
(CPU), live against that build in tests/test_gpu_backends.py (GPU), and by finite differences.
"""
import numpy as np


def _window(c, C, rng):
    """channels [lo, hi] a channel normalises over (cuda_lrn_layer.cu:54-55: integer range/2 on both sides, clipped)"""
    half = rng // 2
    return max(0, c - half), min(C - 1, c + half)


def lrn_forward(x, rng, k, alpha, beta):
    """cuda_lrn_layer.cu:35-68.  x [C][B][A] -> (y, local_scale), FP32 like the kernel's float arithmetic."""
    x = np.asarray(x, dtype=np.float32)
    C = x.shape[0]
    sq = x.astype(np.float64) ** 2
    scale = np.empty_like(x)
    for c in range(C):
        lo, hi = _window(c, C, rng)
        scale[c] = (k + alpha * sq[lo:hi + 1].sum(axis=0) / rng).astype(np.float32)
    y = (x / np.power(scale.astype(np.float64), beta)).astype(np.float32)
    return y, scale


def lrn_backward(x, y, dy, scale, rng, alpha, beta):
    """cuda_lrn_layer.cu:71-101: dx = dy / s^beta - 2*alpha*beta/range * x * sum_window(dy * y / s)."""
    x, y, dy = (np.asarray(a, dtype=np.float64) for a in (x, y, dy))
    s = np.asarray(scale, dtype=np.float64)
    C = x.shape[0]
    ratio = dy * y / s
    dx = np.empty_like(x)
    for c in range(C):
        lo, hi = _window(c, C, rng)
        dx[c] = dy[c] / np.power(s[c], beta) - 2.0 * alpha * beta * x[c] * ratio[lo:hi + 1].sum(axis=0) / rng
    return dx.astype(np.float32)
