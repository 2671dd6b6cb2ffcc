"""CPU restatement (NumPy) of the reference's algorithm for the conv / pool / group-norm / dense /
softmax-cross-entropy training path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (cianna_b200/) imports this file; it may be used from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, as the CHECKER of the CUDA path.

Parity status: PINNED.  The reference ships no golden vectors, so every function
here is checked (tests/test_oracle.py) against (a) the unmodified reference CPU back-ends compiled into
oracle/_ref/ by oracle/build_ref.sh and driven through oracle/ref_probe.c, when that build is present,
and (b) the fixtures under tests/golden/ that were produced by the same reference build
(tests/golden/make_golden.py).

All tensors use the REFERENCE layouts:
  activations of conv / pool / norm layers : [C][B][H*W]            (src/conv_layer.c:239-245)
  network input / dense activations        : [B][n + 1], bias node last (src/auxil.c:320-329)
  conv filters                             : [N][k*k*C + 1], column c*k*k + ky*k + kx, bias weight last
  dense weights                            : [in + 1][n + 1]          (src/dense_layer.c:253-268)
Accumulations are done in float64 like the NAIV back-end (double `h`, src/naiv/naiv_conv_layer.c:214-227).
"""
import numpy as np

EPS_GN = 0.001  # src/naiv/naiv_norm_layer.c:91,110


# ------------------------------------------------------------------ activations
def relu_forward(x, length, saturation=800.0, leak=0.05):
    """src/activ_functions.c:378-413 (conv layout branch): in place on [C][B][A]; samples >= length -> 0."""
    y = np.where(x <= 0, x * leak, np.where(x > saturation, saturation + (x - saturation) * leak, x)).astype(np.float32)
    y[:, length:, :] = 0
    return y


def relu_deriv(delta, value, length, saturation=800.0, leak=0.05):
    """CUDA semantics (src/cuda/cuda_activ_functions.cu:72-111): the test is on the activated VALUE.
    The CPU twin tests `deriv > saturation` instead (src/activ_functions.c:431) - differs only above 800."""
    d = np.where((value <= 0) | (value > saturation), delta * leak, delta).astype(np.float32)
    d[:, length:, :] = 0
    return d


def softmax_conv(x, length):
    """src/activ_functions.c:790-830: one softmax per sample over ALL filters and positions of [C][B][A]."""
    y = np.zeros_like(x, dtype=np.float32)
    for b in range(min(length, x.shape[1])):
        v = x[:, b, :].astype(np.float32)
        e = np.exp(v - v.max(), dtype=np.float32)
        y[:, b, :] = e / e.sum(dtype=np.float32)
    return y


def softmax_dense(x, length):
    """src/activ_functions.c:760-789: dense layout [B][n+1], bias node forced to 0."""
    y = np.zeros_like(x, dtype=np.float32)
    for b in range(min(length, x.shape[0])):
        v = x[b, :-1].astype(np.float32)
        e = np.exp(v - v.max(), dtype=np.float32)
        y[b, :-1] = e / e.sum(dtype=np.float32)
    return y


def output_delta_conv(out, target, length):
    """delta = o - t with the conv-layout target remap pos = a + (c + b*C)*A
    (src/activ_functions.c:866-880; same expression for the quadratic loss :505-519)."""
    C, B, A = out.shape
    t = target.reshape(B, C, A).transpose(1, 0, 2)
    d = (out - t).astype(np.float32)
    d[:, length:, :] = 0
    return d


def cross_entropy_conv(out, target, length):
    """per-element -t*log(max(o, 1e-6)), src/activ_functions.c:905-936; returns [C][B][A]."""
    C, B, A = out.shape
    t = target.reshape(B, C, A).transpose(1, 0, 2)
    e = (-t * np.log(np.maximum(out, 0.000001))).astype(np.float32)
    e[:, length:, :] = 0
    return e


def quadratic_conv(out, target, length):
    C, B, A = out.shape
    t = target.reshape(B, C, A).transpose(1, 0, 2)
    e = (0.5 * (out - t) ** 2).astype(np.float32)
    e[:, length:, :] = 0
    return e


# ------------------------------------------------------------------ convolution
def im2col(x, first_layer, B, C, H, W, k, stride, pad, bias_value):
    """Gather form of im2col_fct (src/naiv/naiv_conv_layer.c:35-97): rows = (b, oy, ox), columns
    c*k*k + ky*k + kx, last column = bias_value.  x is [B][C*H*W+1] when first_layer else [C][B][H*W]."""
    if first_layer:
        img = x[:, : C * H * W].reshape(B, C, H, W)
    else:
        img = x.reshape(C, B, H, W).transpose(1, 0, 2, 3)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    xp = np.zeros((B, C, H + 2 * pad, W + 2 * pad), dtype=np.float64)
    xp[:, :, pad : pad + H, pad : pad + W] = img
    col = np.zeros((B, Ho, Wo, C * k * k + 1), dtype=np.float64)
    for ky in range(k):
        for kx in range(k):
            patch = xp[:, :, ky : ky + stride * Ho : stride, kx : kx + stride * Wo : stride]  # [B][C][Ho][Wo]
            col[:, :, :, ky * k + kx : C * k * k : k * k] = patch.transpose(0, 2, 3, 1)
    col[:, :, :, -1] = bias_value
    return col.reshape(B * Ho * Wo, C * k * k + 1), Ho, Wo


def conv_forward(x, filters, first_layer, B, C, H, W, k, stride, pad, bias_value):
    """output[f][b*A + a] = sum_j im2col[b*A+a][j] * filters[f][j]  (naiv_conv_layer.c:214-227).
    Returns (pre-activation output [N][B][A], im2col matrix)."""
    col, Ho, Wo = im2col(x, first_layer, B, C, H, W, k, stride, pad, bias_value)
    out = (col @ filters.astype(np.float64).T).T
    return out.reshape(filters.shape[0], B, Ho * Wo).astype(np.float32), col


def conv_backward_data(delta, filters, B, C, H, W, k, stride, pad):
    """delta_prev[c][b][iy][ix] = sum_{f,ky,kx} W[f][c*k*k+ky*k+kx] * delta[f][b][oy][ox] with
    oy*stride + ky - pad = iy  (full correlation with rotated filters, naiv_conv_layer.c:262-332)."""
    N = filters.shape[0]
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    d = delta.reshape(N, B, Ho, Wo).astype(np.float64)
    w = filters[:, : C * k * k].reshape(N, C, k, k).astype(np.float64)
    dxp = np.zeros((C, B, H + 2 * pad, W + 2 * pad), dtype=np.float64)
    for ky in range(k):
        for kx in range(k):
            contrib = np.einsum("fc,fbyx->cbyx", w[:, :, ky, kx], d)
            dxp[:, :, ky : ky + stride * Ho : stride, kx : kx + stride * Wo : stride] += contrib
    return dxp[:, :, pad : pad + H, pad : pad + W].reshape(C, B, H * W).astype(np.float32)


def sgd_update(weights, update, grad, lr, batch, momentum, weight_decay, skip_last=0):
    """update = lr/B*grad + momentum*update ; update += lr*wd*w ; w -= update
    (naiv_conv_layer.c:351-369 + update_weights, src/auxil.c:654-667); skip_last = is_pivot."""
    upd = (lr / batch * grad + momentum * update).astype(np.float32)
    w = weights.astype(np.float32).copy()
    flat_u, flat_w = upd.reshape(-1), w.reshape(-1)
    n = flat_w.size - skip_last
    flat_u[:n] += np.float32(lr * weight_decay) * flat_w[:n]
    flat_w[:n] -= flat_u[:n]
    return w, upd


def conv_weight_grad(col, delta):
    """grad[f][j] = sum_b im2col[b][j] * delta[f][b]  (naiv_conv_layer.c:351-365)."""
    N = delta.shape[0]
    return delta.reshape(N, -1).astype(np.float64) @ col


# ------------------------------------------------------------------ pooling
def pool_forward(x, B, C, H, W, p, stride, pad, ptype):
    """max: first strict maximum in (y, x) scan order, map = y*p + x, -1/0 when the window is empty
    (src/naiv/naiv_pool_layer.c:30-115); avg: mean over the in-bound elements (:118-186)."""
    Ho = (H + 2 * pad - p) // stride + 1
    Wo = (W + 2 * pad - p) // stride + 1
    img = x.reshape(C, B, H, W)
    out = np.zeros((C, B, Ho, Wo), dtype=np.float32)
    pmap = -np.ones((C, B, Ho, Wo), dtype=np.int32)
    if ptype == "MAX":
        found = np.zeros((C, B, Ho, Wo), dtype=bool)
        for py in range(p):
            for px in range(p):
                for oy in range(Ho):
                    iy = oy * stride + py - pad
                    if iy < 0 or iy >= H:
                        continue
                    for ox in range(Wo):
                        ix = ox * stride + px - pad
                        if ix < 0 or ix >= W:
                            continue
                        v = img[:, :, iy, ix]
                        take = (~found[:, :, oy, ox]) | (v > out[:, :, oy, ox])
                        out[:, :, oy, ox] = np.where(take, v, out[:, :, oy, ox])
                        pmap[:, :, oy, ox] = np.where(take, py * p + px, pmap[:, :, oy, ox])
                        found[:, :, oy, ox] = True
    else:
        acc = np.zeros((C, B, Ho, Wo), dtype=np.float64)
        cnt = np.zeros((Ho, Wo), dtype=np.int64)
        for py in range(p):
            for px in range(p):
                for oy in range(Ho):
                    iy = oy * stride + py - pad
                    if iy < 0 or iy >= H:
                        continue
                    for ox in range(Wo):
                        ix = ox * stride + px - pad
                        if ix < 0 or ix >= W:
                            continue
                        acc[:, :, oy, ox] += img[:, :, iy, ix]
                        cnt[oy, ox] += 1
        out = (acc / cnt).astype(np.float32)
    return out.reshape(C, B, Ho * Wo), pmap.reshape(C, B, Ho * Wo)


def pool_backward(delta, pmap, B, C, H, W, p, stride, pad, ptype):
    """gather form: src/naiv/naiv_pool_layer.c:195-318 (avg divides by the FULL window volume)."""
    Ho = (H + 2 * pad - p) // stride + 1
    Wo = (W + 2 * pad - p) // stride + 1
    d = delta.reshape(C, B, Ho, Wo)
    m = pmap.reshape(C, B, Ho, Wo)
    dx = np.zeros((C, B, H, W), dtype=np.float32)
    for oy in range(Ho):
        for ox in range(Wo):
            for py in range(p):
                iy = oy * stride + py - pad
                if iy < 0 or iy >= H:
                    continue
                for px in range(p):
                    ix = ox * stride + px - pad
                    if ix < 0 or ix >= W:
                        continue
                    if ptype == "MAX":
                        dx[:, :, iy, ix] += np.where(m[:, :, oy, ox] == py * p + px, d[:, :, oy, ox], 0)
                    else:
                        dx[:, :, iy, ix] += d[:, :, oy, ox] / np.float32(p * p)
    return dx.reshape(C, B, H * W)


# ------------------------------------------------------------------ group normalisation
def group_norm_forward(x, gamma, beta, group_size, set_off, length):
    """src/naiv/naiv_norm_layer.c:47-147: stats per (sample, group), eps 1e-3, gamma/beta per group."""
    C, B, A = x.shape
    G = C // group_size
    xg = x.reshape(G, group_size, B, A).astype(np.float64)
    mean = xg.mean(axis=(1, 3))  # [G][B]
    var = ((xg - mean[:, None, :, None]) ** 2).mean(axis=(1, 3))
    mean32, var32 = mean.astype(np.float32), var.astype(np.float32)
    xh = (xg - mean32[:, None, :, None].astype(np.float64)) / np.sqrt(var32.astype(np.float64) + EPS_GN)[:, None, :, None]
    y = gamma.astype(np.float64)[:, None, None, None] * xh + beta.astype(np.float64)[:, None, None, None]
    if set_off > 0:
        y[G - set_off :] = xg[G - set_off :]
    y = y.reshape(C, B, A).astype(np.float32)
    y[:, length:, :] = 0
    return y, mean32.T.copy(), var32.T.copy()  # stats as [B][G]


def group_norm_backward(x, delta, gamma, mean, var, group_size, set_off, length):
    """src/naiv/naiv_norm_layer.c:86-104,150-195: returns (delta_in, d_gamma[B][G], d_beta[B][G])."""
    C, B, A = x.shape
    G = C // group_size
    n = group_size * A
    xg = x.reshape(G, group_size, B, A).astype(np.float64)
    dg = delta.reshape(G, group_size, B, A).astype(np.float64)
    mu = mean.T.astype(np.float64)[:, None, :, None]
    rstd = 1.0 / np.sqrt(var.T.astype(np.float64) + EPS_GN)[:, None, :, None]
    d_beta = dg.sum(axis=(1, 3)).astype(np.float32)  # [G][B]
    d_gamma = ((dg * (xg - mu)).sum(axis=(1, 3)) * rstd[:, 0, :, 0]).astype(np.float32)
    dxg = (1.0 / n) * gamma.astype(np.float64)[:, None, None, None] * rstd * (
        n * dg - d_beta.astype(np.float64)[:, None, :, None] - (xg - mu) * rstd * d_gamma.astype(np.float64)[:, None, :, None])
    if set_off > 0:
        dxg[G - set_off :] = dg[G - set_off :]
    dx = dxg.reshape(C, B, A).astype(np.float32)
    dx[:, length:, :] = 0
    return dx, d_gamma.T.copy(), d_beta.T.copy()


def group_norm_update(gamma, beta, gamma_upd, beta_upd, d_gamma, d_beta, lr, batch, momentum, set_off=0):
    """host loop of naiv_norm_layer.c:245-262 (no weight decay on gamma / beta)."""
    G = gamma.size
    gu, bu = gamma_upd.copy(), beta_upd.copy()
    g, b = gamma.copy(), beta.copy()
    for j in range(G - set_off):
        gu[j] = momentum * gu[j] + lr * (d_gamma[:, j].astype(np.float64).sum() / batch)
        bu[j] = momentum * bu[j] + lr * (d_beta[:, j].astype(np.float64).sum() / batch)
        g[j] -= gu[j]
        b[j] -= bu[j]
    return g, b, gu, bu


# ------------------------------------------------------------------ dense
def flatten_for_dense(x, bias_value):
    """flat[b][c*A + a] = x[c][b][a], flat[b][C*A] = bias (src/naiv/naiv_dense_layer.c:33-53)."""
    C, B, A = x.shape
    flat = np.empty((B, C * A + 1), dtype=np.float32)
    flat[:, :-1] = x.transpose(1, 0, 2).reshape(B, C * A)
    flat[:, -1] = bias_value
    return flat


def dense_forward(flat_in, weights):
    """out[b][i] = sum_j W[j][i] * in[b][j] (naiv_dense_layer.c:191-205); pre-activation [B][n+1]."""
    return (flat_in.astype(np.float64) @ weights.astype(np.float64)).astype(np.float32)


def relu_forward_dense(x, length, saturation=800.0, leak=0.05):
    """dense layout branch of ReLU_activation_fct (src/activ_functions.c:388-399): bias node -> 0."""
    y = np.where(x <= 0, x * leak, np.where(x > saturation, saturation + (x - saturation) * leak, x)).astype(np.float32)
    y[:, -1] = 0
    y[length:] = 0
    return y


def relu_deriv_dense(delta, value, length, saturation=800.0, leak=0.05):
    d = np.where((value <= 0) | (value > saturation), delta * leak, delta).astype(np.float32)
    d[:, -1] = 0
    d[length:] = 0
    return d


def output_delta_dense(out, target, length):
    """delta[b][i] = o - t for the n real outputs, 0 for the bias node (src/activ_functions.c:505-513)."""
    d = np.zeros_like(out, dtype=np.float32)
    d[:, :-1] = out[:, :-1] - target
    d[length:] = 0
    return d


def quadratic_dense(out, target, length):
    """per-element 0.5*(o-t)^2 of a dense output layer, bias node and tail samples 0 (src/activ_functions.c:534-546)."""
    e = np.zeros_like(out, dtype=np.float32)
    e[:, :-1] = 0.5 * (out[:, :-1] - target) ** 2
    e[length:] = 0
    return e


# ------------------------------------------------------------------ logistic activation
def _logistic(x, beta, saturation):
    """src/activ_functions.c:636-639: the EXPONENT's argument -beta*x is clamped from above, all in float"""
    t = np.minimum(-np.float32(beta) * x.astype(np.float32), np.float32(saturation)).astype(np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(t, dtype=np.float32))).astype(np.float32)


def logistic_forward(x, length, beta=1.0, saturation=6.0):
    """conv layout [C][B][A] (src/activ_functions.c:644-655); defaults of set_logistic_activ (:593-594)."""
    y = _logistic(x, beta, saturation)
    y[:, length:, :] = 0
    return y


def logistic_forward_dense(x, length, beta=1.0, saturation=6.0):
    """dense layout [B][n+1] (src/activ_functions.c:632-643): bias node -> 0"""
    y = _logistic(x, beta, saturation)
    y[:, -1] = 0
    y[length:] = 0
    return y


def logistic_deriv(delta, value, length, beta=1.0):
    """src/activ_functions.c:668-694: delta *= beta * y * (1 - y) on the activated value"""
    d = (delta * (np.float32(beta) * value * (1.0 - value.astype(np.float64)))).astype(np.float32)
    d[:, length:, :] = 0
    return d


def logistic_deriv_dense(delta, value, length, beta=1.0):
    d = (delta * (np.float32(beta) * value * (1.0 - value.astype(np.float64)))).astype(np.float32)
    d[:, -1] = 0
    d[length:] = 0
    return d
