"""Full-size fixture of the HEADLINE configuration: one training step of Darknet19 at 448 px, batch 16 (the batch of
examples/ImageNET/imagenet_train.py:45-103 upstream), run by the UNMODIFIED reference CPU back-end (oracle/_ref/omp,
C_BLAS, FP32) on seeded inputs and seeded weights.

Run where /root/reference exists (needs ~6 GB and about a minute on 8 cores):   python tests/golden/make_golden_darknet19.py

The tensors of this network are 100-400 MB each, so the fixture is COMPACT: per layer the L2 norm and a fixed sample of
4096 elements (seeded positions) of the output, of the delta and - for conv layers - of the momentum buffer after the
step (= the weight update), plus in full: the class probabilities, the per-sample loss, every group-norm layer's
statistics (mean / var per sample and group), d_gamma / d_beta and updated gamma / beta.  Weights and inputs are NOT
stored: both sides regenerate them from the seeds below (`seeded_weights`, `seeded_batch`); checksums guard the draw.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))
BATCH, SIZE, CLASSES = 16, 448, 1000
HYPER = dict(lr=0.003, momentum=0.9, weight_decay=0.0002)   # examples/ImageNET/imagenet_train.py:101-103 upstream
NSAMPLE = 4096


def seeded_weights(kind, layer_index, shape):
    """conv: He-normal rows [nb_filters][k*k*C + 1] (bias column drawn like the rest); norm: gamma 1 +- 0.2, beta +- 0.1"""
    rng = np.random.default_rng(5000 + layer_index)
    if kind == "conv":
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(2.0 / shape[1]))).astype(np.float32)
    g = 1.0 + 0.2 * rng.standard_normal(shape, dtype=np.float32)
    b = 0.1 * rng.standard_normal(shape, dtype=np.float32)
    return np.concatenate([g, b]).astype(np.float32)


def seeded_batch(batch=BATCH, size=SIZE, classes=CLASSES, seed=77, bias=0.1):
    """dataset-layout batch [B][3*size*size + 1], pixel values (U[0,255) - 100) / 155 as examples/ImageNET/aux_fct.py:122"""
    rng = np.random.default_rng(seed)
    n = size * size * 3
    x = np.empty((batch, n + 1), dtype=np.float32)
    x[:, :n] = (rng.random((batch, n), dtype=np.float32) * 255.0 - 100.0) / 155.0
    x[:, n] = bias
    t = np.zeros((batch, classes), dtype=np.float32)
    t[np.arange(batch), rng.integers(0, classes, batch)] = 1.0
    return x, t


def sample_positions(layer_index, what, size):
    return np.random.default_rng(9000 + 10 * layer_index + what).integers(0, size, NSAMPLE)


def summarize(out, key, arr, layer_index, what):
    a = np.asarray(arr, dtype=np.float32).ravel()
    pos = sample_positions(layer_index, what, a.size)
    out[key + "_sample"] = a[pos].copy()
    out[key + "_l2"] = np.array([np.sqrt(np.sum(a.astype(np.float64) ** 2))])
    out[key + "_absmax"] = np.array([np.abs(a).max()])


def compare(g, key, arr, layer_index, what):
    """deviation of a full tensor from the fixture's summary of it: `sample` max |a - ref| over the sampled positions /
    max |ref tensor|, `q98` the same at the 98 % quantile, `l2` relative difference of the whole-tensor L2 norms"""
    a = np.asarray(arr, dtype=np.float32).ravel()
    pos = sample_positions(layer_index, what, a.size)
    ref = g[key + "_sample"].astype(np.float64)
    scale = max(float(g[key + "_absmax"][0]), 1e-30)
    d = np.abs(a[pos].astype(np.float64) - ref) / scale
    l2 = float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
    ref_l2 = max(float(g[key + "_l2"][0]), 1e-30)
    return {"sample": float(d.max()), "q98": float(np.quantile(d, 0.98)), "l2": abs(l2 - ref_l2) / ref_l2}


def full_dev(a, ref):
    return float(np.abs(np.asarray(a, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def selfdev():
    """The reference against ITSELF: the same step on its C_NAIV back-end (plain loops, another summation order than
    OpenBLAS) compared with the C_BLAS fixture, quantity by quantity, with the comparison the GPU test applies to the
    product -> tests/golden/darknet19_448_b16_selfdev.npz.  This is the floor any second implementation of the same
    arithmetic sits on: the FP32 forward pass of this 43-layer network reproduces to ~1e-6..1e-5 only, and every
    leaky-ReLU / max-pool decision taken on a value within that distance of its threshold flips - a full-size error on
    one delta element, spread by the layers below it."""
    from cianna_b200 import configs
    from oracle import ref_driver as rd
    g = dict(np.load(os.path.join(HERE, "darknet19_448_b16.npz")))
    spec = configs.darknet19(BATCH, SIZE, CLASSES)
    t0 = time.time()
    ref = rd.RefNet(spec, "C_NAIV", variant="omp")
    for l in range(ref.n_layers):
        t = ref.layer_type(l)
        if t == rd.CONV:
            w = ref.weights_view(l)
            w[...] = seeded_weights("conv", l, w.shape)
        elif t == rd.NORM:
            ga, be = ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")
            gb = seeded_weights("norm", l, ga.shape)
            ga[...], be[...] = gb[:ga.size], gb[ga.size:]
    x, tgt = seeded_batch()
    ref.forward(x)
    print("reference (C_NAIV) forward: %.1f s" % (time.time() - t0)); t0 = time.time()
    out = {}

    def put(name, e):
        for k, v in e.items():
            out["%s_%s" % (name, k)] = np.array([v])
    kinds = [k for k, _ in spec["layers"]]
    for l, k in enumerate(kinds):
        put("out_%d_%s" % (l, k), compare(g, "out_%d" % l, ref.output(l), l, 0))
        if ref.layer_type(l) == rd.NORM:
            out["mean_%d" % l] = np.array([full_dev(ref.norm_view(l, "mean"), g["mean_%d" % l])])
            out["var_%d" % l] = np.array([full_dev(ref.norm_view(l, "var"), g["var_%d" % l])])
    out["probs"] = np.array([full_dev(ref.output(ref.n_layers - 1), g["probs"])])
    loss = float(ref.loss(tgt).sum(axis=(0, 2)).mean())
    out["loss"] = np.array([abs(loss - float(g["loss"].mean())) / float(g["loss"].mean())])
    ref.backward(tgt, HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"])
    print("reference (C_NAIV) backward: %.1f s" % (time.time() - t0))
    for l, k in enumerate(kinds):
        put("delta_%d_%s" % (l, k), compare(g, "delta_%d" % l, ref.delta(l), l, 1))
        if k == "conv":
            put("m1_%d" % l, compare(g, "m1_%d" % l, ref.moment_view(l), l, 2))
            w1 = ref.weights_view(l)
            put("dw_%d" % l, compare(g, "dw_%d" % l, w1 - seeded_weights("conv", l, w1.shape), l, 3))
        elif k == "norm":
            out["dgamma_%d" % l] = np.array([full_dev(ref.norm_view(l, "d_gamma"), g["dgamma_%d" % l])])
            out["dbeta_%d" % l] = np.array([full_dev(ref.norm_view(l, "d_beta"), g["dbeta_%d" % l])])
            w1 = np.concatenate([ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")])
            out["gn_w1_%d" % l] = np.array([full_dev(w1, g["w1_%d" % l])])
    path = os.path.join(HERE, "darknet19_448_b16_selfdev.npz")
    np.savez_compressed(path, **out)
    worst = sorted(((float(v[0]), k) for k, v in out.items()), reverse=True)[:12]
    print("wrote", path, "largest self-deviations:", worst)


def main():
    from cianna_b200 import configs
    from oracle import ref_driver as rd
    spec = configs.darknet19(BATCH, SIZE, CLASSES)
    t0 = time.time()
    ref = rd.RefNet(spec, "C_BLAS", variant="omp")
    out = {}
    wsum = 0.0
    for l in range(ref.n_layers):
        t = ref.layer_type(l)
        if t == rd.CONV:
            w = ref.weights_view(l)
            w[...] = seeded_weights("conv", l, w.shape)
            wsum += float(np.abs(w).sum(dtype=np.float64))
        elif t == rd.NORM:
            g, b = ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")
            gb = seeded_weights("norm", l, g.shape)
            g[...], b[...] = gb[:g.size], gb[g.size:]
            wsum += float(np.abs(gb).sum(dtype=np.float64))
    out["weights_abs_sum"] = np.array([wsum])
    x, tgt = seeded_batch()
    out["x_abs_sum"] = np.array([np.abs(x).sum(dtype=np.float64)])
    out["t_argmax"] = tgt.argmax(axis=1).astype(np.int32)
    print("network + seeded weights: %.1f s" % (time.time() - t0)); t0 = time.time()
    ref.forward(x)
    print("reference forward: %.1f s" % (time.time() - t0)); t0 = time.time()
    last = ref.n_layers - 1
    for l in range(ref.n_layers):
        summarize(out, "out_%d" % l, ref.output(l), l, 0)
        if ref.layer_type(l) == rd.NORM:
            out["mean_%d" % l] = ref.norm_view(l, "mean").copy()
            out["var_%d" % l] = ref.norm_view(l, "var").copy()
    out["probs"] = ref.output(last).copy()                      # [1000][16][1]
    out["loss"] = ref.loss(tgt).sum(axis=(0, 2)).astype(np.float32)   # per sample
    ref.backward(tgt, HYPER["lr"], HYPER["momentum"], HYPER["weight_decay"])
    print("reference backward: %.1f s" % (time.time() - t0)); t0 = time.time()
    for l in range(ref.n_layers):
        summarize(out, "delta_%d" % l, ref.delta(l), l, 1)
        t = ref.layer_type(l)
        if t == rd.CONV:
            summarize(out, "m1_%d" % l, ref.moment_view(l), l, 2)
            w1 = ref.weights_view(l)
            summarize(out, "dw_%d" % l, w1 - seeded_weights("conv", l, w1.shape), l, 3)
        elif t == rd.NORM:
            out["w1_%d" % l] = np.concatenate([ref.norm_view(l, "gamma"), ref.norm_view(l, "beta")]).copy()
            out["dgamma_%d" % l] = ref.norm_view(l, "d_gamma").copy()
            out["dbeta_%d" % l] = ref.norm_view(l, "d_beta").copy()
    path = os.path.join(HERE, "darknet19_448_b16.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024), "loss per sample", out["loss"][:4])


if __name__ == "__main__":
    if "--selfdev" in sys.argv:
        selfdev()
    else:
        main()
