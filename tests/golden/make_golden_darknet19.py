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
    main()
