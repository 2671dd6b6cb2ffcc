"""CPU-only, world_size = 2 over gloo: the data-parallel scheme of the host library (cianna_b200/host/network.c) -
shard the mini-batch, lay the RAW weight gradients and the group-norm (d_gamma, d_beta) sums out in ONE arena, exchange
it in the buckets and in the order `cb_dp_plan` dictates, then run the optimizer with lr / (B_local * world) - must
reproduce the single-process step on the full batch.

What is the product's here: the arena layout and the bucket plan come from libcianna_host.so's own planner
(cb_dp_plan: pure C, callable without a device; prepare_training uses the same function).  What is stood in: gloo for
NCCL, and the gradients / optimizer arithmetic are the oracle's (the CUDA kernels cannot run here; the GPU-side check of
the same step is scripts/dp_check.py, run by tests/test_gpu_dp.py on boxes with >= 2 GPUs).
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cianna_oracle as co
from oracle.oracle_net import OracleNet
from tests import netdefs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(spec, seed):
    rng = np.random.default_rng(seed)
    net = OracleNet(spec)
    for L in net.layers:
        if L["kind"] == "conv":
            L["weights"] = (rng.standard_normal(L["weights"].shape) * 0.2).astype(np.float32)
        if L["kind"] == "norm":
            L["gamma"] = (1 + 0.1 * rng.standard_normal(L["gamma"].shape)).astype(np.float32)
    return net


def _raw_grads(net, x, t):
    """forward + backward with lr = 0: weights untouched, raw gradients recovered from the saved tensors"""
    net.forward(x)
    net.backward(t, 0.0, 0.0, 0.0)
    out = {}
    for L in net.layers:
        if L["kind"] == "conv":
            out[L["idx"]] = co.conv_weight_grad(L["col"], L["delta"])
        elif L["kind"] == "norm":
            out[L["idx"]] = np.stack([L["d_gamma"].astype(np.float64).sum(0), L["d_beta"].astype(np.float64).sum(0)])
    return out


def _plan(lens, head):
    """libcianna_host.so's arena / bucket planner"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    H = ctypes.CDLL(os.path.join(root, "cianna_b200", "libcianna_host.so"))
    n = len(lens)
    c_len = (ctypes.c_size_t * n)(*lens)
    off = (ctypes.c_size_t * n)()
    beg, bl, first = (ctypes.c_size_t * 8)(), (ctypes.c_size_t * 8)(), (ctypes.c_int * 8)()
    H.cb_dp_plan.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    nb = H.cb_dp_plan(n, c_len, head, off, beg, bl, first)
    return list(off), [(beg[i], bl[i], first[i]) for i in range(nb)]


def test_bucket_plan_covers_the_arena_in_backward_order():
    # Darknet19's 19 weight-gradient slices (filters x (k*k*C) + bias column) and its 18 group-norm layers' sums
    cfg = [(32, 27), (64, 288), (128, 576), (64, 128), (128, 576), (256, 1152), (128, 256), (256, 1152), (512, 2304), (256, 512),
           (512, 2304), (256, 512), (512, 2304), (1024, 4608), (512, 1024), (1024, 4608), (512, 1024), (1024, 4608), (1000, 1024)]
    lens = [n * k + n for n, k in cfg]
    head = 2 * 450
    off, buckets = _plan(lens, head)
    assert 1 <= len(buckets) <= 8
    assert off[0] >= head and all(o % 64 == 0 for o in off)
    assert all(off[i + 1] >= off[i] + lens[i] for i in range(len(lens) - 1))
    total = off[-1] + (lens[-1] + 63) // 64 * 64
    # contiguous, disjoint, complete; listed in the order the backward sweep completes them (last layers first)
    ordered = sorted(buckets)
    assert ordered[0][0] == 0 and sum(b[1] for b in buckets) == total
    for (b0, l0, _), (b1, _, _) in zip(ordered, ordered[1:]):
        assert b0 + l0 == b1
    assert [b[2] for b in buckets] == sorted((b[2] for b in buckets), reverse=True) and buckets[-1][2] == 0
    for beg, ln, first in buckets[:-1]:
        assert beg == off[first]
    # few large buckets (launch latency, not link count), a small last one (it sits on the critical path)
    assert len(buckets) <= 6 and buckets[-1][1] * 4 <= 8 << 20
    # degenerate shapes
    assert _plan([10], 0)[1] == [(0, 64, 0)]
    off1, b1 = _plan([100, 200], 6)
    assert b1[-1][0] == 0 and sum(x[1] for x in b1) == off1[-1] + 256


def _worker(rank, world, port, spec_full, x, t, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = spec_full["batch"] // world
    spec = dict(spec_full, batch=B)
    net = _make(spec, 5)
    sl = slice(rank * B, (rank + 1) * B)
    grads = _raw_grads(net, x[sl], t[sl])
    lr, mom, wd = 0.05, 0.9, 0.001
    result = {}
    # the arena of prepare_training: group-norm sums at the head (layer order), then the weight-gradient slices
    convs = [i for i in sorted(grads) if net.layers[i]["kind"] == "conv"]
    norms = [i for i in sorted(grads) if net.layers[i]["kind"] == "norm"]
    lens = [grads[i].size for i in convs]
    head = sum(grads[i].size for i in norms)
    off, buckets = _plan(lens, head)
    arena = np.zeros(off[-1] + (lens[-1] + 63) // 64 * 64, dtype=np.float64)
    pos, norm_off = 0, {}
    for i in norms:
        norm_off[i] = pos
        arena[pos:pos + grads[i].size] = grads[i].ravel()
        pos += grads[i].size
    for i, o in zip(convs, off):
        arena[o:o + grads[i].size] = grads[i].ravel()
    for beg, ln, _ in buckets:                               # cb_dp_layer_done: ONE all-reduce per bucket, backward order
        g = torch.from_numpy(arena[beg:beg + ln])
        dist.all_reduce(g, op=dist.ReduceOp.SUM)             # raw gradients, never momentum
    for idx in sorted(grads):
        L = net.layers[idx]
        if L["kind"] == "conv":
            o = off[convs.index(idx)]
            g = arena[o:o + grads[idx].size].reshape(grads[idx].shape)
            w, _ = co.sgd_update(L["weights"], L["update"], g, lr, B * world, mom, wd)   # hyper[0] = lr / (B * world)
            result[idx] = w
        else:
            gsum = arena[norm_off[idx]:norm_off[idx] + grads[idx].size].reshape(grads[idx].shape)
            gam = L["gamma"] - (lr * gsum[0] / (B * world)).astype(np.float32)
            bet = L["beta"] - (lr * gsum[1] / (B * world)).astype(np.float32)
            result[idx] = np.concatenate([gam, bet])
    if rank == 0:
        ret.update({k: v for k, v in result.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_step_equals_full_batch_step():
    spec = netdefs.mini_darknet(batch=4, size=8, classes=5)
    rng = np.random.default_rng(0)
    n = 8 * 8 * 3
    x = np.empty((4, n + 1), np.float32)
    x[:, :n] = rng.random((4, n), dtype=np.float32) - 0.4
    x[:, n] = 0.1
    t = np.zeros((4, 5), np.float32)
    t[np.arange(4), rng.integers(0, 5, 4)] = 1
    # single process, full batch
    ref = _make(spec, 5)
    ref.forward(x)
    ref.backward(t, 0.05, 0.9, 0.001)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), spec, x, t, ret), nprocs=2, join=True)
    assert len(ret) > 0
    for idx, w in ret.items():
        L = ref.layers[idx]
        full = L["weights"] if L["kind"] == "conv" else np.concatenate([L["gamma"], L["beta"]])
        assert np.abs(w - full).max() <= 2e-6 * max(1.0, np.abs(full).max()), idx
