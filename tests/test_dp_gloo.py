"""CPU-only, world_size = 2 over gloo: the data-parallel arithmetic the host library applies
(cianna_b200/host/network.c: apply_updates + set_hyper) - shard the mini-batch, sum the RAW weight gradients and the
group-norm (d_gamma, d_beta) sums across ranks, then run the optimizer with lr / (B_local * world) - must reproduce the
single-process step on the full batch.  Gradients come from the oracle (the checker), exchanged with torch.distributed.
"""
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
    for idx in sorted(grads):
        g = torch.from_numpy(np.ascontiguousarray(grads[idx], dtype=np.float64))
        dist.all_reduce(g, op=dist.ReduceOp.SUM)          # cb200_dp_allreduce: raw gradients, never momentum
        L = net.layers[idx]
        if L["kind"] == "conv":
            w, _ = co.sgd_update(L["weights"], L["update"], g.numpy(), lr, B * world, mom, wd)   # hyper[0] = lr / (B * world)
            result[idx] = w
        else:
            gsum = g.numpy()
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
