"""Launched under torchrun (one rank per GPU).  Every rank builds the same small Darknet-shaped network; ONLY rank 0
receives the reference weights, the other ranks start from different (rank-seeded) ones, so the step can only agree if
cb_dp_init really broadcasts rank 0's parameters.  Then every rank trains TWO steps (momentum carried) on its shard of
a global batch (bucketed NCCL all-reduce of the raw gradients inside the host library), and rank 0 compares the updated
weights with two single-GPU steps on the whole batch (network id 1, same process); every rank also checks that its
replica equals rank 0's bit for bit.  Prints DP_CHECK_OK on success."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cianna_b200 import CIANNA as cnn, cabi, utils  # noqa: E402
from tests import netdefs  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cabi.check(cabi.lib().cb200_init(local))
H = cnn.host()
mode = sys.argv[1] if len(sys.argv) > 1 else "off"
Bg = 8 * world
spec_g = netdefs.tc_darknet(batch=Bg, size=16, classes=16)
spec_l = dict(spec_g, batch=Bg // world)
kinds = [k for k, _ in spec_g["layers"]]
rng = np.random.default_rng(1234)            # same stream on every rank
n = 16 * 16 * 3
x = np.empty((Bg, n + 1), np.float32); x[:, :n] = rng.random((Bg, n), dtype=np.float32) - 0.4; x[:, n] = 0.1
t = np.zeros((Bg, 16), np.float32); t[np.arange(Bg), rng.integers(0, 16, Bg)] = 1
weights = {}
with utils.Quiet():
    utils.build_network(cnn, spec_l, "C_CUDA", mode, network=0)
rng_rank = np.random.default_rng(999 + rank)
for i, k in enumerate(kinds):
    if k in ("conv", "norm"):
        w = cnn.layer_weights(i, network=0)
        if k == "conv":
            weights[i] = (rng.standard_normal(w.shape) * 0.1).astype(np.float32)
        else:
            weights[i] = np.concatenate([1 + 0.2 * rng.standard_normal(w.size // 2), 0.1 * rng.standard_normal(w.size // 2)]).astype(np.float32)
        # rank 0 holds the parameters of the run; every other rank starts somewhere else (as time-seeded initialisers would)
        cnn.set_layer_weights(i, weights[i] if rank == 0 else (rng_rank.standard_normal(w.shape) * 0.3).astype(np.float32), network=0)
idbuf = (ctypes.c_char * 128)()
if rank == 0:
    H.cb_dp_unique_id(idbuf)
tt = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
dist.broadcast(tt, 0)
idbuf = (ctypes.c_char * 128).from_buffer_copy(bytes(tt.cpu().tolist()))
H.cb_dp_init(cnn._net(0), idbuf, rank, world)
b = Bg // world
for step in range(2):
    xs = np.roll(x, step, axis=0)
    ts = np.roll(t, step, axis=0)
    cnn.load_batch(xs[rank * b:(rank + 1) * b], ts[rank * b:(rank + 1) * b], network=0)
    cnn.forward_batch(network=0)
    cnn.backward_batch(0.05, 0.9, 0.001, network=0)
got = {i: cnn.layer_weights(i, network=0) for i, k in enumerate(kinds) if k in ("conv", "norm")}
ok = True
# replicas stay identical: same start (broadcast), same all-reduced gradients, same optimizer arithmetic
for i in sorted(got):
    mine = torch.from_numpy(got[i].copy()).cuda()
    ref0 = mine.clone()
    dist.broadcast(ref0, 0)
    if not torch.equal(mine, ref0):
        ok = False
        print("rank", rank, "layer", i, "replica differs from rank 0 by", float((mine - ref0).abs().max()))
if rank == 0:
    with utils.Quiet():
        utils.build_network(cnn, spec_g, "C_CUDA", mode, network=1)
    for i, w in weights.items():
        cnn.set_layer_weights(i, w, network=1)
    for step in range(2):
        cnn.load_batch(np.roll(x, step, axis=0), np.roll(t, step, axis=0), network=1)
        cnn.forward_batch(network=1)
        cnn.backward_batch(0.05, 0.9, 0.001, network=1)
    tol = 2e-5 if mode == "off" else 4e-3      # two chained steps
    worst = 0.0
    for i in got:
        ref = cnn.layer_weights(i, network=1)
        e = float(np.abs(got[i] - ref).max() / max(np.abs(ref).max(), 1e-30))
        worst = max(worst, e)
        if not e < tol:
            ok = False
            print("layer", i, kinds[i], "rel err", e)
    print("DP_CHECK_%s world=%d mode=%s worst_rel_err=%.3e" % ("OK" if ok else "FAIL", world, mode, worst))
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
if rank == 0 and int(flag.item()) != 0:
    print("DP_CHECK_FAIL on %d rank(s)" % int(flag.item()))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
