#!/usr/bin/env python3
"""Headline benchmark: Darknet19 448 px FP16C_FP32A training images/s (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU implementation

One "step" = one mini-batch of the full training hot path (layout import, forward, loss, backward, gradient
exchange, optimizer) over synthetic ImageNet-shaped data with random-init weights.

  value  whole-job images/s with the batches already resident in HBM (dynamic_load = 0 semantics), timed with CUDA
         events on the compute stream over exactly K steps, max over ranks
  e2e    the same metric through the reference-facing API (cnn.train: per step a host->device copy of the batch from
         pinned host memory and a device->host read of the per-sample loss), host buffers in, loss out
  roofline  dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs / CUDA-event time of its launches,
         against the measured dense bf16 peak of MEASURED_PEAKS.json.  The events are taken live in an attribution pass
         of the same K steps run just before the timed region with the weight-gradient overlap switched off, because in
         the timed region itself the weight gradients share the GPU with the rest of the backward sweep
  cpu_baseline  the compiled reference (oracle/_ref, OpenBLAS back-end) on the host cores, bounded sample
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "darknet19_448_train_images_per_sec"
UNIT = "images/s"
TRAIN_GFLOP_PER_IMG = 66.71     # algorithmic 2*M*N*K incl. bias column, fwd + dgrad + wgrad
REF_BATCH = 16                  # images per step of the reference arm / cpu_baseline sample
HYPER = dict(learning_rate=0.003, momentum=0.9, weight_decay=0.0002)   # examples/ImageNET/imagenet_train.py:101-103 upstream


def darknet19_spec(batch, size=448, classes=1000):
    from cianna_b200 import configs
    return configs.darknet19(batch, size, classes)


def synth_batches(nb_batch, batch, size, classes, seed):
    """inputs (U[0,255) - 100)/155 like examples/ImageNET/aux_fct.py:122 upstream, one-hot targets"""
    rng = np.random.default_rng(seed)
    n = nb_batch * batch
    x = ((rng.random((n, size * size * 3), dtype=np.float32) * 255.0 - 100.0) / 155.0).astype(np.float32)
    t = np.zeros((n, classes), dtype=np.float32)
    t[np.arange(n), rng.integers(0, classes, n)] = 1.0
    return x, t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []          # (arrival time, line)
        self.t0 = self.t1 = None

    def begin(self):
        """the timed region starts (the process was started earlier: nvidia-smi needs 0.1-1 s to deliver its first line,
        longer when eight ranks start one each - the timed region of ten 16 ms steps would be over by then)"""
        self.t0 = time.monotonic()

    def end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = (self.t1 if self.t1 is not None else time.monotonic()) + 0.12      # a line describes the 100 ms before it
        window = [ln for t, ln in self.lines if t0 <= t <= t1]
        scope = "timed region"
        if not window:
            # a region shorter than the sampling period: the closest samples under the same load (the steps of the
            # attribution pass and the warm-up step right before the region)
            window = [ln for t, ln in self.lines if t0 - 1.5 <= t <= t1 + 0.3]
            scope = "timed region +- 1.5 s (same workload running)"
        for ln in window:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "scope": scope, "reasons": sorted(reasons)}


def workload_config(args, world):
    """the same `config` object on both arms (ours and --impl reference)"""
    B = args.batch
    return {"workload": "Darknet19 ImageNet classifier 448px %s training, synthetic data" % args.precision, "batch_per_gpu": B,
            "global_batch": B * world, "image_size": args.size, "classes": 1000, "parallelism": "dp%d" % world,
            "l2_policy": "inputs larger than L2 (activations %.1f GB per step)" % (36.5e6 * 2 * B / 1e9),
            "train_gflop_per_image": TRAIN_GFLOP_PER_IMG}


def kernel_source_id():
    """sha256 over the CUDA sources + C-ABI header: ties an ncu summary under profiles/ to the build it was taken on
    (the GPU box has no .git, so a commit id would not be checkable there)"""
    import glob
    import hashlib
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(ROOT, "cianna_b200", "csrc", "*")) + [os.path.join(ROOT, "include", "cianna_b200.h")]):
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1398.3), d.get("hbm_gbs", 6547.5), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """the UNMODIFIED reference (oracle/_ref/omp: src/*.c + OpenBLAS back-end) through its own Python API on the host cores"""
    if rank != 0:
        return
    from oracle import ref_driver as rd, ref_loader
    if not ref_loader.available("omp"):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/omp/CIANNA.so is not built (needs /root/reference at build time)"}))
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(cores))
    b = REF_BATCH          # upstream's own batch for this network (examples/ImageNET/imagenet_train.py:45), whatever --steps is
    spec = darknet19_spec(b, args.size, 1000)
    cnn, _ = ref_loader.load("omp")
    with rd._Quiet():
        rd.build_network(cnn, spec, "C_BLAS", "off", network=0)
    kw = dict(nb_iter=1, control_interv=1000, shuffle_every=0, silent=1, network=0, confmat=0, save_every=0, **HYPER)

    def run(nb, seed):
        x, t = synth_batches(nb, b, args.size, 1000, seed)
        with rd._Quiet():
            cnn.create_dataset("TRAIN", nb * b, x, t, network=0, silent=1)
            t0 = time.perf_counter()
            cnn.train(**kw)
            return time.perf_counter() - t0

    if args.warmup > 0:
        run(args.warmup, 1)
    dt = run(args.steps, 2)
    val = args.steps * b / dt
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gpu_launches": 0,
            "config": workload_config(args, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "%d training steps of %d images each (bounded sample of the workload's %d-image steps; the CPU back-end is FP32 "
                                       "whatever the precision mode), unmodified reference built from /root/reference/src: C_BLAS back-end, OpenBLAS + OpenMP "
                                       "on %d threads" % (args.steps, b, args.batch, cores)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline_sample(size):
    """bounded CPU sample for the main line (rank 0, N = 1): 2 training steps of REF_BATCH images with the compiled
    reference (the same per-step sample as the --impl reference arm)"""
    from oracle import ref_driver as rd, ref_loader
    if not ref_loader.available("omp"):
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: oracle/_ref/omp not built"}
    code = r"""
import sys, os, time, json
sys.path.insert(0, %r)
import bench
from oracle import ref_driver as rd, ref_loader
cnn, _ = ref_loader.load("omp")
spec = bench.darknet19_spec(bench.REF_BATCH, %d, 1000)
with rd._Quiet():
    rd.build_network(cnn, spec, "C_BLAS", "off", network=0)
x, t = bench.synth_batches(2, bench.REF_BATCH, %d, 1000, 3)
with rd._Quiet():
    cnn.create_dataset("TRAIN", 2 * bench.REF_BATCH, x, t, network=0, silent=1)
    t0 = time.perf_counter()
    cnn.train(nb_iter=1, control_interv=1000, shuffle_every=0, silent=1, network=0, confmat=0, save_every=0, **bench.HYPER)
    dt = time.perf_counter() - t0
sys.stderr.write("CPUBASE " + json.dumps({"dt": dt}) + "\n")
""" % (ROOT, size, size)
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OPENBLAS_NUM_THREADS=str(cores))
    try:
        r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=600)
        for ln in r.stderr.splitlines():
            if ln.startswith("CPUBASE "):
                dt = json.loads(ln[8:])["dt"]
                return {"value": 2.0 * REF_BATCH / dt, "unit": UNIT, "cores": cores, "kind": "reference",
                        "sample": "1 epoch of 2 training steps x %d images (Darknet19-448, FP32, unmodified reference C_BLAS back-end, OpenBLAS + OpenMP on %d threads), %.1f s" % (REF_BATCH, cores, dt)}
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "failed: " + r.stderr[-200:]}
    except subprocess.TimeoutExpired:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "timed out after 600 s"}


def inference_object(imgs, ms_inf, ms_inf_e2e, ms_inf_serial, fam_inf, args, peak_tf, peak_hbm, peak_src):
    """Darknet19-448 inference images/s at 1 GPU (BASELINE.json's second metric) with its own roofline: the forward
    implicit-GEMM family against the tensor peak and the group-norm (+ fused pool) family against the HBM peak"""
    out = {"metric": "darknet19_448_inference_images_per_sec", "value": imgs / (ms_inf / 1000.0), "unit": UNIT,
           "e2e": imgs / (ms_inf_e2e / 1000.0), "ms_per_step": ms_inf / args.steps,
           "effective_tflops": imgs / (ms_inf / 1000.0) * 22.36 / 1e3, "fwd_gflop_per_image": 22.36}
    fams = {}
    for k, v in fam_inf.items():
        if v["launches"] == 0 or v["ms"] <= 0:
            continue
        rate = v["work"] / (v["ms"] / 1000.0)
        conv = k.startswith("conv")
        fams[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                   ("tflops" if conv else "gbs"): rate / (1e12 if conv else 1e9), "share_of_step": v["ms"] / ms_inf_serial}
    out["kernel_families"] = fams
    conv = fam_inf.get("conv_fwd_tcgen05")
    if conv and conv["launches"]:
        ach = conv["work"] / (conv["ms"] / 1000.0) / 1e12
        out["roofline"] = {"bound": "tensor", "kernel": "conv_igemm_kernel (forward launches)", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                           "frac": ach / peak_tf, "traffic": None, "peak_source": peak_src, "launches": conv["launches"],
                           "avg_launch_ms": conv["ms"] / conv["launches"], "share_of_step": conv["ms"] / ms_inf_serial,
                           "timing": "CUDA events around each launch in an attribution pass of the same K forward steps"}
    gn = fam_inf.get("group_norm")
    if gn and gn["launches"]:
        ach = gn["work"] / (gn["ms"] / 1000.0) / 1e9
        out["roofline_hbm"] = {"bound": "hbm", "kernel": "norm_apply / norm_pool_fwd kernels", "achieved": ach, "peak": peak_hbm, "unit": "GB/s",
                               "frac": ach / peak_hbm, "traffic": None, "launches": gn["launches"], "share_of_step": gn["ms"] / ms_inf_serial}
    return out


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    from cianna_b200 import CIANNA as cnn
    from cianna_b200 import cabi
    from cianna_b200 import utils

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        dist = dist_mod
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = cabi.lib()
    cabi.check(L.cb200_init(local_rank))
    H = cnn.host()
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    H.cb_train_steps.argtypes = [vp, ci, cf, cf, cf, ci, ci]
    H.cb_forward_steps.argtypes = [vp, ci, ci, ci]
    L.cb200_profile_collect.argtypes = [ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]

    B = args.batch
    spec = darknet19_spec(B, args.size, 1000)
    quiet = utils.Quiet()
    with quiet:
        utils.build_network(cnn, spec, "C_CUDA", args.precision, network=0, dynamic_load=1)
    net = cnn._net(0)
    if world > 1:
        import torch
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            H.cb_dp_unique_id(idbuf)
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        idbuf = (ctypes.c_char * 128).from_buffer_copy(bytes(t.cpu().tolist()))
        H.cb_dp_init(net, idbuf, rank, world)
    cnn.set_TC_scale_factor(256.0, network=0)     # examples/ImageNET/imagenet_train.py upstream: TC_scale_factor=256

    nb = min(args.steps, args.host_batches)
    x, t = synth_batches(nb, B, args.size, 1000, seed=100 + rank)
    with quiet:
        cnn.create_dataset("TRAIN", nb * B, x, t, network=0, silent=1)
        cnn.create_dataset("TEST", min(nb, 2) * B, x[: min(nb, 2) * B], t[: min(nb, 2) * B], network=0, silent=1)
    del x, t
    lr, mom, wd = HYPER["learning_rate"], HYPER["momentum"], HYPER["weight_decay"]

    def barrier():
        cabi.check(L.cb200_device_sync())
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        ev0, ev1 = ctypes.c_void_p(), ctypes.c_void_p()
        cabi.check(L.cb200_event_create(ctypes.byref(ev0)))
        cabi.check(L.cb200_event_create(ctypes.byref(ev1)))
        barrier()
        cabi.check(L.cb200_event_record(ev0, None))
        fn()
        cabi.check(L.cb200_event_record(ev1, None))
        ms = ctypes.c_float()
        cabi.check(L.cb200_event_elapsed_ms(ev0, ev1, ctypes.byref(ms)))
        barrier()
        v = float(ms.value)
        if dist is not None:
            import torch
            mine = torch.tensor([v], device="cuda")
            every = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(every, mine)
            per_rank_ms[:] = [float(e.item()) for e in every]
            v = max(per_rank_ms)          # the job is as slow as its slowest rank
        else:
            per_rank_ms[:] = [v]
        return v

    per_rank_ms = []

    # ---- device-resident steps
    H.cb_last_step_loss.restype = ctypes.c_float
    H.cb_last_step_loss.argtypes = [vp]
    H.cb_train_steps(net, 1, lr, mom, wd, 1, 0)
    loss_first = float(H.cb_last_step_loss(net))
    H.cb_train_steps(net, max(args.warmup - 1, 0), lr, mom, wd, 1, 0)
    barrier()
    # (a) attribution pass: the same K steps with the weight-gradient kernels back on the compute stream, so that every
    #     CUDA-event pair brackets ONE kernel family running alone (in the timed pass below the weight gradients overlap
    #     the rest of the backward sweep on a low-priority stream and per-kernel durations are not separable)
    sampler = ClockSampler(local_rank)
    sampler.start()
    H.cb_set_wgrad_overlap.argtypes = [ctypes.c_void_p, ctypes.c_int]
    H.cb_set_wgrad_overlap(net, 0)
    L.cb200_profile_reset()
    L.cb200_profile_enable(1)
    ms_serial = timed(lambda: H.cb_train_steps(net, args.steps, lr, mom, wd, 1, 0))
    L.cb200_profile_enable(0)
    H.cb_set_wgrad_overlap(net, 1)
    H.cb_train_steps(net, 1, lr, mom, wd, 1, 0)
    barrier()
    # (b) the timed region of `value`
    L.cb200_launch_count(1)
    sampler.begin()
    ms_res = timed(lambda: H.cb_train_steps(net, args.steps, lr, mom, wd, 1, 0))
    sampler.end()
    clocks = sampler.stop()
    rank_ms_res = [m / args.steps for m in per_rank_ms]
    rank_clocks = [clocks]
    if dist is not None:
        rank_clocks = [None] * world
        dist.all_gather_object(rank_clocks, clocks)
        clocks = dict(clocks)
        reasons = sorted(set(r for c in rank_clocks for r in (c or {}).get("reasons", [])))
        clocks["reasons"] = reasons                                   # union over the ranks
        clocks["per_rank"] = [{"sm_mhz": (c or {}).get("sm_mhz"), "reasons": (c or {}).get("reasons")} for c in rank_clocks]
    launches = int(L.cb200_launch_count(0))
    fam = {}
    for f, name in ((0, "conv_fwd_tcgen05"), (1, "conv_dgrad_tcgen05"), (2, "conv_wgrad_tcgen05"), (3, "conv_fwd_simt"), (4, "conv_dgrad_simt"),
                    (5, "conv_wgrad_simt"), (6, "pool"), (7, "group_norm")):
        ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        cabi.check(L.cb200_profile_collect(f, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n)))
        fam[name] = {"ms": ms.value, "work": work.value, "launches": int(n.value)}
    L.cb200_profile_reset()
    loss_last = float(H.cb_last_step_loss(net))     # the timed region's last step (same batches cycled: must be finite and below loss_first)

    # ---- end to end through the reference-facing API: cnn.train over host-resident batches (dynamic_load = 1)
    def e2e_run():
        full, rem = divmod(args.steps, nb)
        with quiet:
            if full:
                cnn.train(nb_iter=full, control_interv=10 ** 6, shuffle_every=0, silent=1, network=0, TC_scale_factor=256.0, **HYPER)
        if rem:
            H.cb_train_steps(net, rem, lr, mom, wd, 0, 1)
    H.cb_train_steps(net, min(args.warmup, 2), lr, mom, wd, 0, 1)
    with quiet:   # one untimed epoch through train(): also takes the per-layer perf_eval sample of the first epoch
        cnn.train(nb_iter=1, control_interv=10 ** 6, shuffle_every=0, silent=1, network=0, TC_scale_factor=256.0, **HYPER)
    ms_e2e = timed(e2e_run)

    # ---- inference (forward only), device resident and end to end
    H.cb_forward_steps(net, 2, 1, 0)
    ms_inf = timed(lambda: H.cb_forward_steps(net, args.steps, 1, 0))
    ms_inf_e2e = timed(lambda: H.cb_forward_steps(net, args.steps, 0, 1))
    # attribution pass of the inference step (events around every launch; not the timed number)
    L.cb200_profile_reset()
    L.cb200_profile_enable(1)
    ms_inf_serial = timed(lambda: H.cb_forward_steps(net, args.steps, 1, 0))
    L.cb200_profile_enable(0)
    fam_inf = {}
    for f, name in ((0, "conv_fwd_tcgen05"), (3, "conv_fwd_simt"), (6, "pool"), (7, "group_norm")):
        ms, work, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        cabi.check(L.cb200_profile_collect(f, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n)))
        fam_inf[name] = {"ms": ms.value, "work": work.value, "launches": int(n.value)}
    L.cb200_profile_reset()

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    imgs = args.steps * B * world
    value = imgs / (ms_res / 1000.0)
    peak_tf, peak_hbm, peak_src = measured_peaks()
    # forward and data gradient are launches of the same implicit-GEMM kernels (conv_igemm_kernel per tap, conv_halo_kernel
    # for 3x3 on large maps, conv_first_fwd_kernel for the first layer); the weight gradient has its own kernels
    kern = {}
    for name, members in (("conv_igemm_kernel", ("conv_fwd_tcgen05", "conv_dgrad_tcgen05")), ("conv_wgrad_kernel", ("conv_wgrad_tcgen05",))):
        agg = {"ms": sum(fam[m]["ms"] for m in members), "work": sum(fam[m]["work"] for m in members), "launches": sum(fam[m]["launches"] for m in members)}
        if agg["launches"] > 0:
            kern[name] = agg
    if kern:
        dom = max(kern, key=lambda k: kern[k]["ms"])
        ach = kern[dom]["work"] / (kern[dom]["ms"] / 1000.0) / 1e12
        traffic, ncu_note = None, None
        tpath = os.path.join(ROOT, "profiles", "r2_conv_metrics_summary.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            key = {"conv_igemm_kernel": "igemm", "conv_wgrad_kernel": "wgrad"}[dom]
            if tj.get("kernel_source_id") != kernel_source_id():
                # an ncu capture of another build says nothing about this one: no traffic figure rather than a stale one
                ncu_note = {"stale": "profiles/r2_conv_metrics_summary.json was captured on kernel sources %s, this build is %s: traffic withheld"
                                     % (tj.get("kernel_source_id"), kernel_source_id())}
            elif tj.get("batch", 128) == B and args.size == 448 and key in tj:
                traffic = tj[key]["avg_dram_bytes_per_launch"]
                ncu_note = {"source": "profiles/r2_step_metrics_b128.csv (ncu, dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one step)",
                            "kernels": tj[key].get("kernels"), "tensor_pipe_pct_time_weighted": tj[key]["tensor_pipe_pct_time_weighted"]}
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": traffic,
                "peak_source": peak_src, "launches": kern[dom]["launches"], "avg_launch_ms": kern[dom]["ms"] / kern[dom]["launches"],
                "algorithmic_flops_per_launch": kern[dom]["work"] / kern[dom]["launches"],
                "share_of_step": kern[dom]["ms"] / ms_serial, "timing": "CUDA events around each launch in an attribution pass of the same K steps with the weight-gradient overlap off (%.2f ms/step serial vs %.2f overlapped)" % (ms_serial / args.steps, ms_res / args.steps), "ncu": ncu_note}
    else:
        dom = max(fam, key=lambda k: fam[k]["ms"])
        ach = fam[dom]["work"] / max(fam[dom]["ms"], 1e-9) * 1e3 / 1e12
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None, "peak_source": peak_src}
    families = {}
    for k, v in fam.items():
        if v["launches"] == 0:
            continue
        rate = v["work"] / (v["ms"] / 1000.0)
        families[k] = {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                       ("tflops" if k.startswith("conv") else "gbs"): rate / (1e12 if k.startswith("conv") else 1e9)}
    es = 2 if args.precision != "off" else 4
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"off": "f32", "FP16C_FP32A": "f16", "BF16C_FP32A": "bf16"}[args.precision], "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": imgs / (ms_e2e / 1000.0), "unit": UNIT, "h2d_bytes_per_step": B * ((args.size * args.size * 3 + 1) + 1000) * es,
                    "d2h_bytes_per_step": B * 4, "api": "cianna_b200.CIANNA.train (dynamic_load=1)"},
            "per_rank_ms_per_step": rank_ms_res,
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "kernel_families": families,
            "effective_tflops": value * TRAIN_GFLOP_PER_IMG / 1e3,
            "inference": inference_object(imgs, ms_inf, ms_inf_e2e, ms_inf_serial, fam_inf, args, peak_tf, peak_hbm, peak_src),
            "loss": {"first_step": loss_first, "last_timed_step": loss_last,
                     "note": "mean cross-entropy of the step's batch; %d synthetic batches are cycled, ln(1000) = 6.91 at random init" % nb}}
    if not (np.isfinite(loss_last) and np.isfinite(loss_first)):
        line["invalid"] = "non-finite loss in the timed region"
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_sample(args.size)
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "skipped (reported at N=1 only)"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=448)
    ap.add_argument("--precision", default="FP16C_FP32A", choices=["off", "FP16C_FP32A", "BF16C_FP32A"])
    ap.add_argument("--host-batches", type=int, default=4, help="distinct synthetic batches kept in pinned host memory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
