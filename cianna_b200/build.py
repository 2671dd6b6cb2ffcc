"""Builds the native pieces in-tree (no JIT cache, no site-packages install):

  cianna_b200/libcianna_b200.so   CUDA core + C-ABI (include/cianna_b200.h), nvcc, sm_100a only
  cianna_b200/libcianna_host.so   host-side C library mirroring the reference C API (gcc, C99)

Usage: python -m cianna_b200.build [--force]
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
OBJ = os.path.join(HERE, "build")
CORE_SO = os.path.join(HERE, "libcianna_b200.so")
HOST_SO = os.path.join(HERE, "libcianna_host.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build_core(force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(ROOT, "include", "cianna_b200.h")]
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o])
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(lambda c: _run(c, log=c[-1] + ".log"), jobs))
    if force or jobs or not os.path.exists(CORE_SO):
        _run([NVCC, "-shared", "-o", CORE_SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return CORE_SO


def build_host(force=False):
    srcs = sorted(glob.glob(os.path.join(HOST, "*.c")))
    if not srcs:
        return None
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + [os.path.join(ROOT, "include", "cianna_b200.h")]
    if force or _newer(HOST_SO, srcs + hdrs + [CORE_SO]):
        _run(["gcc", "-O2", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-fPIC", "-shared", "-Wall", "-Wno-unused-result",
              "-I", os.path.join(ROOT, "include"), "-o", HOST_SO] + srcs +
             ["-L", HERE, "-lcianna_b200", "-Wl,-rpath,$ORIGIN", "-lm", "-Wl,-Bsymbolic"])
    return HOST_SO


def build_all(force=False):
    core = build_core(force)
    host = build_host(force)
    return core, host


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
