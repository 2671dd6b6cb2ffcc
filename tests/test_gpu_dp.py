"""GPU, >= 2 devices: one training step sharded over 2 ranks (NCCL all-reduce inside the host library) must reproduce
the single-GPU step on the whole batch (scripts/dp_check.py).  Skipped on single-GPU boxes; the CPU-side arithmetic is
covered by tests/test_dp_gloo.py."""
import ctypes
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        return ctypes.CDLL(os.path.join(ROOT, "cianna_b200", "libcianna_b200.so")).cb200_device_count()
    except OSError:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["off", "FP16C_FP32A"])
def test_two_gpu_step_equals_single_gpu_step(mode):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dp_check.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "DP_CHECK_OK" in r.stdout and "DP_CHECK_FAIL" not in r.stdout, r.stdout[-3000:]
