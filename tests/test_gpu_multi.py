"""Multi-GPU parity (needs >= 2 GPUs; skipped on a single-GPU box): launches tests/mgpu_check.py under torchrun,
one rank per GPU, once per transport (NVLink peer windows, NCCL)."""
import os
import subprocess
import sys

import pytest

from sparsebench_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        return api.lib().sbDeviceCount()
    except Exception:
        return 0


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_multi_gpu_parity(mode):
    n = min(_gpus(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    env = dict(os.environ, SB_COMM=mode)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    port = 29600 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert r.stdout.count("mgpu_check: PASS") == n
