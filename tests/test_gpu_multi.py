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


def test_reference_driver_on_two_gpus():
    """integration/_build/sparseBench-CRS-B200 = the reference's own main.c linked against the library, launched as two
    ranks (torchrun --no-python exports RANK / WORLD_SIZE / LOCAL_RANK, commInit picks them up): the printed residuals
    must equal the single-rank reference run on the same global 16 x 16 x 16 problem."""
    import json
    import re
    if _gpus() < 2:
        pytest.skip("needs at least two GPUs")
    exe = os.path.join(ROOT, "integration", "_build", "sparseBench-CRS-B200")
    if not os.path.exists(exe):
        pytest.skip("integration/_build not built")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_stdout.json")))["gen16_eps"]["lines"]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    port = 29800 + (os.getpid() % 100)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), exe, "-x", "16", "-y", "16", "-z", "8", "-i", "60", "-e", "1e-6"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    keep = re.compile(r"^(Initial Residual|Iteration =|Solution performed|Difference between)")
    lines = [re.sub(r"and took .*", "and took", ln) for ln in r.stdout.splitlines() if keep.match(ln)]
    num = re.compile(r"[-+]?\d+\.\d+E[-+]\d+")
    assert len(lines) == len(gold), (lines, gold)
    for mine, ref in zip(lines, gold):
        assert num.split(mine) == num.split(ref), (mine, ref)
        for a, b in zip(num.findall(mine), num.findall(ref)):
            assert abs(float(a) - float(b)) <= 2e-6 * max(abs(float(a)), abs(float(b))), (mine, ref)
