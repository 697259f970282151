"""World-size 2 and 3 on CPU over torch.distributed/gloo: the multi-rank host logic (partition lists, halo exchange
indexing) through a real multi-process transport. See tests/gloo_partition_check.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_halo_exchange_over_gloo(world):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env["OMP_NUM_THREADS"] = "1"
    port = 29700 + (os.getpid() % 200) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "gloo_partition_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert r.stdout.count("gloo_partition_check: PASS") == world
