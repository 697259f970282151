"""bench.py's reference arm runs on the CPU (oracle/_ref/libref_CRS_fast.so): check the output contract here --
exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_CRS_fast.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("workload,world", [("sell256", 1), ("strong512sell", 2)])
def test_reference_arm_prints_one_json_line(workload, world):
    env = dict(os.environ, SB_REF_THREADS="2")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    if world > 1:
        env.update(RANK="0", WORLD_SIZE=str(world), LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(world), "--steps", "6",
                        "--warmup", "3", "--ref-planes", "8", "--workload", workload], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cg_gflops" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == world and d["steps"] == 6 and d["warmup"] == 3 and d["value"] > 0
    assert d["scaling"] == ("strong" if workload.startswith("strong") else "weak")
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["same_config"] is False and 0 < d["cpu_baseline"]["sampled_fraction"] < 1      # --ref-planes: a labelled sample


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, env=env)
    assert r.returncode == 0 and r.stdout == ""


@pytest.mark.gpu
def test_b200_arm_contract_on_a_small_workload():
    """the product arm on one GPU (64^3 smoke size): one JSON line carrying every key of the measurement contract"""
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "sell64", "--steps", "20", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["gpu_launches"] == 60 and d["value"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert d["cg"]["residual_final"] < 0.1 * d["cg"]["residual_initial"]      # the timed iterations really solved something
