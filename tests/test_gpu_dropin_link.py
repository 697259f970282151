"""The drop-in boundary for real: integration/_build/sparseBench-<FMT>-B200 is the reference's OWN driver (main.c,
parameter.c, profiler.c, util.c -- compiled where they lie, see integration/Makefile) linked against
libsparsebench_b200.so. Its solver output must match the stdout of the reference's own executable
(tests/golden/ref_stdout.json, made by tests/golden/make_ref_stdout.py): same lines, same iteration count, residuals
equal in the 7 digits %E prints (last digit may differ by rounding)."""
import json
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_stdout.json")))
KLEIN = os.path.join(ROOT, "tests", "golden", "reference_fixtures", "matrix_band_klein.mtx")
KEEP = re.compile(r"^(Initial Residual|Iteration =|Solution performed|Difference between)")
NUM = re.compile(r"[-+]?\d+\.\d+E[-+]\d+|nan|-nan")


def close_lines(a, b):
    ta, tb = NUM.split(a), NUM.split(b)
    if ta != tb:
        return False
    for x, y in zip(NUM.findall(a), NUM.findall(b)):
        if "nan" in x or "nan" in y:
            if ("nan" in x) != ("nan" in y):
                return False
            continue
        fx, fy = float(x), float(y)
        if abs(fx - fy) > 2e-6 * max(abs(fx), abs(fy)):
            return False
    return True


@pytest.mark.parametrize("fmt", ["CRS", "CCRS", "SCS"])
@pytest.mark.parametrize("case", sorted(GOLD))
def test_reference_driver_linked_against_the_library(fmt, case):
    exe = os.path.join(ROOT, "integration", "_build", "sparseBench-%s-B200" % fmt)
    if not os.path.exists(exe):
        pytest.skip("integration/_build not built (needs the reference sources: make -C integration)")
    args = [a if a != "<klein>" else KLEIN for a in GOLD[case]["args"]]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [re.sub(r"and took .*", "and took", ln) for ln in r.stdout.splitlines() if KEEP.match(ln)]
    gold = GOLD[case]["lines"]
    assert len(lines) == len(gold), (lines, gold)
    for mine, ref in zip(lines, gold):
        assert close_lines(mine, ref), (mine, ref)
    assert "Function   Rate(MB/s)  Rate(MFlop/s)  Walltime(s)" in r.stdout      # the reference's own profilerPrint ran


def test_reference_driver_bmx_round_trip(tmp_path):
    """main.c:42-52 (`-c file.mtx` writes file.bmx) and main.c:72-76 (`.bmx` input) through the linked reference driver:
    the klein matrix converted to .bmx and solved from there prints the klein golden (its values are exact in float32)"""
    import shutil
    exe = os.path.join(ROOT, "integration", "_build", "sparseBench-CRS-B200")
    if not os.path.exists(exe):
        pytest.skip("integration/_build not built")
    mtx = str(tmp_path / "klein.mtx")
    shutil.copy(KLEIN, mtx)
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([exe, "-c", mtx], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and os.path.exists(str(tmp_path / "klein.bmx")), r.stdout + r.stderr
    r = subprocess.run([exe, "-m", str(tmp_path / "klein.bmx"), "-i", "10"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [re.sub(r"and took .*", "and took", ln) for ln in r.stdout.splitlines() if KEEP.match(ln)]
    gold = GOLD["klein"]["lines"]
    assert len(lines) == len(gold) and all(close_lines(a, b) for a, b in zip(lines, gold)), (lines, gold)
