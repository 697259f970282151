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


SPMV_LINE = re.compile(r"^spMVM:\s+([-+\d.eE]+|inf|nan)\s+([-+\d.eE]+|inf|nan)\s+([-+\d.eE]+)", re.M)


def _run_spmv_mode(exe, n, iters):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([exe, "-t", "spmv", "-x", str(n), "-y", str(n), "-z", str(n), "-i", str(iters)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Test type: SPMVM" in r.stdout and "Function   Rate(MB/s)  Rate(MFlop/s)  Walltime(s)" in r.stdout
    m = SPMV_LINE.search(r.stdout)
    assert m, r.stdout[-1500:]
    mbs, mflops, wall2 = float(m.group(1)), float(m.group(2)), float(m.group(3))
    # the reference's accounting (main.c:183-189, profiler.c:127-139): 12 B and 2 flop per ALLOCATED non-zero (27 per row),
    # k = itermax. Walltime is printed with two decimals only; the rate columns carry the full precision.
    wall = 1e-6 * 12 * 27 * n ** 3 * iters / mbs
    assert abs(wall - wall2) <= 0.0051 and abs(mflops / mbs - 2.0 / 12.0) <= 1e-6
    return wall


@pytest.mark.parametrize("fmt,n", [("SCS", 256), ("CRS", 128), ("CCRS", 128)])
def test_reference_driver_spmv_mode(fmt, n):
    """`-t spmv` of the UNMODIFIED main.c (main.c:200-216): it allocate()s x and y, fills them with plain host stores
    (:208-211) and calls spMVM itermax-1 times inside PROFILE(); the reference's profilerPrint reports the rate. Works
    because the ABI-level allocate() hands out unified memory that spMVM moves to the GPU once (that one-time move is
    inside the first PROFILE region; two runs with different iteration counts separate it from the steady state).
    SELL at 256^3: the steady-state time per call the driver reports is within 3 % (+ 15 us: PROFILE's getTimeStamp
    pair drains the device around every call) of the same kernel launched back to back through the API here."""
    import ctypes as C

    import numpy as np

    from sparsebench_b200 import api
    exe = os.path.join(ROOT, "integration", "_build", "sparseBench-%s-B200" % fmt)
    if not os.path.exists(exe):
        pytest.skip("integration/_build not built")
    # the same kernel through the API, x = 1, back to back
    g = api.matrixGenerate(n, n, n, device=True)
    fmt_id = {"CRS": api.FMT_CRS, "SCS": api.FMT_SCS, "CCRS": api.FMT_CCRS}[fmt]
    A = api.convertMatrix(fmt_id, g, 32, 256)
    if fmt != "CCRS":
        api.lib().sbFreeGMatrix(C.byref(g))
    x, y = api.to_device(np.ones(n ** 3)), api.DeviceBuffer(8 * (n ** 3 + 64))
    t = api.EventTimer()
    for _ in range(3):
        api.spMVM(A, x, y)
    t.start()
    for _ in range(100):
        api.spMVM(A, x, y)
    direct_ms = t.stop_ms() / 100
    api.destroyMatrix(A)
    x.free(); y.free()
    if fmt == "CCRS":
        api.lib().sbFreeGMatrix(C.byref(g))
    # The first PROFILE()d call also pays for first-launch work (module load, the one-time move of x and y) and, every
    # few runs, for a GPU that dropped its clocks while the driver spent seconds generating the matrix on the host
    # (0.05 - 1 s until they are back up). Two iteration counts separate the steady state from the one-time part, the
    # minimum over repeated runs drops the clock ramps; pairs of runs are added until the estimate is sane.
    # The 3 % claim is made at 256^3 (0.82 ms per call). At 128^3 a call takes 0.11 ms and the driver's own report
    # (a rate printed with two decimals over 100 calls) resolves no better than ~10 us per call on a busy box: those
    # cases check that the driver runs the real kernel at about its rate (10 % + 30 us).
    rel, slack = (1.03, 0.020) if n >= 256 else (1.10, 0.030)
    short, long_ = 31, 131
    shorts, longs = [], []
    for attempt in range(6):
        shorts.append(_run_spmv_mode(exe, n, short))
        longs.append(_run_spmv_mode(exe, n, long_))
        per_call_ms = (min(longs) - min(shorts)) / (long_ - short) * 1e3
        if attempt >= 1 and 0.9 * direct_ms <= per_call_ms <= rel * direct_ms + slack:
            break
    first_call_extra_ms = min(shorts) * 1e3 - (short - 1) * per_call_ms
    print("%s %d^3 -t spmv: %.4f ms per call reported by the reference driver (steady state; one-time move of x, y to the GPU "
          "%.2f ms), %.4f ms back to back through the API" % (fmt, n, per_call_ms, first_call_extra_ms, direct_ms))
    # PROFILE's getTimeStamp pair drains the device before and after every call: ~15 us of launch + wake-up per call
    assert per_call_ms <= rel * direct_ms + slack, (per_call_ms, direct_ms)
    assert per_call_ms >= 0.9 * direct_ms, (per_call_ms, direct_ms)          # and it really ran the kernel


@pytest.mark.parametrize("fmt,variant,tol", [("CRS", "f32", 2e-3), ("SCS", "f32", 2e-3), ("CRS", "u64", 2e-6)])
def test_reference_driver_of_a_type_variant(fmt, variant, tol):
    """The unmodified main.c compiled with the reference's type switches (-DPRECISION=1 / -DUINT_TYPE=2, util.h:35-53)
    and linked against the matching variant library, against the reference's own executable built with the same
    switches (oracle/_ref/sparseBench-CRS-ref-<v>, run here on the CPU): same lines, same iteration count, residuals
    equal within the precision's tolerance."""
    exe = os.path.join(ROOT, "integration", "_build", "sparseBench-%s-B200-%s" % (fmt, variant))
    ref = os.path.join(ROOT, "oracle", "_ref", "sparseBench-CRS-ref-%s" % variant)
    if not (os.path.exists(exe) and os.path.exists(ref)):
        pytest.skip("variant driver / reference executable not built (make -C integration; make -C oracle ref)")
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    args = ["-x", "16", "-y", "16", "-z", "16", "-i", "20"]
    out = {}
    for name, cmd in (("ref", [ref] + args), ("b200", [exe] + args)):
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        out[name] = [re.sub(r"and took .*", "and took", ln) for ln in r.stdout.splitlines() if KEEP.match(ln)]
    assert len(out["ref"]) == len(out["b200"]) >= 12, out
    first = None
    for a, b in zip(out["b200"], out["ref"]):
        assert NUM.split(a) == NUM.split(b), (a, b)
        for x, y in zip(NUM.findall(a), NUM.findall(b)):
            fx, fy = float(x), float(y)
            first = fy if first is None else first
            assert abs(fx - fy) <= tol * max(abs(fx), abs(fy)) + 1e-5 * tol * first, (a, b)
