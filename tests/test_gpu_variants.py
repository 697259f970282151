"""The reference's compile-time type switches as build variants (util.h:35-53: PRECISION=1 -> float values, UINT_TYPE=2
-> unsigned long long indices): libsparsebench_b200_{f32,u64,f32u64}.so are the same sources compiled with the other
types. Each variant is checked in its own process (the C ABI's struct layouts depend on the types) by
tests/variant_check.py against the reference's own sources compiled with the same switches."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("variant", ["f32", "u64", "f32u64"])
def test_type_variant_against_the_reference_built_with_the_same_switches(variant):
    lib = os.path.join(ROOT, "sparsebench_b200", "libsparsebench_b200_%s.so" % variant)
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_CRS_%s.so" % variant)
    assert os.path.exists(lib), lib + " missing: python -m sparsebench_b200.build"
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libref_CRS_%s.so not built (needs the reference sources: make -C oracle ref)" % variant)
    env = dict(os.environ, SB_VARIANT=variant)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_check.py")], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "variant_check %r: PASS" % variant in r.stdout
