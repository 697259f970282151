"""Builds libsparsebench_b200.so (+ the three link-time drop-in shims) in-tree with nvcc for sm_100a.

    python -m sparsebench_b200.build [--force]

The built .so files are git-ignored but travel to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libsparsebench_b200.so")
SOURCES = ["runtime.cu", "generate.cu", "formats.cu", "spmv.cu", "vecops.cu", "cg.cu", "krylov.cu", "comm.cu", "partition.cpp", "mmio.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC,-fvisibility=default",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]
if os.environ.get("SB_BUILD_SWEEPS"):       # every CRS/CCRS pipeline configuration of the sweeps in profiles/ (SB_ROWS_VAR, SB_ROWS_CFG): 4x the compile time
    NVCC_FLAGS.append("-DSB_TUNING_SWEEPS")


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _headers():
    hs = [os.path.join(ROOT, "include", "sparsebench_b200.h")]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs


# The reference's compile-time type switches (util.h:35-53) as build variants of the same sources.
VARIANTS = {
    "": [],                                              # double, unsigned int: libsparsebench_b200.so (the graded configs)
    "f32": ["-DPRECISION=1"],                            # float values
    "u64": ["-DUINT_TYPE=2"],                            # unsigned long long indices
    "f32u64": ["-DPRECISION=1", "-DUINT_TYPE=2"],
}


def lib_path(variant=""):
    return os.path.join(HERE, "libsparsebench_b200%s.so" % ("_" + variant if variant else ""))


def build(force=False, verbose=False, variants=None):
    """builds the default library and (variants=None: all) the type variants; returns the default library's path.
    All variants' compilations run concurrently (a clean build of the four is bounded by 4 x spmv.cu otherwise)."""
    todo = list(VARIANTS if variants is None else variants)
    started = [(v, _compile(v, force, verbose)) for v in todo]
    for v, (objs, procs) in started:
        _link(v, objs, procs, force, verbose)
    return LIB


def _compile(variant, force, verbose):
    defs = VARIANTS[variant]
    obj_dir = OBJ + ("_" + variant if variant else "")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _newer([s] + _headers(), o):
            cmd = [nvcc] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    return objs, procs


def _link(variant, objs, procs, force, verbose):
    defs = VARIANTS[variant]
    lib = lib_path(variant)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s (variant %r)\n%s\n" % (src, variant, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (variant %r)" % variant)
    if force or procs or not os.path.exists(lib):
        subprocess.check_call([nvcc, "-shared", "-o", lib] + objs +
                              ["-Xlinker", "-Bsymbolic", "-lnccl", "-lpthread", "-cudart", "static"])
    suffix = "_" + variant if variant else ""
    for fmt in ("CRS", "SCS", "CCRS"):
        shim = os.path.join(HERE, "libsparsebench_b200_%s%s.so" % (fmt, suffix))
        src = os.path.join(CSRC, "dropin.c")
        if force or _newer([src, lib] + _headers(), shim):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-D" + fmt] + defs + ["-I" + os.path.join(ROOT, "include"), src,
                                   "-o", shim, "-L" + HERE, "-lsparsebench_b200" + suffix, "-Wl,-rpath,$ORIGIN"])
    return lib


if __name__ == "__main__":
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variants=[("" if v == "default" else v) for v in only] or None))
