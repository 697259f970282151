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
SOURCES = ["runtime.cu", "generate.cu", "formats.cu", "spmv.cu", "vecops.cu", "cg.cu", "comm.cu", "partition.cpp", "mmio.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC,-fvisibility=default",
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def _headers():
    hs = [os.path.join(ROOT, "include", "sparsebench_b200.h")]
    hs += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(o)
        if force or _newer([s] + _headers(), o):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or not os.path.exists(LIB):
        subprocess.check_call([nvcc, "-shared", "-o", LIB] + objs +
                              ["-Xlinker", "-Bsymbolic", "-lnccl", "-lpthread", "-cudart", "static"])
    for fmt in ("CRS", "SCS", "CCRS"):
        shim = os.path.join(HERE, "libsparsebench_b200_%s.so" % fmt)
        src = os.path.join(CSRC, "dropin.c")
        if force or _newer([src, LIB] + _headers(), shim):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-D" + fmt, "-I" + os.path.join(ROOT, "include"), src,
                                   "-o", shim, "-L" + HERE, "-lsparsebench_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
