"""Latency of the communication kernels in isolation (run under torchrun, >= 2 GPUs):
commExchange (halo put + wait/unpack, comm.c:627-651) and the one-double all-reduce behind ddot (comm.c:653-662)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsebench_b200 import api  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    L = api.lib()
    comm = api.Comm()
    L.commInit(C.byref(comm), 0, None)
    nx = ny = 256
    nz = 8
    g = api.matrixGenerate(nx, ny, nz, rank, world, device=True)
    L.commPartition(C.byref(comm), C.byref(g))
    n = nx * ny * nz
    x = api.to_device(np.arange(n + comm.externalCount, dtype=np.float64))
    t = api.EventTimer()
    reps = 200
    for _ in range(20):
        L.commExchange(C.byref(comm), n, x.ptr)
    L.sbDeviceSynchronize()
    t.start()
    for _ in range(reps):
        L.commExchange(C.byref(comm), n, x.ptr)
    us_ex = t.stop_ms() * 1e3 / reps
    L.sbCommAllreduceDevice.argtypes = [C.POINTER(api.Comm), C.c_void_p, C.c_int, C.c_int]
    d = api.to_device(np.ones(8))
    for _ in range(20):
        L.sbCommAllreduceDevice(C.byref(comm), d.ptr, 1, api.OP_MAX)
    L.sbDeviceSynchronize()
    t.start()
    for _ in range(reps):
        L.sbCommAllreduceDevice(C.byref(comm), d.ptr, 1, api.OP_MAX)
    us_ar = t.stop_ms() * 1e3 / reps
    print("[rank %d/%d] mode %s: exchange %.2f us (%d doubles out), all-reduce %.2f us" % (
        rank, world, os.environ.get("SB_COMM", "peer"), us_ex, comm.totalSendCount, us_ar), flush=True)
    L.commFinalize(C.byref(comm))


if __name__ == "__main__":
    main()
