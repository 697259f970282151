"""World-size>1 host logic over a real multi-process transport, on CPU (torch.distributed / gloo):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tests/gloo_partition_check.py

Every rank builds its own row block (matrixGenerate, host arrays), runs commPartition in its three collective steps
(sbPartitionLocal -> all-gather of the per-owner counts -> request lists to the owners -> sbPartitionFinish, i.e.
comm.c:414-625 with gloo in place of MPI/NCCL), and checks its lists bit for bit against the oracle; then performs
the halo exchange of comm.c:627-651 with those lists over gloo and checks the received values. No GPU needed."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import orc  # noqa: E402
from sparsebench_b200 import api  # noqa: E402


def partition_over_gloo(L, g, rank, world):
    """commPartition in its three collective steps (comm.c:414-625 with gloo in place of MPI/NCCL)."""
    # step 1: all-gather of the first rows (comm.c:496)
    starts = [None] * world
    dist.all_gather_object(starts, int(g.startRow))
    starts = np.array(starts, np.uint32)
    want = np.zeros(world, np.int32)
    plan = L.sbPartitionLocal(C.byref(g), rank, world, starts.ctypes.data, want.ctypes.data)
    # step 2: all-gather of the per-owner counts
    wants = [None] * world
    dist.all_gather_object(wants, want.tolist())
    wm = np.array(wants, np.int32)                      # [requester][owner]
    # step 3: request lists travel requester -> owner (comm.c:130-161)
    mine = {}
    for s in range(world):
        cnt = C.c_int(0)
        ptr = L.sbPartitionRequestSlice(plan, wm.ctypes.data, s, C.byref(cnt))
        assert cnt.value == wm[rank, s]
        mine[s] = [ptr[i] for i in range(cnt.value)]
    allreq = [None] * world
    dist.all_gather_object(allreq, mine)
    received = []
    for q in range(world):                              # ascending requester
        received += allreq[q][rank]
    received = np.array(received + [0], np.int32)
    comm = api.Comm()
    comm.rank, comm.size = rank, world
    L.sbPartitionFinish(plan, C.byref(comm), wm.ctypes.data, received.ctypes.data)
    return comm, starts


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = api.lib()
    bad = []
    for (nx, ny, nz, use7) in [(3, 3, 2, False), (4, 5, 4, True), (8, 8, 3, False), (1, 4, 4, False)]:
        g = api.matrixGenerate(nx, ny, nz, rank, world, use7)
        n = nx * ny * nz
        comm, _starts = partition_over_gloo(L, g, rank, world)
        # oracle: the reference algorithm for all ranks at once
        omats = [orc.generate(nx, ny, nz, r, world, use7) for r in range(world)]
        part = orc.Partition(omats)
        d, o = comm.lists(), part.ranks[rank]
        tag = "%dx%dx%d/%d" % (nx, ny, nz, int(use7))
        for f in ("externalCount", "totalSendCount"):
            if d[f] != o[f]:
                bad.append("%s %s" % (tag, f))
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend"):
            if not np.array_equal(d[f], o[f]):
                bad.append("%s %s" % (tag, f))
        if not np.array_equal(api.gmatrix_arrays(g)[1], omats[rank].col):
            bad.append("%s renumbered columns" % tag)
        if g.nc != n + comm.externalCount:
            bad.append("%s nc" % tag)
        # halo exchange with the lists (comm.c:627-651): x = global row id, so halo slot j must hold externalsReordered[j]
        x = np.zeros(n + comm.externalCount)
        x[:n] = rank * n + np.arange(n)
        out = {}
        for i, dest in enumerate(d["destinations"]):
            lo = d["sdispls"][i]
            out[int(dest)] = x[d["elementsToSend"][lo:lo + d["sendCounts"][i]]].tolist()
        allout = [None] * world
        dist.all_gather_object(allout, out)
        for i, src in enumerate(d["sources"]):
            vals = allout[int(src)][rank]
            x[n + d["rdispls"][i]:n + d["rdispls"][i] + d["recvCounts"][i]] = vals
        if not np.array_equal(x[n:], np.asarray(o["externalsReordered"], np.float64)):
            bad.append("%s halo values" % tag)
    # arrow matrix: every rank meets the last rank's column first (owners out of ascending order). The reference's
    # lists are inconsistent there; the product's must deliver x[global id] into the slot of every external.
    import matrices
    N = 12 * world + 5
    lo, rp, col, val = matrices.arrow_blocks(N, world)[rank]
    g = api.gmatrix_from_csr(rp, col.copy(), val, startRow=lo, totalNr=N)
    comm, starts = partition_over_gloo(L, g, rank, world)
    everything = [None] * world
    dist.all_gather_object(everything, (len(rp) - 1, col.tolist(), api.gmatrix_arrays(g)[1].tolist(), comm.lists()))
    if rank == 0:
        sem = matrices.check_partition_semantics(starts.astype(np.int64), [e[0] for e in everything], [e[1] for e in everything],
                                                 [e[2] for e in everything], [e[3] for e in everything])
        bad += ["arrow: " + m for m in sem]
    for b in bad:
        print("[rank %d] FAIL %s" % (rank, b), flush=True)
    print("[rank %d] gloo_partition_check: %s" % (rank, "PASS" if not bad else "FAIL"), flush=True)
    dist.destroy_process_group()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
