"""CPU-side checks of the product: the C-ABI library loads without a GPU, exports every symbol the header
declares, and its host-only entry points (matrixGenerate, the commPartition halves) match the oracle / the
reference outputs bit for bit. No compute call needs a device here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import orc
from sparsebench_b200 import _lib, api
import matrices

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sparsebench_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = text[text.index("/* ") if "/* " in text else 0:]
    names = re.findall(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return sorted(set(names))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 40, names
    for expected in ("allocate", "getTimeStamp", "waxpby", "ddot", "commPartition", "commExchange", "commReduction",
                     "commInit", "commFinalize", "matrixGenerate", "sbCRS_spMVM", "sbSCS_convertMatrix", "sbSolveCG"):
        assert expected in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


@pytest.mark.parametrize("variant", ["f32", "u64", "f32u64"])
def test_type_variant_libraries_load_and_export_every_declared_symbol(variant):
    """util.h:35-53 as build variants: same sources, other CG_FLOAT / CG_UINT; loaded RTLD_LOCAL next to the default one"""
    L = C.CDLL(os.path.join(ROOT, "sparsebench_b200", "libsparsebench_b200_%s.so" % variant))
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, missing
    for fmt in ("CRS", "SCS", "CCRS"):
        S = C.CDLL(os.path.join(ROOT, "sparsebench_b200", "libsparsebench_b200_%s_%s.so" % (fmt, variant)))
        assert all(hasattr(S, n) for n in ("convertMatrix", "spMVM", "solveCG"))


@pytest.mark.parametrize("fmt", ["CRS", "SCS", "CCRS"])
def test_dropin_shims_export_reference_names(fmt):
    S = _lib.load_dropin(fmt)
    for n in ("convertMatrix", "spMVM", "solveCG"):       # matrix.h:57, solver.h:11-13
        assert hasattr(S, n)


def test_struct_layouts_match_reference_headers():
    # CG_UINT = unsigned int, pointers 8 bytes: CRSMatrix.h:9-16, SCSMatrix.h:13-27, matrix.h:29-35, comm.h:27-46
    assert C.sizeof(api.GMatrix) == 48 and C.sizeof(api.CRSMatrix) == 56 and C.sizeof(api.CCRSMatrix) == 48
    assert C.sizeof(api.SCSMatrix) == 32 + 16 + 24 + 32 and api.SCSMatrix.C.offset == 48
    assert C.sizeof(api.Parameter) == 32 and api.ENTRY_DTYPE.itemsize == 16
    assert api.Comm.elementsToSend.offset == 24 and api.Comm.communicator.offset == 96


@pytest.mark.parametrize("nx,ny,nz,rank,size,use7", [(4, 3, 2, 0, 1, False), (5, 4, 3, 1, 3, False), (3, 3, 3, 2, 3, True),
                                                     (1, 1, 5, 0, 2, False), (16, 16, 8, 3, 4, False)])
def test_host_matrixGenerate_bit_exact(nx, ny, nz, rank, size, use7):
    g = api.matrixGenerate(nx, ny, nz, rank, size, use7)
    rp, col, val = api.gmatrix_arrays(g)
    m = orc.generate(nx, ny, nz, rank, size, use7)
    assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val)
    n = nx * ny * nz
    assert (g.nr, g.nc, g.nnz, g.totalNr, g.totalNnz, g.startRow, g.stopRow) == \
        (n, n, 27 * n, n * size, 27 * n * size, n * rank, n * rank + n - 1)      # matrix.c:114-120
    api.lib().sbFreeGMatrix(C.byref(g))


def partition_serial(mats):
    """Drives sbPartitionLocal / sbPartitionRequestSlice / sbPartitionFinish for all ranks in one process."""
    L = api.lib()
    P = len(mats)
    starts = np.array([g.startRow for g in mats], np.uint32)
    want = np.zeros((P, P), np.int32)
    plans = []
    for r, g in enumerate(mats):
        w = np.zeros(P, np.int32)
        plans.append(L.sbPartitionLocal(C.byref(g), r, P, starts.ctypes.data, w.ctypes.data))
        want[r] = w
    comms = []
    for r in range(P):
        received = []
        for s in range(P):                     # ascending requester
            if want[s, r] > 0:
                cnt = C.c_int(0)
                ptr = L.sbPartitionRequestSlice(plans[s], want.ctypes.data, r, C.byref(cnt))
                assert cnt.value == want[s, r]
                received += [ptr[i] for i in range(cnt.value)]
        received = np.array(received + [0], np.int32)
        comms.append((r, received))
    out = []
    for r, received in comms:
        c = api.Comm()
        c.rank, c.size = r, P
        L.sbPartitionFinish(plans[r], C.byref(c), want.ctypes.data, received.ctypes.data)
        out.append(c)
    return out


MPI_CASES = [(3, 3, 3, 2, False), (2, 4, 3, 2, False), (4, 5, 4, 3, True), (8, 16, 16, 4, False), (1, 4, 4, 4, False)]


@pytest.mark.parametrize("P,nx,ny,nz,use7", MPI_CASES)
def test_partition_lists_bit_exact_vs_unmodified_comm_c(golden, P, nx, ny, nz, use7):
    """comm.c:414-625 pinned by the shim-driven reference run (tests/golden/make_golden.py)."""
    mats = [api.matrixGenerate(nx, ny, nz, r, P, use7) for r in range(P)]
    comms = partition_serial(mats)
    cfg = "mpi_P%d_%dx%dx%d_%d_" % (P, nx, ny, nz, int(use7))
    for r in range(P):
        d = comms[r].lists()
        sc = golden[cfg + "r%d_scalars" % r]
        assert [mats[r].nr, mats[r].nc, d["externalCount"], d["totalSendCount"]] == list(sc[:4])
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls", "elementsToSend"):
            assert np.array_equal(golden[cfg + "r%d_%s" % (r, f)], d[f]), (r, f)
        rp, col, _ = api.gmatrix_arrays(mats[r])
        assert np.array_equal(golden[cfg + "r%d_cols" % r], col)
        assert np.array_equal(golden[cfg + "r%d_rowPtr" % r], rp)


def _irregular_blocks(sort_cols):
    rng = np.random.default_rng(11)
    N, P = 300, 5
    rows = []
    for i in range(N):
        cols = {i, max(i - 1, 0), min(i + 1, N - 1), int(rng.integers(0, N)), int(rng.integers(0, N))}
        rows.append(sorted(cols) if sort_cols else sorted(cols, key=lambda c: (c * 7919) % N))
    bounds = [0, 50, 120, 130, 220, 300]
    blocks = []
    for r in range(P):
        lo, hi = bounds[r], bounds[r + 1]
        rp = np.zeros(hi - lo + 1, np.uint32)
        col = []
        for i in range(lo, hi):
            col += rows[i]
            rp[i - lo + 1] = len(col)
        blocks.append((lo, rp, np.array(col, np.uint32), rng.standard_normal(len(col))))
    return N, blocks


def _partition_blocks(N, blocks):
    omats = [orc.Csr(rp, col.copy(), val, startRow=lo, totalNr=N) for lo, rp, col, val in blocks]
    gmats = [api.gmatrix_from_csr(rp, col.copy(), val, startRow=lo, totalNr=N) for lo, rp, col, val in blocks]
    part = orc.Partition(omats)
    comms = partition_serial(gmats)
    starts = [lo for lo, _, _, _ in blocks]
    nrs = [len(rp) - 1 for _, rp, _, _ in blocks]
    new_cols = [api.gmatrix_arrays(g)[1] for g in gmats]
    lists = [c.lists() for c in comms]
    bad = matrices.check_partition_semantics(starts, nrs, [b[2] for b in blocks], new_cols, lists)
    assert not bad, bad
    return part, omats, starts, new_cols, lists


@pytest.mark.parametrize("sort_cols", [True, False])
def test_partition_on_irregular_matrix(sort_cols):
    """A banded + random-long-range matrix split over 5 ranks of unequal size. The lists must be self-consistent
    (tests/matrices.py:check_partition_semantics) and equal the reference's (oracle restatement of comm.c:414-625)
    on every rank on which the reference itself is self-consistent, i.e. meets its owners in ascending order."""
    N, blocks = _irregular_blocks(sort_cols)
    part, omats, starts, new_cols, lists = _partition_blocks(N, blocks)
    compared = 0
    for r in range(len(blocks)):
        o = part.ranks[r]
        assert lists[r]["externalCount"] == o["externalCount"]
        for f in ("sources", "recvCounts", "rdispls", "destinations", "sendCounts", "sdispls"):
            assert np.array_equal(lists[r][f], o[f]), (r, f)          # counts and topology never depend on the order
        if matrices.owners_ascending(starts, o["externalsReordered"]):
            assert np.array_equal(new_cols[r], omats[r].col), r
            compared += 1
        # what I send to d is defined by d's halo layout: equal to the reference's when d's layout is the reference's
        for i, d in enumerate(lists[r]["destinations"]):
            if matrices.owners_ascending(starts, part.ranks[int(d)]["externalsReordered"]):
                lo, cnt = int(lists[r]["sdispls"][i]), int(lists[r]["sendCounts"][i])
                assert np.array_equal(lists[r]["elementsToSend"][lo:lo + cnt], o["elementsToSend"][lo:lo + cnt]), (r, int(d))
    assert compared >= (1 if sort_cols else 0)


@pytest.mark.parametrize("P", [3, 4])
def test_partition_arrow_matrix_owners_out_of_order(P):
    """Arrow matrix: every rank meets the LAST rank's column first. The reference then addresses the wrong ranks
    (its elementsToSend leave the local row range); the product lays the halo groups out in ascending owner order
    and must be self-consistent."""
    N = 40
    blocks = matrices.arrow_blocks(N, P)
    part, omats, starts, new_cols, lists = _partition_blocks(N, blocks)
    # the reference's own lists are inconsistent here: some rank is asked for rows it does not own
    broken = any(len(part.ranks[r]["elementsToSend"]) and
                 (part.ranks[r]["elementsToSend"].min() < 0 or part.ranks[r]["elementsToSend"].max() >= len(blocks[r][1]) - 1)
                 for r in range(P))
    assert broken, "the arrow matrix no longer exercises the out-of-order case"


# ------------------------------------------------------------------------------------------- MatrixMarket path
def _write_mm(path, field, symm, entries, n):
    with open(path, "w") as f:
        f.write("%%%%MatrixMarket matrix coordinate %s %s\n%% comment line\n%d %d %d\n" % (field, symm, n, n, len(entries)))
        for (r, c, v) in entries:
            f.write("%d %d\n" % (r + 1, c + 1) if field == "pattern" else "%d %d %r\n" % (r + 1, c + 1, v))


def test_mm_reader_bit_exact_on_reference_fixtures(fixtures_dir):
    """MMMatrixRead + matrixConvertfromMM (matrix.c:123-269) through the C ABI vs the oracle's restatement."""
    from oracle import mmio
    for name in ["test%d.mtx" % t for t in range(11)] + ["matrix_band_klein.mtx"]:
        path = os.path.join(fixtures_dir, name)
        g = api.matrixRead(path)
        rp, col, val = api.gmatrix_arrays(g)
        m = mmio.read_mm(path)
        assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val), name
        assert (g.nr, g.nc, g.nnz, g.startRow, g.stopRow) == (m.nr, m.nr, m.nnz, 0, m.nr - 1)
        api.lib().sbFreeGMatrix(C.byref(g))


@pytest.mark.parametrize("field,symm", [("real", "general"), ("real", "symmetric"), ("integer", "symmetric"),
                                        ("pattern", "general"), ("pattern", "symmetric")])
def test_mm_reader_fields_symmetry_and_duplicates(tmp_path, field, symm):
    """mirroring of off-diagonals (matrix.c:208-212), pattern values = 1, unsorted input, duplicate entries kept in
    file order by the two stable sorts (:220-228)"""
    from oracle import mmio
    rng = np.random.default_rng(5)
    n = 23
    entries = []
    for _ in range(140):
        r, c = int(rng.integers(0, n)), int(rng.integers(0, n))
        if symm == "symmetric" and c > r:
            r, c = c, r
        v = float(rng.integers(-9, 10)) if field == "integer" else float(rng.standard_normal())
        entries.append((r, c, v))
    entries += [(i, i, 4.0) for i in range(n)] + entries[:5]          # every row non-empty; 5 duplicates
    path = str(tmp_path / "m.mtx")
    _write_mm(path, field, symm, entries, n)
    g = api.matrixRead(path)
    rp, col, val = api.gmatrix_arrays(g)
    m = mmio.read_mm(path)
    assert np.array_equal(rp, m.rowPtr) and np.array_equal(col, m.col) and np.array_equal(val, m.val)
    from oracle import ref
    if ref.available("CRS"):
        mr = ref.csr_from_gmatrix(ref.read_mm(path))                    # the reference's own reader
        assert np.array_equal(rp, mr.rowPtr) and np.array_equal(col, mr.col) and np.array_equal(val, mr.val)


# ------------------------------------------------------------------------------------------- .bmx files
def _parse_bmx(path):
    """independent reader of the layout matrixBinfile.c:38-105 writes"""
    raw = open(path, "rb").read()
    assert raw[:22] == b"# SparseBench DataFile" and raw[22:24] == b"\0\0"
    nr, nnz = np.frombuffer(raw, np.uint32, 2, 24)
    rp = np.frombuffer(raw, np.uint32, nr + 1, 32)
    rec = np.frombuffer(raw, np.dtype([("col", np.uint32), ("val", np.float32)]), nnz, 32 + 4 * (nr + 1))
    assert len(raw) == 32 + 4 * (nr + 1) + 8 * nnz
    return int(nr), int(nnz), rp, rec["col"], rec["val"]


@pytest.mark.parametrize("P", [1, 2, 3, 7])
def test_bmx_write_then_read_per_rank(tmp_path, fixtures_dir, P):
    """matrixBinWrite (single rank, main.c:42-52) and matrixBinRead for every rank of P (row blocks of sizeOfRank)"""
    L = api.lib()
    one = api.Comm()
    one.rank, one.size = 0, 1
    g = api.matrixRead(os.path.join(fixtures_dir, "matrix_band_klein.mtx"))
    rp0, col0, val0 = api.gmatrix_arrays(g)
    path = str(tmp_path / "klein.bmx")
    L.matrixBinWrite(C.byref(g), C.byref(one), path.encode())
    nr, nnz, rp, col, val = _parse_bmx(path)
    assert (nr, nnz) == (g.nr, int(rp0[-1])) and np.array_equal(rp, rp0) and np.array_equal(col, col0)
    assert np.array_equal(val, val0.astype(np.float32))
    start = 0
    for r in range(P):
        c = api.Comm()
        c.rank, c.size = r, P
        m = api.GMatrix()
        L.matrixBinRead(C.byref(m), C.byref(c), path.encode())
        m._device = False
        n = nr // P + (1 if nr % P > r else 0)
        assert (m.nr, m.nc, m.startRow, m.stopRow, m.totalNr, m.totalNnz) == (n, n, start, start + n - 1, nr, nnz)
        rpl, cl, vl = api.gmatrix_arrays(m)
        lo, hi = int(rp0[start]), int(rp0[start + n])
        assert m.nnz == hi - lo and np.array_equal(rpl, rp0[start:start + n + 1] - lo)
        assert np.array_equal(cl, col0[lo:hi]) and np.array_equal(vl, val0[lo:hi].astype(np.float32).astype(np.float64))
        L.sbFreeGMatrix(C.byref(m))
        start += n


# ------------------------------------------------------------------------------------------- .bmx pinned by the reference
@pytest.fixture(scope="module")
def ref_files():
    """outputs of the reference's own matrixBinWrite / matrixBinRead / multi-rank commDistributeMatrix (shim ranks),
    tests/golden/make_file_golden.py"""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_files.npz"))


def test_bmx_written_bytes_equal_the_reference_file(tmp_path, fixtures_dir):
    """matrixBinWrite (matrixBinfile.c:38-105): the file this library writes for klein is byte-identical to the one the
    reference's own code wrote through MPI-IO (tests/golden/reference_fixtures/klein_ref.bmx)."""
    L = api.lib()
    one = api.Comm()
    one.rank, one.size = 0, 1
    g = api.matrixRead(os.path.join(fixtures_dir, "matrix_band_klein.mtx"))
    path = str(tmp_path / "klein.bmx")
    L.matrixBinWrite(C.byref(g), C.byref(one), path.encode())
    mine, ref_bytes = open(path, "rb").read(), open(os.path.join(fixtures_dir, "klein_ref.bmx"), "rb").read()
    assert len(ref_bytes) == 2820 and mine == ref_bytes


@pytest.mark.parametrize("P", [1, 2, 3, 7])
def test_bmx_read_equals_the_reference_reader(ref_files, fixtures_dir, P):
    """matrixBinRead (matrixBinfile.c:107-236) of the reference-written file, every rank of P: header fields, local
    row pointers, columns and float32-rounded values equal what the reference's own reader returned on P shim ranks."""
    L = api.lib()
    path = os.path.join(fixtures_dir, "klein_ref.bmx")
    live = None
    from oracle import ref
    if ref.available("mpi_CRS"):
        live = ref.mpi_bmx_read(P, path)
    for r in range(P):
        c = api.Comm()
        c.rank, c.size = r, P
        m = api.GMatrix()
        L.matrixBinRead(C.byref(m), C.byref(c), path.encode())
        m._device = False
        key = "bmx_klein_P%d_r%d_" % (P, r)
        assert [m.nr, m.nc, m.nnz, m.totalNr, m.totalNnz, m.startRow, m.stopRow] == list(ref_files[key + "scalars"])
        rp, col, val = api.gmatrix_arrays(m)
        assert np.array_equal(rp, ref_files[key + "rowPtr"]) and np.array_equal(col, ref_files[key + "cols"])
        assert np.array_equal(val, ref_files[key + "vals"])
        if live is not None:
            assert np.array_equal(rp, live[r]["rowPtr"]) and np.array_equal(col, live[r]["cols"]) and np.array_equal(val, live[r]["vals"])
        L.sbFreeGMatrix(C.byref(m))


def test_single_rank_mm_path_equals_the_reference_mpi_build(ref_files, fixtures_dir):
    """MMMatrixRead + commDistributeMatrix + matrixConvertfromMM on ONE rank against the reference's -D_MPI build
    (which, unlike its non-MPI branch, fills totalNr / totalNnz: comm.c:366-368 vs :404-410)."""
    g = api.matrixRead(os.path.join(fixtures_dir, "matrix_band_klein.mtx"))
    key = "mm_matrix_band_klein_P1_r0_"
    assert [g.nr, g.nc, g.nnz, g.totalNr, g.totalNnz, g.startRow, g.stopRow] == list(ref_files[key + "scalars"])
    rp, col, val = api.gmatrix_arrays(g)
    assert np.array_equal(rp, ref_files[key + "rowPtr"]) and np.array_equal(col, ref_files[key + "cols"]) and np.array_equal(val, ref_files[key + "vals"])
