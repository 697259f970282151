"""Test matrices shared by the GPU parity tests."""
import numpy as np

from oracle import orc


def irregular_spd(n, seed=7, max_extra=20):
    """Symmetric, strictly diagonally dominant, random long-range couplings, 2..2*max_extra entries per row."""
    rng = np.random.default_rng(seed)
    rowsets = [set([i]) for i in range(n)]
    for i in range(n):
        for j in rng.integers(0, n, int(rng.integers(0, max_extra))):
            rowsets[i].add(int(j)); rowsets[int(j)].add(i)
        if i + 1 < n:
            rowsets[i].add(i + 1); rowsets[i + 1].add(i)
    rp = np.zeros(n + 1, np.uint32)
    cols, vals = [], []
    for i in range(n):
        for c in sorted(rowsets[i]):
            cols.append(c)
            vals.append(float(len(rowsets[i]) + len(rowsets[c])) if c == i else -1.0 / (1 + ((i + c) % 3)))
        rp[i + 1] = len(cols)
    return orc.Csr(rp, np.array(cols, np.uint32), np.array(vals))


def arrow_blocks(N, P, bounds=None):
    """Arrow-shaped matrix (diagonal + dense last column/row + a band) split into P row blocks. Every rank except
    the last meets the LAST rank's column before any lower rank's: owners first appear out of ascending order, the
    case the reference's halo lists get wrong (comm.c:40-114 vs :148-158)."""
    if bounds is None:
        bounds = [N * r // P for r in range(P + 1)]
    blocks = []
    for r in range(P):
        lo, hi = bounds[r], bounds[r + 1]
        rp = np.zeros(hi - lo + 1, np.uint32)
        col, val = [], []
        for i in range(lo, hi):
            cs = [N - 1, i] if i != N - 1 else list(range(N - 1, -1, -1))     # last column first, then the diagonal
            for c in (i - 3, i + 3, i - 1, i + 1):
                if 0 <= c < N and c not in cs:
                    cs.append(c)
            for c in cs:
                col.append(c)
                val.append(float(2 * N) if c == i else -1.0)
            rp[i - lo + 1] = len(col)
        blocks.append((lo, rp, np.array(col, np.uint32), np.array(val)))
    return blocks


def owners_ascending(startRows, ids):
    """True when the owners of `ids` (halo slots in order) never decrease."""
    starts = np.asarray(startRows, np.int64)
    own = np.searchsorted(starts, np.asarray(ids, np.int64), side="right") - 1
    return bool(np.all(np.diff(own) >= 0))


def check_partition_semantics(startRows, nrs, global_cols, new_cols, lists):
    """Self-consistency of a partition, independent of any oracle. Per rank r: global_cols[r] are the column ids
    before commPartition, new_cols[r] after, lists[r] the Comm lists. Checks (1) local columns are shifted by
    startRow, (2) every external id maps to ONE halo slot >= nr, slots grouped by ascending owner, first-encounter
    order inside a group, (3) a halo exchange driven by the lists (comm.c:627-651) puts x[global id] into the slot
    of every external. Returns a list of failure strings."""
    bad = []
    P = len(nrs)
    slot_global = []
    for r in range(P):
        lo, nr = int(startRows[r]), int(nrs[r])
        g, c = np.asarray(global_cols[r], np.int64), np.asarray(new_cols[r], np.int64)
        local = (g >= lo) & (g < lo + nr)
        if not np.array_equal(c[local], g[local] - lo):
            bad.append("rank %d: local columns are not col - startRow" % r)
        ext_g, ext_c = g[~local], c[~local]
        nExt = lists[r]["externalCount"]
        sg = np.full(nExt, -1, np.int64)
        if len(ext_c) and (ext_c.min() < nr or ext_c.max() >= nr + nExt):
            bad.append("rank %d: halo column outside [nr, nr+externalCount)" % r)
            slot_global.append(sg)
            continue
        first = {}
        for gid, cc in zip(ext_g.tolist(), ext_c.tolist()):
            if sg[cc - nr] not in (-1, gid):
                bad.append("rank %d: halo slot %d holds two ids" % (r, cc - nr))
            sg[cc - nr] = gid
            first.setdefault(gid, len(first))
        if np.any(sg < 0) or len(first) != nExt:
            bad.append("rank %d: %d externals but %d slots used" % (r, len(first), int(np.sum(sg >= 0))))
        if not owners_ascending(startRows, sg):
            bad.append("rank %d: halo groups not in ascending owner order" % r)
        own = np.searchsorted(np.asarray(startRows, np.int64), sg, side="right") - 1
        for o in np.unique(own):
            enc = [first[int(x)] for x in sg[own == o]]
            if enc != sorted(enc):
                bad.append("rank %d: owner %d group not in first-encounter order" % (r, int(o)))
        slot_global.append(sg)
    # exchange: x = global row id
    xs = [np.concatenate([startRows[r] + np.arange(nrs[r], dtype=np.float64), np.full(lists[r]["externalCount"], -7.0)])
          for r in range(P)]
    for r in range(P):
        d = lists[r]
        for i, dest in enumerate(d["destinations"]):
            lo = int(d["sdispls"][i])
            el = d["elementsToSend"][lo:lo + int(d["sendCounts"][i])]
            if len(el) and (el.min() < 0 or el.max() >= nrs[r]):
                bad.append("rank %d: elementsToSend outside the local rows" % r)
                continue
            dd = lists[int(dest)]
            k = list(dd["sources"]).index(r)
            if int(dd["recvCounts"][k]) != len(el):
                bad.append("rank %d -> %d: send/recv count mismatch" % (r, int(dest)))
                continue
            at = int(nrs[int(dest)]) + int(dd["rdispls"][k])
            xs[int(dest)][at:at + len(el)] = xs[r][el]
    for r in range(P):
        if not np.array_equal(xs[r][int(nrs[r]):], slot_global[r].astype(np.float64)):
            bad.append("rank %d: exchanged halo values differ from the ids the columns refer to" % r)
    return bad
