// Host half of commPartition (replaces comm.c:414-625 with buildIndexMapping :40-114 and
// buildElementsToSend :116-182). Pure integer work on the host, no device and no communicator: the two
// small exchanges it needs (an all-gather of per-owner counts and the request lists) are done by the
// caller -- NCCL inside commPartition (comm.cu), torch.distributed/gloo in the CPU tests.
//
// The outputs (halo numbering, renumbered column ids, elementsToSend, neighbour lists and displacements)
// are bit-identical to the reference's; its unbalanced binary search tree (bstree.c, O(ext^2) on stencil
// input) is replaced by an open-addressing hash set with the same first-encounter ordinals.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "sb_partition.h"

namespace sb {

namespace {
struct OrdinalSet {
  std::vector<idx_t> key;
  std::vector<int> ord;
  idx_t mask = 0;
  explicit OrdinalSet(size_t expect)
  {
    size_t cap = 16;
    while (cap < 2 * expect + 8) cap <<= 1;
    key.assign(cap, 0);
    ord.assign(cap, -1);
    mask = (idx_t)(cap - 1);
  }
  idx_t slot(idx_t k) const
  {
    idx_t h = (k * 2654435761u) & mask;
    while (ord[h] >= 0 && key[h] != k) h = (h + 1) & mask;
    return h;
  }
};
} // namespace

void PartitionPlan::build(const idx_t* extRefs, size_t nRefs, int rank_, int size_, idx_t nr_, idx_t startRow_,
    const idx_t* startRows)
{
  rank = rank_; size = size_; nr = nr_; startRow = startRow_;
  // step 1 (comm.c:452-473): ordinals in first-encounter order
  OrdinalSet seen(nRefs);
  extGlobal.clear();
  for (size_t i = 0; i < nRefs; i++) {
    const idx_t c = extRefs[i];
    const idx_t h = seen.slot(c);
    if (seen.ord[h] < 0) {
      seen.key[h] = c;
      seen.ord[h] = (int)extGlobal.size();
      extGlobal.push_back(c);
    }
  }
  const int nExt = (int)extGlobal.size();
  // step 2 (comm.c:496-520): owner = largest rank whose first row is <= the id
  std::vector<int> owner((size_t)nExt);
  want.assign((size_t)size, 0);
  for (int i = 0; i < nExt; i++) {
    int o = size - 1;
    while (o > 0 && startRows[o] > extGlobal[(size_t)i]) o--;
    owner[(size_t)i] = o;
    want[(size_t)o]++;
  }
  // step 3 (comm.c:40-114): halo slots grouped by owner, first-encounter order inside a group. The reference emits
  // the groups in order of first appearance but then ships and receives them as if they were in ASCENDING owner
  // order (sources / rdispls of comm.c:522-580 and the slices of :148-158 are running sums over ascending ranks):
  // whenever a higher owner is met before a lower one its lists address the wrong ranks. The groups are therefore
  // laid out in ascending owner order here -- bit-identical to the reference in every case in which the reference
  // is self-consistent (all generated stencils, row-sorted banded input), and correct in the others.
  std::vector<int> groupStart((size_t)size, 0);
  int cursor = 0;
  for (int o = 0; o < size; o++) {
    groupStart[(size_t)o] = cursor;
    cursor += want[(size_t)o];
  }
  localId.assign((size_t)nExt, 0);
  requests.assign((size_t)nExt, 0);
  std::vector<int> fill = groupStart;
  for (int i = 0; i < nExt; i++) {
    const int slot = fill[(size_t)owner[(size_t)i]]++;
    localId[(size_t)i] = (idx_t)((int)nr + slot);
    requests[(size_t)slot] = (int)extGlobal[(size_t)i];        // externalsReordered (comm.c:108-110)
  }
  // lookup table for the column rewrite
  lookupKey.swap(seen.key);
  lookupOrd.swap(seen.ord);
  lookupMask = seen.mask;
}

idx_t PartitionPlan::renumber(idx_t col, idx_t stopRow) const
{
  if (col >= startRow && col <= stopRow) return col - startRow;       // comm.c:100-101
  idx_t h = (col * 2654435761u) & lookupMask;
  while (lookupOrd[h] >= 0 && lookupKey[h] != col) h = (h + 1) & lookupMask;
  return localId[(size_t)lookupOrd[h]];                               // comm.c:102-104
}

void PartitionPlan::finish(CommLists& out, const int* wantMatrix, const int* received) const
{
  // topology (comm.c:522-580): sources = owners I need, destinations = ranks that need me, ascending;
  // displacements are running sums in that order (comm.c:135,:150)
  out.externalCount = (int)extGlobal.size();
  out.sources.clear(); out.recvCounts.clear(); out.rdispls.clear();
  out.destinations.clear(); out.sendCounts.clear(); out.sdispls.clear();
  int racc = 0, sacc = 0;
  for (int s = 0; s < size; s++) {
    const int in = wantMatrix[(size_t)rank * size + s];
    const int outc = wantMatrix[(size_t)s * size + rank];
    if (in > 0) { out.sources.push_back(s); out.recvCounts.push_back(in); out.rdispls.push_back(racc); racc += in; }
    if (outc > 0) { out.destinations.push_back(s); out.sendCounts.push_back(outc); out.sdispls.push_back(sacc); sacc += outc; }
  }
  out.totalSendCount = sacc;
  out.elementsToSend.resize((size_t)sacc);
  for (int i = 0; i < sacc; i++) {
    const long long local = (long long)received[i] - (long long)startRow;                        // comm.c:164-166
    if (local < 0 || local >= (long long)nr) {   // would become a device gather index: never let a foreign row through
      fprintf(stderr, "sparsebench_b200: commPartition: rank %d was asked for row %d, which it does not own (rows %u..%u)\n",
          rank, received[i], startRow, startRow + nr - 1);
      exit(EXIT_FAILURE);
    }
    out.elementsToSend[(size_t)i] = (int)local;
  }
}

void PartitionPlan::requestSlice(const int* wantMatrix, int source, const int** ptr, int* count) const
{
  // the slice shipped to `source` starts at my rdispls for it: a running sum over ascending sources
  // (comm.c:148-158) = the start of that owner's group (groups are laid out in ascending owner order)
  int off = 0;
  for (int s = 0; s < source; s++) off += wantMatrix[(size_t)rank * size + s];
  *ptr = requests.data() + off;
  *count = wantMatrix[(size_t)rank * size + source];
}

} // namespace sb
