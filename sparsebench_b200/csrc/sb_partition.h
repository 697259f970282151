// Host-side partition plan shared by partition.cpp (algorithm) and comm.cu (transport + device paths).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "sb_types.h"

namespace sb {

struct CommLists {
  int externalCount = 0, totalSendCount = 0;
  std::vector<int> sources, recvCounts, rdispls, destinations, sendCounts, sdispls, elementsToSend;
};

struct PartitionPlan {
  int rank = 0, size = 1;
  idx_t nr = 0, startRow = 0;
  std::vector<idx_t> extGlobal;   // external global ids, first-encounter order (comm.c:452-473)
  std::vector<idx_t> localId;     // halo slot (>= nr) of each ordinal (comm.c:57-81)
  std::vector<int> want;             // want[owner] = number of externals owned by `owner` (comm.c:496-520)
  std::vector<int> requests;         // externalsReordered: global id held by halo slot j (comm.c:108-110)
  std::vector<idx_t> lookupKey;
  std::vector<int> lookupOrd;
  idx_t lookupMask = 0;

  void build(const idx_t* extRefs, size_t nRefs, int rank, int size, idx_t nr, idx_t startRow,
      const idx_t* startRows);
  idx_t renumber(idx_t col, idx_t stopRow) const;
  void requestSlice(const int* wantMatrix, int source, const int** ptr, int* count) const;
  void finish(CommLists& out, const int* wantMatrix, const int* received) const;
};

} // namespace sb
