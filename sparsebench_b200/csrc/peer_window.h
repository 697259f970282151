// Layout of the per-rank control window that peers map through CUDA IPC (comm.cu), and the descriptor of one
// all-reduce epoch over those windows. Shared by the communication layer and by the kernels that have an
// all-reduce fused into their epilogue (push) or prologue (collect).
#pragma once
#include <stdint.h>

namespace sb {

constexpr int kMaxRanks = 64;
constexpr int kRedDepth = 4;         // epochs in flight: a rank is never more than two epochs ahead of a slow reader

struct CtrlWindow {
  unsigned long long haloFlag[kMaxRanks];              // [source rank] arrival counter of slot-based halo exchanges
  unsigned long long haloAck[kMaxRanks];               // [dest rank]   counter of exchanges that dest has copied out of its slot
  unsigned long long directFlag[kMaxRanks];            // [source rank] arrival counter of direct (registered-vector) exchanges
  unsigned long long redFlag[kRedDepth][kMaxRanks];    // [epoch % depth][rank] epoch of the value below
  double redVal[kRedDepth][kMaxRanks];
};

// One sum all-reduce of a double: every rank stores its partial into slot [rank] of every peer's window (push),
// any later kernel sums the `size` slots of its own window in rank order (collect) -- bit-identical on all ranks.
struct PeerReduce {
  int size = 0, rank = 0;            // size == 0: not in use
  unsigned long long epoch = 0;
  CtrlWindow* mine = nullptr;
  CtrlWindow* const* peers = nullptr;   // device array of `size` mapped windows (own window at [rank])
  unsigned long long* trace = nullptr;  // SB_SYNC_TRACE: [0] += ns block 0 waited in the collect, [1] += 1
};

} // namespace sb
