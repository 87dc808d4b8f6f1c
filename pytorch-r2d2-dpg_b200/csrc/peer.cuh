#pragma once
#include "common.cuh"

namespace r2d2 {

// Gradient exchange of the data-parallel learner over NVLink peer memory: the two flat gradient blocks
// (learner.py:113-114 critic, 127-128 actor) are the only data that crosses GPUs.  Every rank owns one symmetric buffer
// (same layout on every rank, mapped into every process):
//   [ flags 4 KB | critic grads | actor grads | critic sums | actor sums ]     (blocks padded to 4*world floats)
// and three kernels per block and iteration run IN the learner's stream:
//   signal  after the BPTT that produced the block: "my gradients are complete" -> flag word on every peer
//   reduce  rank r sums slice r of the block over all ranks (peer loads, fixed rank order) and stores the sum into the
//           `sums` block of EVERY rank (peer stores), then raises "slice r delivered" on every peer
//   wait    before the optimiser kernel reads `sums`
// Only the owner adds a slice, so all ranks see bit-identical sums.  The three points sit at different places of the
// iteration (learner.cu) with independent work in between: a rank only waits when a peer is later than that slack.
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerCritic = 0, kPeerActor = 1;

struct PeerLayout {
  size_t bytes = 0, off_flags = 0, off_grads[2] = {0, 0}, off_sums[2] = {0, 0};
  long long padded[2] = {0, 0};   // floats per block, multiple of 4 * world
};
PeerLayout peer_layout(long long n_critic, long long n_actor, int world);

struct PeerPtrs { char* base[kPeerMaxWorld]; };

struct PeerExchange {
  int rank = 0, world = 1;
  PeerPtrs ptrs{};
  PeerLayout lay;
  unsigned epoch[2] = {0, 0};          // signals sent so far per block
  bool reduce_pending[2] = {false, false};
  bool wait_pending[2] = {false, false};
  float* grads(int block) const { return reinterpret_cast<float*>(ptrs.base[rank] + lay.off_grads[block]); }
  float* sums(int block) const { return reinterpret_cast<float*>(ptrs.base[rank] + lay.off_sums[block]); }
};

int peer_signal(PeerExchange& x, int block, cudaStream_t stream);
int peer_reduce(PeerExchange& x, int block, cudaStream_t stream);   // no-op unless a signal of this block is pending
int peer_wait(PeerExchange& x, int block, cudaStream_t stream);     // runs a pending reduce first
// 0 = fine; 1 = a bounded wait for a peer's flag expired (the results of that iteration are garbage)
int peer_status(const PeerExchange& x, int* out, cudaStream_t stream);
// diagnostics, ns summed since the last reset: [0..1] slice-sum kernel waiting for the peers' signal (critic, actor),
// [2..3] slice-sum kernel in total, [4..5] wait kernel
int peer_counters(const PeerExchange& x, unsigned long long* out6, int reset, cudaStream_t stream);

}  // namespace r2d2
