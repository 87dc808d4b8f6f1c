#pragma once
#include "common.cuh"
#include "net.cuh"
#include "peer.cuh"

namespace r2d2 {

struct Learner {
  r2d2_learner_config cfg;
  NetShape actor_sh, critic_sh;
  int rows = 0;            // T' = burn_in + learning + n_step
  int step = 0;            // completed learner iterations (learner.py:82)
  int launches_phase[3] = {0, 0, 0};
  // side stream for work that depends on the batch and the weights only (the input projections of the online critic
  // chain and of the actor's DPG chain): it fills the SMs the persistent scans of the target chains leave idle
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_c1_inputs = nullptr, ev_a1_inputs = nullptr;
  bool overlap_inputs = false;
  bool overlap_actor_inputs = true;  // off while the caller defers the actor's optimiser step behind the next critic phase
  bool a1_inputs_pending = false;    // a1's input projection of this iteration was issued on the side stream
  bool actor_forward_done = false;   // learner_actor_forward already ran for the current iteration
  int launches_actor_forward = 0;
  float* arena = nullptr;
  size_t arena_floats = 0;
  // batch (filled by replay_sample or by the caller).  Two slots: while the phases of iteration i still read slot s, the
  // caller may fill slot 1-s with batch i+1 and run its target chains early (learner_target_phase) - they read the
  // target nets only, so they are the independent work between "critic gradients complete" and "sums needed" of the
  // data-parallel gradient exchange.  obs .. leaf_idx below always point into slots[cur_slot].
  struct BatchSlot {
    float *obs = nullptr, *act = nullptr, *rew = nullptr, *term = nullptr, *states = nullptr, *uniforms = nullptr;
    long long* leaf_idx = nullptr;
  } slots[2];
  int cur_slot = 0;
  int targets_slot = -1;        // slot whose target-chain outputs (q_next) are current; -1: none
  int c1_inputs_slot = -1;      // slot whose online-critic input projection is in flight / done on the side stream
  int launches_target = 0;      // launches of the last learner_target_phase
  bool target_phase_standalone = false;
  float *obs = nullptr, *act = nullptr, *rew = nullptr, *term = nullptr, *states = nullptr, *uniforms = nullptr;
  long long* leaf_idx = nullptr;
  // intermediates / results
  float *act_tc = nullptr, *q = nullptr, *q_next = nullptr, *target = nullptr, *dq = nullptr, *mu = nullptr,
        *q_pi = nullptr, *dq_pi = nullptr, *dpre_actor = nullptr, *td_sq = nullptr, *priority = nullptr,
        *losses = nullptr;
  ChainWs ws_ta, ws_tc, ws_c1, ws_a1, ws_c2;
  // data-parallel learner: gradient blocks live in a peer-mapped buffer and are summed by peer.cu's kernels in this
  // learner's own stream (null: single GPU, or the caller reduces cfg.*_grads itself between the phases)
  PeerExchange* peer = nullptr;
  const float* optimiser_grads(int block) const {
    return peer ? peer->sums(block) : (block == kPeerCritic ? cfg.critic_grads : cfg.actor_grads);
  }
};

int learner_create(Learner** out, const r2d2_learner_config* cfg);
int learner_destroy(Learner* l);
int learner_select_batch(Learner* l, int slot);
// target chains of the batch in `slot` (learner.py:87,94-95,106): q_next for the next learner_critic_phase on that slot.
// Reads the target nets and the batch only.  learner_critic_phase runs it itself when the current slot has none.
int learner_target_phase(Learner* l, int slot, cudaStream_t stream);
// forget a target phase that ran ahead (the caller is about to overwrite that batch)
int learner_discard_prefetch(Learner* l, cudaStream_t stream);
int learner_critic_phase(Learner* l, cudaStream_t stream);
int learner_actor_forward(Learner* l, cudaStream_t stream);
int learner_actor_phase(Learner* l, float grad_scale, cudaStream_t stream);
int learner_finish_phase(Learner* l, float grad_scale, cudaStream_t stream);
// peer_bases[k] = rank k's exchange buffer (peer_layout(...).bytes, zeroed, mapped into this process); moves the
// learner's gradient blocks into peer_bases[rank]
int learner_attach_peers(Learner* l, int rank, int world, void* const* peer_bases);

}  // namespace r2d2
