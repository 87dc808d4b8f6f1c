// Learner iteration engine: learner.py:84-139 (sample excluded - see replay.cu) as three phases so a
// data-parallel caller can all-reduce the flat gradient buffers between them (NCCL lives in the host code).
//
//   phase 1  target_actor chain  -> target_critic chain -> online critic chain -> fused TD/priority
//            kernel -> critic BPTT                                          (learner.py:86-113)
//   phase 2  critic Adam; actor chain from the zero state with two cell steps per row; critic on the
//            actor's actions with the post-step weights; dgrad through the critic; actor BPTT
//                                                                             (learner.py:114-127)
//   phase 3  actor Adam; hard target update every `target_update_interval` steps (learner.py:128-132)
//
// The actor burn-in of learner.py:92 is skipped: its state is discarded at learner.py:117 before any use.
#include "learner.cuh"

#include <cmath>
#include <stdlib.h>

#include "elementwise.cuh"
#include "gemm.cuh"
#include "lstm_scan.cuh"

namespace r2d2 {

static size_t align64(size_t n) { return (n + 63) & ~(size_t)63; }

int learner_create(Learner** out, const r2d2_learner_config* cfg) {
  R2D2_REQUIRE(out && cfg, "null");
  R2D2_REQUIRE(cfg->obs_size > 0 && cfg->n_actions > 0 && cfg->hidden > 0 && cfg->hidden % 4 == 0, "sizes");
  R2D2_REQUIRE(cfg->batch > 0 && cfg->burn_in >= 0 && cfg->learning > 0 && cfg->n_step > 0, "window");
  R2D2_REQUIRE(cfg->actor_params && cfg->critic_params && cfg->target_actor_params && cfg->target_critic_params,
               "parameter buffers");
  R2D2_REQUIRE(cfg->actor_grads && cfg->critic_grads && cfg->actor_exp_avg && cfg->actor_exp_avg_sq &&
                   cfg->critic_exp_avg && cfg->critic_exp_avg_sq, "gradient / Adam buffers");
  Learner* l = new Learner();
  l->cfg = *cfg;
  l->actor_sh = NetShape{cfg->obs_size, cfg->n_actions, cfg->hidden, false};
  l->critic_sh = NetShape{cfg->obs_size, cfg->n_actions, cfg->hidden, true};
  const int B = cfg->batch, Bn = cfg->burn_in, L = cfg->learning, n = cfg->n_step;
  const int O = cfg->obs_size, A = cfg->n_actions, H = cfg->hidden;
  const int Tt = Bn + n + L, Tc = Bn + L, Tw = Bn + L + n;
  l->rows = Tw;
  // one allocation, carved
  size_t total = 0;
  const size_t n_obs = align64((size_t)Tw * B * O), n_act = align64((size_t)Tw * B * A), n_rt = align64((size_t)Tw * B);
  const size_t n_states = align64((size_t)8 * B * H), n_lba = align64((size_t)L * B * A);
  const size_t n_acttc = align64((size_t)Tt * B * A);
  total += 2 * (n_obs + n_act + 2 * n_rt + n_states + align64(B) /*uniforms*/ + align64(2 * (size_t)B) /*leaf idx (int64)*/);
  total += n_acttc + 8 * n_lba + align64((size_t)L * B) + align64(B) + 64;
  const size_t ws_ta = ChainWs::floats(l->actor_sh, Tt, B, 1), ws_tc = ChainWs::floats(l->critic_sh, Tt, B, 1);
  const size_t ws_c1 = ChainWs::floats(l->critic_sh, Tc, B, 1), ws_a1 = ChainWs::floats(l->actor_sh, L, B, 2);
  const size_t ws_c2 = ChainWs::floats(l->critic_sh, L, B, 1);
  total += ws_ta + ws_tc + ws_c1 + ws_a1 + ws_c2;
  R2D2_CUDA_TRY(cudaMalloc(&l->arena, total * sizeof(float)));
  R2D2_CUDA_TRY(cudaMemset(l->arena, 0, total * sizeof(float)));
  l->arena_floats = total;
  float* p = l->arena;
  auto take = [&](size_t nfl) { float* r = p; p += align64(nfl); return r; };
  for (auto& b : l->slots) {
    b.obs = take(n_obs); b.act = take(n_act); b.rew = take(n_rt); b.term = take(n_rt);
    b.states = take(n_states); b.uniforms = take(B);
    b.leaf_idx = reinterpret_cast<long long*>(take(2 * (size_t)B));
  }
  learner_select_batch(l, 0);
  l->act_tc = take(n_acttc);
  l->q = take(n_lba); l->q_next = take(n_lba); l->target = take(n_lba); l->dq = take(n_lba);
  l->mu = take(n_lba); l->q_pi = take(n_lba); l->dq_pi = take(n_lba); l->dpre_actor = take(n_lba);
  l->td_sq = take((size_t)L * B); l->priority = take(B); l->losses = take(64);
  l->ws_ta = ChainWs::carve(take(ws_ta), l->actor_sh, Tt, B, 1);
  l->ws_tc = ChainWs::carve(take(ws_tc), l->critic_sh, Tt, B, 1);
  l->ws_c1 = ChainWs::carve(take(ws_c1), l->critic_sh, Tc, B, 1);
  l->ws_a1 = ChainWs::carve(take(ws_a1), l->actor_sh, L, B, 2);
  l->ws_c2 = ChainWs::carve(take(ws_c2), l->critic_sh, L, B, 1);
  l->ws_ta.inference_only = true;  // target nets: no BPTT, the scan keeps only what the heads read (learner.py:94-95,106)
  l->ws_tc.inference_only = true;
  l->ws_c1.keep_z1_image = true;   // chains whose weights get gradients: the l1 kernel also leaves z1 as the B operand of dW_ih
  l->ws_a1.keep_z1_image = true;
  {   // R2D2_OVERLAP_INPUTS=0 disables the side stream (A/B)
    const char* e = getenv("R2D2_OVERLAP_INPUTS");
    l->overlap_inputs = !(e && e[0] == '0');
    if (l->overlap_inputs) {
      int lo = 0, hi = 0;
      R2D2_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      R2D2_CUDA_TRY(cudaStreamCreateWithPriority(&l->side, cudaStreamNonBlocking, lo));   // lowest priority: leftovers only
      R2D2_CUDA_TRY(cudaEventCreateWithFlags(&l->ev_fork, cudaEventDisableTiming));
      R2D2_CUDA_TRY(cudaEventCreateWithFlags(&l->ev_c1_inputs, cudaEventDisableTiming));
      R2D2_CUDA_TRY(cudaEventCreateWithFlags(&l->ev_a1_inputs, cudaEventDisableTiming));
    }
  }
  *out = l;
  return R2D2_OK;
}

int learner_destroy(Learner* l) {
  if (!l) return R2D2_OK;
  if (l->side) { cudaStreamSynchronize(l->side); cudaStreamDestroy(l->side); }
  if (l->ev_fork) cudaEventDestroy(l->ev_fork);
  if (l->ev_c1_inputs) cudaEventDestroy(l->ev_c1_inputs);
  if (l->ev_a1_inputs) cudaEventDestroy(l->ev_a1_inputs);
  cudaFree(l->arena);
  delete l->peer;
  delete l;
  return R2D2_OK;
}

int learner_select_batch(Learner* l, int slot) {
  R2D2_REQUIRE(l && (slot == 0 || slot == 1), "batch slot");
  const Learner::BatchSlot& b = l->slots[slot];
  l->cur_slot = slot;
  l->obs = b.obs; l->act = b.act; l->rew = b.rew; l->term = b.term; l->states = b.states; l->uniforms = b.uniforms;
  l->leaf_idx = b.leaf_idx;
  return R2D2_OK;
}

int learner_discard_prefetch(Learner* l, cudaStream_t st) {
  R2D2_REQUIRE(l, "null");
  // whatever the side stream still writes into the online critic's workspace for the dropped batch goes first
  if (l->c1_inputs_slot >= 0 && l->side) R2D2_CUDA_TRY(cudaStreamWaitEvent(st, l->ev_c1_inputs, 0));
  l->c1_inputs_slot = -1;
  l->targets_slot = -1;
  return R2D2_OK;
}

int learner_target_phase(Learner* l, int slot, cudaStream_t st) {
  R2D2_REQUIRE(l && (slot == 0 || slot == 1), "batch slot");
  const r2d2_learner_config& c = l->cfg;
  const int B = c.batch, Bn = c.burn_in, L = c.learning, n = c.n_step, A = c.n_actions, H = c.hidden;
  const int Tt = Bn + n + L, Tc = Bn + L;
  const long long launches0 = launch_count();
  const Learner::BatchSlot& b = l->slots[slot];
  const bool ahead = slot != l->cur_slot;   // batch i+1 while the phases of iteration i are still running on the other slot
  const NetParams Pa_t = NetParams::from_flat(c.target_actor_params, l->actor_sh);
  const NetParams Pc_t = NetParams::from_flat(c.target_critic_params, l->critic_sh);
  const size_t BH = (size_t)B * H;
  const float* st_ta = b.states + 2 * BH;   // states[1] = target_actor (hx, cx)
  const float* st_tc = b.states + 6 * BH;   // states[3] = target_critic

  // target actor over rows [0, Bn+n+L) from its stored state (learner.py:87,94,106); actions for the last L rows
  R2D2_TRY(net_forward_inputs(l->actor_sh, Pa_t, l->ws_ta, b.obs, nullptr, Tt, B, st));
  if (l->overlap_inputs) {
    // fork: input projections that need the batch and weights nothing in this phase changes are issued on a
    // low-priority side stream right where the first persistent scan starts: the scans occupy 7 x 16 of the 148 SMs
    // and the projections take the rest.  Same slot: the online critic chain of this batch (stored actions).  The
    // actor's DPG chain of the iteration in flight: when its weights are final - always when running ahead (the caller
    // completes the previous optimiser step first), else unless the caller defers that step (overlap_actor_inputs).
    bool forked = false;
    auto fork = [&]() -> int {
      if (!forked) {
        R2D2_CUDA_TRY(cudaEventRecord(l->ev_fork, st));
        R2D2_CUDA_TRY(cudaStreamWaitEvent(l->side, l->ev_fork, 0));
        forked = true;
      }
      return R2D2_OK;
    };
    if (!ahead && l->c1_inputs_slot != slot) {
      const NetParams Pc = NetParams::from_flat(c.critic_params, l->critic_sh);
      R2D2_TRY(fork());
      R2D2_TRY(net_forward_inputs(l->critic_sh, Pc, l->ws_c1, b.obs, b.act, Tc, B, l->side));
      R2D2_CUDA_TRY(cudaEventRecord(l->ev_c1_inputs, l->side));
      l->c1_inputs_slot = slot;
    }
    if ((ahead || l->overlap_actor_inputs) && !l->a1_inputs_pending && !l->actor_forward_done) {
      const NetParams Pa = NetParams::from_flat(c.actor_params, l->actor_sh);
      R2D2_TRY(fork());
      R2D2_TRY(net_forward_inputs(l->actor_sh, Pa, l->ws_a1, l->obs + (size_t)Bn * B * c.obs_size, nullptr, L, B, l->side));
      R2D2_CUDA_TRY(cudaEventRecord(l->ev_a1_inputs, l->side));
      l->a1_inputs_pending = true;
    }
  }
  R2D2_TRY(net_forward_scan(l->actor_sh, Pa_t, l->ws_ta, st_ta, st_ta + BH, Tt, B, 1, st));
  // data parallel: this rank's slice sums of whichever gradient block was signalled last - the peers signalled one
  // input projection and one scan ago (peer.cuh); both calls are no-ops without a pending signal
  if (l->peer) {
    R2D2_TRY(peer_reduce(*l->peer, kPeerCritic, st));
    R2D2_TRY(peer_reduce(*l->peer, kPeerActor, st));
  }
  R2D2_CUDA_TRY(cudaMemcpyAsync(l->act_tc, b.act, sizeof(float) * (size_t)(Bn + n) * B * A, cudaMemcpyDeviceToDevice, st));
  R2D2_TRY(net_head_forward(l->actor_sh, Pa_t, l->ws_ta, Bn + n, Tt, B, 1, l->act_tc + (size_t)(Bn + n) * B * A, A, st));
  // target critic: stored actions while burning in, target-actor actions afterwards (learner.py:95,106)
  R2D2_TRY(net_forward(l->critic_sh, Pc_t, l->ws_tc, b.obs, l->act_tc, st_tc, st_tc + BH, Tt, B, 1, st));
  R2D2_TRY(net_head_forward(l->critic_sh, Pc_t, l->ws_tc, Bn + n, Tt, B, 1, l->q_next, A, st));
  l->targets_slot = slot;
  l->launches_target = (int)(launch_count() - launches0);
  return R2D2_OK;
}

int learner_critic_phase(Learner* l, cudaStream_t st) {
  const r2d2_learner_config& c = l->cfg;
  const int B = c.batch, Bn = c.burn_in, L = c.learning, n = c.n_step, A = c.n_actions, H = c.hidden;
  const int Tc = Bn + L;
  const long long launches0 = launch_count();
  const NetParams Pc = NetParams::from_flat(c.critic_params, l->critic_sh);
  const NetParams Gc = NetParams::from_flat(c.critic_grads, l->critic_sh);
  const size_t BH = (size_t)B * H;
  const float* st_c = l->states + 4 * BH;    // states[2] = critic

  l->target_phase_standalone = l->targets_slot == l->cur_slot;
  if (!l->target_phase_standalone) R2D2_TRY(learner_target_phase(l, l->cur_slot, st));
  // online critic over rows [0, Bn+L) with stored actions (learner.py:93,105); burn-in stays on the tape (Q4)
  if (l->c1_inputs_slot == l->cur_slot) R2D2_CUDA_TRY(cudaStreamWaitEvent(st, l->ev_c1_inputs, 0));
  else R2D2_TRY(net_forward_inputs(l->critic_sh, Pc, l->ws_c1, l->obs, l->act, Tc, B, st));
  l->c1_inputs_slot = -1;
  R2D2_TRY(net_forward_scan(l->critic_sh, Pc, l->ws_c1, st_c, st_c + BH, Tc, B, 1, st));
  if (l->peer) R2D2_TRY(peer_reduce(*l->peer, kPeerActor, st));   // no target phase in this call: first scan is this one
  R2D2_TRY(net_head_forward(l->critic_sh, Pc, l->ws_c1, Bn, Tc, B, 1, l->q, A, st));
  l->targets_slot = -1;   // q_next is consumed by the TD kernel below

  TdPriorityParams tp;
  tp.q = l->q; tp.q_next = l->q_next; tp.rew = l->rew; tp.term = l->term;
  tp.target = l->target; tp.dq = l->dq; tp.td_sq = l->td_sq; tp.priority = l->priority; tp.loss_sum = l->losses;
  tp.L = L; tp.B = B; tp.A = A; tp.burn_in = Bn; tp.n_step = n;
  tp.gamma_n = (float)std::pow((double)c.gamma, (double)n);
  tp.eta = c.eta;
  R2D2_TRY(td_priority(tp, st));

  R2D2_CUDA_TRY(cudaMemsetAsync(c.critic_grads, 0, sizeof(float) * l->critic_sh.param_count(), st));
  R2D2_TRY(net_backward(l->critic_sh, Pc, &Gc, l->ws_c1, l->obs, l->act, l->dq, Bn, Tc, B, 1, nullptr, nullptr, st));
  if (l->peer) R2D2_TRY(peer_signal(*l->peer, kPeerCritic, st));
  l->launches_phase[0] = (int)(launch_count() - launches0);
  return R2D2_OK;
}

// The actor's forward chain of the DPG update (learner.py:117,120-123 without the critic call): it reads the actor's
// weights and the observations only - NOT the critic - so a data-parallel caller runs it while the all-reduce of the
// critic gradients is in flight (r2d2_learner_actor_forward), before the critic's optimiser step.
int learner_actor_forward(Learner* l, cudaStream_t st) {
  const r2d2_learner_config& c = l->cfg;
  const int B = c.batch, Bn = c.burn_in, L = c.learning, A = c.n_actions, O = c.obs_size;
  const long long launches0 = launch_count();
  const NetParams Pa = NetParams::from_flat(c.actor_params, l->actor_sh);
  const float* obs_l = l->obs + (size_t)Bn * B * O;  // rows [Bn, Bn+L)
  // actor from the zero state, LSTM stepped twice per row (learner.py:117,122-123); mu = output of the 2nd call
  if (l->a1_inputs_pending) {   // issued on the side stream during the critic phase of this iteration
    R2D2_CUDA_TRY(cudaStreamWaitEvent(st, l->ev_a1_inputs, 0));
    l->a1_inputs_pending = false;
  } else {
    R2D2_TRY(net_forward_inputs(l->actor_sh, Pa, l->ws_a1, obs_l, nullptr, L, B, st));
  }
  // data parallel: sum this rank's slice of the critic gradients between the input projection and the scan - the
  // peers signalled one projection ago, the sums are needed one scan later (learner_actor_phase)
  if (l->peer) R2D2_TRY(peer_reduce(*l->peer, kPeerCritic, st));
  R2D2_TRY(net_forward_scan(l->actor_sh, Pa, l->ws_a1, nullptr, nullptr, L, B, 2, st));
  R2D2_TRY(net_head_forward(l->actor_sh, Pa, l->ws_a1, 0, L, B, 2, l->mu, A, st));
  l->actor_forward_done = true;
  l->launches_actor_forward = (int)(launch_count() - launches0);
  return R2D2_OK;
}

int learner_actor_phase(Learner* l, float grad_scale, cudaStream_t st) {
  const r2d2_learner_config& c = l->cfg;
  const int B = c.batch, Bn = c.burn_in, L = c.learning, A = c.n_actions, O = c.obs_size;
  const long long launches0 = launch_count();
  const NetParams Pa = NetParams::from_flat(c.actor_params, l->actor_sh);
  const NetParams Ga = NetParams::from_flat(c.actor_grads, l->actor_sh);
  const NetParams Pc = NetParams::from_flat(c.critic_params, l->critic_sh);
  const long long LBA = (long long)L * B * A;

  int extra = 0;
  if (!l->actor_forward_done) R2D2_TRY(learner_actor_forward(l, st));   // single-GPU order: same kernels, same results
  else extra = l->launches_actor_forward;
  l->actor_forward_done = false;
  const long long launches1 = launch_count();
  (void)launches1;
  if (l->peer) R2D2_TRY(peer_wait(*l->peer, kPeerCritic, st));
  R2D2_TRY(adam_step(c.critic_params, l->optimiser_grads(kPeerCritic), c.critic_exp_avg, c.critic_exp_avg_sq,
                     (long long)l->critic_sh.param_count(), l->step + 1, c.critic_lr, 0.9f, 0.999f, 1e-8f,
                     grad_scale, st));                                                     // learner.py:114
  {
    // the other slot already holds the next batch (its target chains ran ahead): the input projection of ITS online
    // critic chain needs the weights Adam just wrote and nothing else - side stream, under the scans of this phase
    const int other = 1 - l->cur_slot;
    if (l->overlap_inputs && l->targets_slot == other && l->c1_inputs_slot != other) {
      const Learner::BatchSlot& nb = l->slots[other];
      R2D2_CUDA_TRY(cudaEventRecord(l->ev_fork, st));
      R2D2_CUDA_TRY(cudaStreamWaitEvent(l->side, l->ev_fork, 0));
      R2D2_TRY(net_forward_inputs(l->critic_sh, Pc, l->ws_c1, nb.obs, nb.act, Bn + L, B, l->side));
      R2D2_CUDA_TRY(cudaEventRecord(l->ev_c1_inputs, l->side));
      l->c1_inputs_slot = other;
    }
  }

  const float* obs_l = l->obs + (size_t)Bn * B * O;  // rows [Bn, Bn+L)
  // critic (post-Adam weights, zero state) on the actor's actions; loss = mean(-Q) (learner.py:118,123-124)
  R2D2_TRY(net_forward(l->critic_sh, Pc, l->ws_c2, obs_l, l->mu, nullptr, nullptr, L, B, 1, st));
  R2D2_TRY(net_head_forward(l->critic_sh, Pc, l->ws_c2, 0, L, B, 1, l->q_pi, A, st));
  R2D2_TRY(scaled_sum(l->q_pi, LBA, -1.0f / (float)LBA, l->losses + 1, st));
  R2D2_TRY(fill_f32(l->dq_pi, LBA, -1.0f / (float)LBA, st));
  // dgrad only through the critic (its weight grads are wasted work in the reference); d_pre(actor) = dQ/da * (1-mu^2)
  R2D2_TRY(net_backward(l->critic_sh, Pc, nullptr, l->ws_c2, obs_l, l->mu, l->dq_pi, 0, L, B, 1, l->dpre_actor,
                        l->mu, st));
  R2D2_CUDA_TRY(cudaMemsetAsync(c.actor_grads, 0, sizeof(float) * l->actor_sh.param_count(), st));
  R2D2_TRY(net_backward(l->actor_sh, Pa, &Ga, l->ws_a1, obs_l, nullptr, l->dpre_actor, 0, L, B, 2, nullptr, nullptr, st));
  if (l->peer) R2D2_TRY(peer_signal(*l->peer, kPeerActor, st));
  l->launches_phase[1] = (int)(launch_count() - launches0) + extra;
  return R2D2_OK;
}

int learner_finish_phase(Learner* l, float grad_scale, cudaStream_t st) {
  const r2d2_learner_config& c = l->cfg;
  const long long launches0 = launch_count();
  if (l->peer) R2D2_TRY(peer_wait(*l->peer, kPeerActor, st));   // runs the slice reduction first if no critic phase did
  R2D2_TRY(adam_step(c.actor_params, l->optimiser_grads(kPeerActor), c.actor_exp_avg, c.actor_exp_avg_sq,
                     (long long)l->actor_sh.param_count(), l->step + 1, c.actor_lr, 0.9f, 0.999f, 1e-8f, grad_scale,
                     st));                                                                 // learner.py:128
  l->step += 1;
  if (c.target_update_interval > 0 && l->step % c.target_update_interval == 0) {           // learner.py:131-132
    R2D2_CUDA_TRY(cudaMemcpyAsync(c.target_actor_params, c.actor_params, sizeof(float) * l->actor_sh.param_count(),
                                  cudaMemcpyDeviceToDevice, st));
    R2D2_CUDA_TRY(cudaMemcpyAsync(c.target_critic_params, c.critic_params, sizeof(float) * l->critic_sh.param_count(),
                                  cudaMemcpyDeviceToDevice, st));
  }
  l->launches_phase[2] = (int)(launch_count() - launches0);
  return R2D2_OK;
}

int learner_attach_peers(Learner* l, int rank, int world, void* const* peer_bases) {
  R2D2_REQUIRE(l && peer_bases, "null argument");
  R2D2_REQUIRE(world >= 2 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "peer rank / world");
  R2D2_REQUIRE(!l->peer, "peers already attached");
  PeerExchange* x = new PeerExchange();
  x->rank = rank;
  x->world = world;
  x->lay = peer_layout((long long)l->critic_sh.param_count(), (long long)l->actor_sh.param_count(), world);
  for (int k = 0; k < world; ++k) {
    if (!peer_bases[k]) { delete x; R2D2_REQUIRE(false, "null peer buffer"); }
    x->ptrs.base[k] = static_cast<char*>(peer_bases[k]);
  }
  l->peer = x;
  l->cfg.critic_grads = x->grads(kPeerCritic);
  l->cfg.actor_grads = x->grads(kPeerActor);
  return R2D2_OK;
}

}  // namespace r2d2
