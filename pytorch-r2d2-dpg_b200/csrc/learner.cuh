#pragma once
#include "common.cuh"
#include "net.cuh"

namespace r2d2 {

struct Learner {
  r2d2_learner_config cfg;
  NetShape actor_sh, critic_sh;
  int rows = 0;            // T' = burn_in + learning + n_step
  int step = 0;            // completed learner iterations (learner.py:82)
  int launches_phase[3] = {0, 0, 0};
  bool actor_forward_done = false;   // learner_actor_forward already ran for the current iteration
  int launches_actor_forward = 0;
  float* arena = nullptr;
  size_t arena_floats = 0;
  // batch (filled by replay_sample or by the caller)
  float *obs = nullptr, *act = nullptr, *rew = nullptr, *term = nullptr, *states = nullptr, *uniforms = nullptr;
  long long* leaf_idx = nullptr;
  // intermediates / results
  float *act_tc = nullptr, *q = nullptr, *q_next = nullptr, *target = nullptr, *dq = nullptr, *mu = nullptr,
        *q_pi = nullptr, *dq_pi = nullptr, *dpre_actor = nullptr, *td_sq = nullptr, *priority = nullptr,
        *losses = nullptr;
  ChainWs ws_ta, ws_tc, ws_c1, ws_a1, ws_c2;
};

int learner_create(Learner** out, const r2d2_learner_config* cfg);
int learner_destroy(Learner* l);
int learner_critic_phase(Learner* l, cudaStream_t stream);
int learner_actor_forward(Learner* l, cudaStream_t stream);
int learner_actor_phase(Learner* l, float grad_scale, cudaStream_t stream);
int learner_finish_phase(Learner* l, float grad_scale, cudaStream_t stream);

}  // namespace r2d2
