#pragma once
#include "common.cuh"

namespace r2d2 {

struct TdPriorityParams {
  const float* q = nullptr;        // [L,B,A] online critic on (o_t, a_t), t in [Bn, Bn+L)      learner.py:105
  const float* q_next = nullptr;   // [L,B,A] target critic on (o_{t+n}, target_actor(o_{t+n}))  learner.py:106
  const float* rew = nullptr;      // [T',B] rewards, already n-step pre-summed by the actor     actor.py:74-76
  const float* term = nullptr;     // [T',B] terminal flags; row t+n-1 gates the bootstrap       learner.py:107
  float* target = nullptr;         // [L,B,A] h(R + gamma^n (1-d) Q')  (optional)
  float* dq = nullptr;             // [L,B,A] d critic_loss / d q = 2 (q - y) / (L*B*A)  (optional)
  float* td_sq = nullptr;          // [L,B] mean over A of squared TD (optional)
  float* priority = nullptr;       // [B] eta*max + (1-eta)*mean over the [b:-1:B] slice (optional)
  float* loss_sum = nullptr;       // scalar: MSE-mean critic loss (zeroed by the call, optional)
  int L = 0, B = 0, A = 0, burn_in = 0, n_step = 0;
  float gamma_n = 0.f;             // gamma ** n_step
  float eta = 0.9f;
};

int td_priority(const TdPriorityParams& p, cudaStream_t stream);
// actor-side next rows (actor.py:74-107), batched over episodes: raw / out [T,B] time-major, n_rows[b] = rows of episode b
// incl. its n_step pad rows; q [T-n_step.., B, A] online critic, q_next [T,B,A] target critic on target-actor actions
int nstep_rewards(const float* raw, const int* n_rows, int T, int B, int n_step, float gamma, float* out, cudaStream_t stream);
int actor_priorities(const float* q, const float* q_next, const float* rew, const float* term, const int* n_rows, int B,
                     int A, int burn_in, int learning, int n_step, float gamma, float eta, int p_max, float* prio,
                     cudaStream_t stream);
// out[n] += sum_m x[m,n] (and out2 if given); accumulates into pre-zeroed buffers
int colsum(const float* x, long long ld, int M, int N, float* out, float* out2, cudaStream_t stream);
int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t stream);
int mul_dtanh(const float* d_out, const float* out, float* d_pre, long long n, cudaStream_t stream);
int adam_step(float* param, const float* grad, float* m, float* v, long long n, int step, float lr, float beta1,
              float beta2, float eps, float grad_scale, cudaStream_t stream);
int fill_f32(float* x, long long n, float value, cudaStream_t stream);
int scaled_sum(const float* x, long long n, float scale, float* out, cudaStream_t stream);

}  // namespace r2d2
