// HBM-bound pieces of the learner path: fused n-step target / value rescaling / TD loss gradient /
// sequence priority (learner.py:107-111,135-138; utils.py:17-21), Adam (learner.py:50-53,114,128),
// bias-gradient column sums, small reductions.
#include "elementwise.cuh"

#include <cmath>

namespace r2d2 {
namespace {

__device__ __forceinline__ float value_rescale(float x) {
  // h(x) = sign(x) * (sqrt(|x| + 1) - 1), utils.py:20-21 (no eps*x term, no inverse anywhere)
  const float m = sqrtf(fabsf(x) + 1.0f) - 1.0f;
  return x > 0.f ? m : (x < 0.f ? -m : 0.f);
}

// fallback (no td_sq buffer to reduce through): grid ceil(B/32) CTAs, 1024 threads = 32 warps; lane -> batch column, warp -> time rows i = w, w+32, ...
// (the kernel moves ~1 MB: it is bound by the length of the per-thread dependent load chain, hence the wide block)
constexpr int TD_WARPS = 32;
__global__ void __launch_bounds__(TD_WARPS * 32) td_priority_column_kernel(TdPriorityParams p) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int L = p.L, B = p.B, A = p.A;
  const float inv_a = 1.0f / (float)A;
  const float grad_scale = 2.0f / ((float)L * (float)B * (float)A);
  float run_max = -INFINITY, run_sum = 0.f, sq_total = 0.f;
  if (b < B) {
    for (int i = w; i < L; i += TD_WARPS) {
      const float r = __ldg(p.rew + (size_t)(p.burn_in + i) * B + b);
      const float d = __ldg(p.term + (size_t)(p.burn_in + i + p.n_step - 1) * B + b);
      const float cont = p.gamma_n * (1.0f - d);
      const size_t base = ((size_t)i * B + b) * A;
      float sq = 0.f;
      for (int a = 0; a < A; ++a) {
        const float q = __ldg(p.q + base + a);
        const float y = value_rescale(r + cont * __ldg(p.q_next + base + a));
        const float diff = q - y;
        if (p.target) p.target[base + a] = y;
        if (p.dq) p.dq[base + a] = grad_scale * diff;
        sq += diff * diff;
      }
      sq_total += sq;
      const float td = sq * inv_a;
      if (p.td_sq) p.td_sq[(size_t)i * B + b] = td;
      // learner.py:137 `average_td_loss[b:-1:B]` drops flat index L*B-1, i.e. (i=L-1, b=B-1)
      if (!(i == L - 1 && b == B - 1)) { run_max = fmaxf(run_max, td); run_sum += td; }
    }
  }
  __shared__ float s_max[TD_WARPS][32], s_sum[TD_WARPS][32], s_sq[TD_WARPS];
  s_max[w][lane] = run_max;
  s_sum[w][lane] = run_sum;
  const float wsq = warp_sum(sq_total);
  if (lane == 0) s_sq[w] = wsq;
  __syncthreads();
  if (w == 0) {
    float mx = s_max[0][lane], sm = s_sum[0][lane];
#pragma unroll
    for (int k = 1; k < TD_WARPS; ++k) { mx = fmaxf(mx, s_max[k][lane]); sm += s_sum[k][lane]; }
    if (b < B && p.priority) {
      const int count = L - ((b == B - 1) ? 1 : 0);
      p.priority[b] = p.eta * mx + (1.0f - p.eta) * (sm / (float)count);  // utils.py:17-18
    }
    if (lane == 0 && p.loss_sum) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < TD_WARPS; ++k) tot += s_sq[k];
      atomicAdd(p.loss_sum, tot / ((float)L * (float)B * (float)A));
    }
  }
}

// ---- the path's TD kernel pair (learner.py:107-111,135-138): every global access is a lane-contiguous 128-byte line.
// Pass 1, one warp per (time row i, 32 consecutive batch elements): the [32 x A] spans of q and q_next are ONE contiguous
// run of 32*A floats each (layout [L][B][A]); the warp copies them into shared memory with unit-stride loads, every lane
// then owns one batch element (A values, stride-A shared reads: conflict free for odd A, 2-way for A = 6), writes the
// target and the loss gradient back into the same shared slots and the warp stores them with unit-stride writes.
// Pass 2 reduces td_sq[L,B] per batch element (max / mean with the [b:-1:B] quirk) and sums the loss: lanes along b.
constexpr int TD1_WARPS = 8;
__global__ void __launch_bounds__(TD1_WARPS * 32) td_elem_kernel(TdPriorityParams p) {
  extern __shared__ float td_smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int L = p.L, B = p.B, A = p.A;
  const int i = blockIdx.y * TD1_WARPS + w;
  const int b0 = blockIdx.x * 32;
  if (i >= L) return;
  const int nb = min(32, B - b0), span = nb * A;
  float* sq_ = td_smem + (size_t)w * 2 * 32 * A;       // q, then (in place) dq
  float* sn_ = sq_ + 32 * A;                           // q_next, then (in place) target
  const size_t base = ((size_t)i * B + b0) * A;
  for (int k = lane; k < span; k += 32) {
    sq_[k] = __ldg(p.q + base + k);
    sn_[k] = __ldg(p.q_next + base + k);
  }
  __syncwarp();
  const float grad_scale = 2.0f / ((float)L * (float)B * (float)A);
  if (lane < nb) {
    const int b = b0 + lane;
    const float r = __ldg(p.rew + (size_t)(p.burn_in + i) * B + b);
    const float d = __ldg(p.term + (size_t)(p.burn_in + i + p.n_step - 1) * B + b);
    const float cont = p.gamma_n * (1.0f - d);
    float sq = 0.f;
    for (int a = 0; a < A; ++a) {
      const float y = value_rescale(r + cont * sn_[lane * A + a]);
      const float diff = sq_[lane * A + a] - y;
      sn_[lane * A + a] = y;
      sq_[lane * A + a] = grad_scale * diff;
      sq += diff * diff;
    }
    p.td_sq[(size_t)i * B + b] = sq / (float)A;
  }
  __syncwarp();
  if (p.target) for (int k = lane; k < span; k += 32) p.target[base + k] = sn_[k];
  if (p.dq) for (int k = lane; k < span; k += 32) p.dq[base + k] = sq_[k];
}

constexpr int TD2_WARPS = 8;
__global__ void __launch_bounds__(TD2_WARPS * 32) td_reduce_kernel(TdPriorityParams p) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  const int L = p.L, B = p.B;
  float run_max = -INFINITY, run_sum = 0.f, tot = 0.f;
  if (b < B) {
    for (int i = w; i < L; i += TD2_WARPS) {
      const float td = p.td_sq[(size_t)i * B + b];
      tot += td;
      // learner.py:137 `average_td_loss[b:-1:B]` drops flat index L*B-1, i.e. (i=L-1, b=B-1)
      if (!(i == L - 1 && b == B - 1)) { run_max = fmaxf(run_max, td); run_sum += td; }
    }
  }
  __shared__ float s_max[TD2_WARPS][32], s_sum[TD2_WARPS][32], s_tot[TD2_WARPS][32];
  s_max[w][lane] = run_max; s_sum[w][lane] = run_sum; s_tot[w][lane] = tot;
  __syncthreads();
  if (w == 0) {
    float mx = s_max[0][lane], sm = s_sum[0][lane], tt = s_tot[0][lane];
#pragma unroll
    for (int k = 1; k < TD2_WARPS; ++k) { mx = fmaxf(mx, s_max[k][lane]); sm += s_sum[k][lane]; tt += s_tot[k][lane]; }
    if (b < B && p.priority) {
      const int count = L - ((b == B - 1) ? 1 : 0);
      p.priority[b] = p.eta * mx + (1.0f - p.eta) * (sm / (float)count);  // utils.py:17-18
    }
    tt = warp_sum(tt);   // critic loss = mean over (i, b, a) of diff^2 = sum of td_sq / (L * B)
    if (lane == 0 && p.loss_sum) atomicAdd(p.loss_sum, tt / ((float)L * (float)B));
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, long long ld, int M, int N,
                                                     int rows_per_block, float* __restrict__ out, float* __restrict__ out2) {
  // block = 32 columns x 8 row lanes; grid.x over column chunks, grid.y over row ranges; atomics into out
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc = 0.f;
  if (col < N)
    for (int r = r0 + w; r < r1; r += 8) acc += __ldg(x + (size_t)r * ld + col);
  __shared__ float sm[8][32];
  sm[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][lane];
    atomicAdd(out + col, t);
    if (out2) atomicAdd(out2 + col, t);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                   float* __restrict__ m, float* __restrict__ v, long long n,
                                                   float grad_scale, float beta1, float beta2, float step_size,
                                                   float inv_bc2_sqrt, float eps) {
  // torch.optim.Adam single-tensor math (learner.py:50,52 defaults): lerp m, addcmul v, sqrt/bc2 + eps, addcdiv
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * grad_scale;
    const float mi = m[i] + (g - m[i]) * (1.0f - beta1);
    const float vi = v[i] * beta2 + (1.0f - beta2) * g * g;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_bc2_sqrt + eps;
    param[i] = param[i] - step_size * (mi / denom);
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ x, long long n, float value) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = value;
}

__global__ void __launch_bounds__(256) scaled_sum_kernel(const float* __restrict__ x, long long n, float scale,
                                                         float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += x[i];
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += sm[k];
    atomicAdd(out, t * scale);
  }
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

__global__ void mul_dtanh_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
                                 float* __restrict__ d_pre, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d_pre[i] = d_out[i] * (1.0f - out[i] * out[i]);
}

// ---- actor-side next rows: n-step reward pre-sum (actor.py:74-76) and the initial priorities of a
// finished episode (actor.py:78-107), batched over episodes (time-major [T,B], one episode per batch column, zero
// padded past its last row).
__global__ void __launch_bounds__(256) nstep_reward_kernel(const float* __restrict__ raw, const int* __restrict__ n_rows,
                                                           int T, int B, int n_step, float gamma, float* __restrict__ out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)T * B) return;
  const int i = (int)(idx / B), b = (int)(idx % B);
  float v = raw[idx];
  if (i < n_rows[b] - n_step) {   // rows of the real episode: discounted sum of the next n raw rewards
    double acc = 0.0, g = 1.0;
    for (int j = 0; j < n_step; ++j) { acc += (double)raw[(size_t)(i + j) * B + b] * g; g *= (double)gamma; }
    v = (float)acc;
  }
  out[idx] = v;
}

// priority k of episode b: 0.9 max + 0.1 mean over j = k+Bn+1 .. k+Bn+L of td_j^2,
// td_j = mean_A(q[j,b,:] - h(R[j,b] + gamma^n (1 - term[j+n-1,b]) q_next[j+n,b,:]))  -  the reference's deque of
// `learning` entries is one step ahead of the window the learner trains on (actor.py:102-107), reproduced here.
__global__ void __launch_bounds__(128) actor_priority_kernel(const float* __restrict__ q, const float* __restrict__ q_next,
                                                             const float* __restrict__ rew, const float* __restrict__ term,
                                                             const int* __restrict__ n_rows, int B, int A, int burn_in,
                                                             int learning, int n_step, float gamma_n, float eta,
                                                             int p_max, float* __restrict__ prio) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p_max) return;
  const int E = n_rows[b] - n_step;
  float out = 0.f;
  if (k < E - (burn_in + learning)) {
    float mx = -INFINITY, sum = 0.f;
    for (int j = k + burn_in + 1; j <= k + burn_in + learning; ++j) {
      const float r = rew[(size_t)j * B + b];
      const float cont = gamma_n * (1.0f - term[(size_t)(j + n_step - 1) * B + b]);
      const float* qj = q + ((size_t)j * B + b) * A;
      const float* qn = q_next + ((size_t)(j + n_step) * B + b) * A;
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc += qj[a] - value_rescale(r + cont * qn[a]);
      const float td = acc / (float)A;
      const float sq = td * td;
      mx = fmaxf(mx, sq);
      sum += sq;
    }
    out = eta * mx + (1.0f - eta) * (sum / (float)learning);
  }
  prio[(size_t)b * p_max + k] = out;
}

}  // namespace

int nstep_rewards(const float* raw, const int* n_rows, int T, int B, int n_step, float gamma, float* out, cudaStream_t stream) {
  R2D2_REQUIRE(raw && n_rows && out && raw != out && T > 0 && B > 0 && n_step > 0, "nstep_rewards args");
  nstep_reward_kernel<<<(unsigned)(((long long)T * B + 255) / 256), 256, 0, stream>>>(raw, n_rows, T, B, n_step, gamma, out);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int actor_priorities(const float* q, const float* q_next, const float* rew, const float* term, const int* n_rows, int B,
                     int A, int burn_in, int learning, int n_step, float gamma, float eta, int p_max, float* prio,
                     cudaStream_t stream) {
  R2D2_REQUIRE(q && q_next && rew && term && n_rows && prio && B > 0 && A > 0 && p_max > 0, "actor_priorities args");
  const float gamma_n = (float)std::pow((double)gamma, (double)n_step);
  actor_priority_kernel<<<dim3(ceil_div(p_max, 128), B), 128, 0, stream>>>(q, q_next, rew, term, n_rows, B, A, burn_in, learning,
                                                                           n_step, gamma_n, eta, p_max, prio);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t stream) {
  add_vec_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(a, b, out, n);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int mul_dtanh(const float* d_out, const float* out, float* d_pre, long long n, cudaStream_t stream) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  mul_dtanh_kernel<<<blocks, 256, 0, stream>>>(d_out, out, d_pre, n);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int td_priority(const TdPriorityParams& p, cudaStream_t stream) {
  R2D2_REQUIRE(p.q && p.q_next && p.rew && p.term, "null input");
  R2D2_REQUIRE(p.L > 0 && p.B > 0 && p.A > 0, "shape");
  if (p.loss_sum) R2D2_CUDA_TRY(cudaMemsetAsync(p.loss_sum, 0, sizeof(float), stream));
  const size_t smem = (size_t)TD1_WARPS * 2 * 32 * p.A * sizeof(float);
  if (p.td_sq && smem <= 48 * 1024) {   // the path's configuration: two line-coalesced passes over L x B x A and L x B
    td_elem_kernel<<<dim3(ceil_div(p.B, 32), ceil_div(p.L, TD1_WARPS)), TD1_WARPS * 32, smem, stream>>>(p);
    count_launch();
    if (p.priority || p.loss_sum) {
      td_reduce_kernel<<<ceil_div(p.B, 32), TD2_WARPS * 32, 0, stream>>>(p);
      count_launch();
    }
  } else {
    td_priority_column_kernel<<<ceil_div(p.B, 32), TD_WARPS * 32, 0, stream>>>(p);
    count_launch();
  }
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int colsum(const float* x, long long ld, int M, int N, float* out, float* out2, cudaStream_t stream) {
  R2D2_REQUIRE(x && out && M > 0 && N > 0, "colsum args");
  const int col_blocks = ceil_div(N, 32);
  int row_blocks = ceil_div(4 * 148, col_blocks);
  if (row_blocks > ceil_div(M, 64)) row_blocks = ceil_div(M, 64);
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = ceil_div(M, row_blocks);
  colsum_kernel<<<dim3(col_blocks, ceil_div(M, rows_per_block)), 256, 0, stream>>>(x, ld, M, N, rows_per_block, out, out2);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int adam_step(float* param, const float* grad, float* m, float* v, long long n, int step, float lr, float beta1,
              float beta2, float eps, float grad_scale, cudaStream_t stream) {
  R2D2_REQUIRE(param && grad && m && v && n > 0 && step >= 1, "adam args");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<blocks, 256, 0, stream>>>(param, grad, m, v, n, grad_scale, beta1, beta2, step_size, inv_bc2_sqrt, eps);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int fill_f32(float* x, long long n, float value, cudaStream_t stream) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  fill_kernel<<<blocks, 256, 0, stream>>>(x, n, value);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int scaled_sum(const float* x, long long n, float scale, float* out, cudaStream_t stream) {
  R2D2_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float), stream));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  scaled_sum_kernel<<<blocks, 256, 0, stream>>>(x, n, scale, out);
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

}  // namespace r2d2
