// Persistent LSTM scan kernels (forward + BPTT) for sm_100a.
//
// Decomposition (H in {32,64,128,256}):  a thread-block cluster of C = H/32 CTAs owns NB batch
// rows for the whole chain.  CTA `rank` owns hidden units [32*rank, 32*rank+32): its 128 gate rows of
// W_hh stay resident in REGISTERS as bf16 hi/lo MMA fragments for the whole launch (loaded once);
// per step the only traffic is the h_t all-gather inside the cluster through distributed shared
// memory plus one hardware cluster barrier.  Batch is split across clusters (grid = C * ceil(B/NB)).
//   forward : D[gate rows(128) x NB] = W_slice[128 x H] * h_{s-1}^T           (swap-AB: M = gate rows)
//   backward: P[H x NB] = W_slice^T[H x 128] * dG_s^T  -> reduce-scatter of fp32 partials via DSMEM
// Tensor cores: mma.sync bf16, 3 passes (hi*hi, lo*hi, hi*lo), fp32 accumulate.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "gemm.cuh"
#include "lstm_scan.cuh"
#include "elementwise.cuh"

namespace cg = cooperative_groups;

namespace r2d2 {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int UNITS_PER_CTA = 32;
constexpr int ROWS_PER_CTA = 128;         // 4 gates x 32 units
constexpr int GT_LD = ROWS_PER_CTA + 4;   // fp32 gate tile [NB][132]: conflict-free fragment writes / unit reads
constexpr int DG_LD = ROWS_PER_CTA + 8;   // bf16 dG tile [NB][136]

__device__ __forceinline__ float accurate_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int H, int NB>
struct FwdSmem {
  static constexpr int HLD = H + 8;                        // bf16 row stride of the h operand tile
  static constexpr int HB_ELEMS = 2 * 2 * NB * HLD;        // [buf][plane][n][k]
  static constexpr int GT_ELEMS = NB * GT_LD;              // fp32
  static constexpr int HS_ELEMS = 2 * NB * UNITS_PER_CTA;  // bf16 staging [plane][n][unit]
  static constexpr int BYTES = HB_ELEMS * 2 + GT_ELEMS * 4 + HS_ELEMS * 2;
};

template <int H, int NB>
__global__ void __launch_bounds__(SCAN_THREADS, 1) lstm_scan_fwd_kernel(ScanFwdParams p) {
  constexpr int C = H / 32, KS = H / 16, NT = NB / 8;
  using SM = FwdSmem<H, NB>;
  constexpr int HLD = SM::HLD;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / C) * NB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int B = p.B, S = p.T * p.repeat;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  float* gt = reinterpret_cast<float*>(smem_raw + SM::HB_ELEMS * 2);
  __nv_bfloat16* hstage = reinterpret_cast<__nv_bfloat16*>(smem_raw + SM::HB_ELEMS * 2 + SM::GT_ELEMS * 4);

  // ---- W_hh slice -> register-resident A fragments (rows: local r = gate*32 + unit; warp w owns r in [16w,16w+16))
  uint32_t a_hi[KS][4], a_lo[KS][4];
  {
    const int gate = w >> 1;
    const int u_lo = (w & 1) * 16 + g;  // local unit of fragment row g; row g+8 -> unit u_lo+8
    const float* w_r0 = p.whh + (size_t)(gate * H + rank * 32 + u_lo) * H;
    const float* w_r1 = w_r0 + (size_t)8 * H;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = ks * 16 + 2 * c;
      float2 v0 = __ldg(reinterpret_cast<const float2*>(w_r0 + k));
      float2 v1 = __ldg(reinterpret_cast<const float2*>(w_r1 + k));
      float2 v2 = __ldg(reinterpret_cast<const float2*>(w_r0 + k + 8));
      float2 v3 = __ldg(reinterpret_cast<const float2*>(w_r1 + k + 8));
      split_pack2(v0.x, v0.y, a_hi[ks][0], a_lo[ks][0]);
      split_pack2(v1.x, v1.y, a_hi[ks][1], a_lo[ks][1]);
      split_pack2(v2.x, v2.y, a_hi[ks][2], a_lo[ks][2]);
      split_pack2(v3.x, v3.y, a_hi[ks][3], a_lo[ks][3]);
    }
  }

  // ---- initial state: full h0 tile -> hb[0]; own units -> hs[0], cs[0], c registers
  for (int idx = tid; idx < NB * H; idx += SCAN_THREADS) {
    const int n = idx / H, k = idx % H, b = b0 + n;
    float v = (b < B && p.h0) ? __ldg(p.h0 + (size_t)b * H + k) : 0.f;
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    hb[(0 * 2 + 0) * NB * HLD + n * HLD + k] = hi;
    hb[(0 * 2 + 1) * NB * HLD + n * HLD + k] = lo;
  }
  const int ug = rank * 32 + lane;  // global hidden unit owned by this thread in the pointwise phase
  float cst[NT];
#pragma unroll
  for (int e = 0; e < NT; ++e) {
    const int b = b0 + w + 8 * e;
    cst[e] = 0.f;
    if (b < B) {
      float hv = p.h0 ? __ldg(p.h0 + (size_t)b * H + ug) : 0.f;
      float cv = p.c0 ? __ldg(p.c0 + (size_t)b * H + ug) : 0.f;
      p.hs[(size_t)b * H + ug] = hv;
      p.cs[(size_t)b * H + ug] = cv;
      cst[e] = cv;
    }
  }
  __syncthreads();
  cluster.sync();  // every CTA of the cluster has started: remote shared memory may be written from here on

  const size_t gstride = (size_t)4 * H;
  for (int s = 0; s < S; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    const int t = s / p.repeat;

    // prefetch this step's input projection for the elements this thread finishes (hides HBM/L2 latency behind the MMAs)
    float gpre[NT][4];
#pragma unroll
    for (int e = 0; e < NT; ++e) {
      const int b = b0 + w + 8 * e;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        gpre[e][q] = (b < B) ? p.gin[((size_t)t * B + b) * gstride + q * H + ug] : 0.f;  // plain load: gates may alias gin
    }

    // ---- tensor-core part: acc[128 x NB] = W_slice * h_{s-1}^T
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    const __nv_bfloat16* hb_hi = hb + (cur * 2 + 0) * NB * HLD;
    const __nv_bfloat16* hb_lo = hb + (cur * 2 + 1) * NB * HLD;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int off = (nt * 8 + g) * HLD + ks * 16 + 2 * c;
        uint32_t bh[2], bl[2];
        bh[0] = *reinterpret_cast<const uint32_t*>(hb_hi + off);
        bh[1] = *reinterpret_cast<const uint32_t*>(hb_hi + off + 8);
        bl[0] = *reinterpret_cast<const uint32_t*>(hb_lo + off);
        bl[1] = *reinterpret_cast<const uint32_t*>(hb_lo + off + 8);
        mma_bf16_16816(acc[nt], a_lo[ks], bh);
        mma_bf16_16816(acc[nt], a_hi[ks], bl);
        mma_bf16_16816(acc[nt], a_hi[ks], bh);
      }
    }
    // fragment -> gate tile gt[n][local row]
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = nt * 8 + 2 * c, r = 16 * w + g;
      gt[n * GT_LD + r] = acc[nt][0];
      gt[(n + 1) * GT_LD + r] = acc[nt][1];
      gt[n * GT_LD + r + 8] = acc[nt][2];
      gt[(n + 1) * GT_LD + r + 8] = acc[nt][3];
    }
    __syncthreads();

    // ---- pointwise LSTM cell: thread = (unit = lane, batch column n = w + 8e); coalesced along units
#pragma unroll
    for (int e = 0; e < NT; ++e) {
      const int n = w + 8 * e, b = b0 + n;
      __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
      if (b < B) {
        const float* gr = gt + n * GT_LD + lane;
        const float ig = accurate_sigmoid(gr[0] + gpre[e][0]);
        const float fg = accurate_sigmoid(gr[32] + gpre[e][1]);
        const float gg = tanhf(gr[64] + gpre[e][2]);
        const float og = accurate_sigmoid(gr[96] + gpre[e][3]);
        const float cn = fg * cst[e] + ig * gg;
        const float hn = og * tanhf(cn);
        cst[e] = cn;
        float* go = p.gates + ((size_t)s * B + b) * gstride + ug;
        go[0] = ig; go[H] = fg; go[2 * H] = gg; go[3 * H] = og;
        p.hs[((size_t)(s + 1) * B + b) * H + ug] = hn;
        p.cs[((size_t)(s + 1) * B + b) * H + ug] = cn;
        if (p.head_in && (s % p.repeat) == p.repeat - 1)
          p.head_in[((size_t)t * B + b) * H + ug] = tanhf(hn);
        split_bf16(hn, hi, lo);
      }
      hstage[(0 * NB + n) * UNITS_PER_CTA + lane] = hi;
      hstage[(1 * NB + n) * UNITS_PER_CTA + lane] = lo;
    }
    __syncthreads();

    // ---- all-gather of h_s inside the cluster: 16-byte DSMEM stores into every CTA's next operand buffer
    if (s + 1 < S) {
      constexpr int VEC_PER_CTA = 2 * NB * 4;  // [plane][n][4 x 16 B]
      for (int vv = tid; vv < C * VEC_PER_CTA; vv += SCAN_THREADS) {
        const int d = vv / VEC_PER_CTA, v = vv % VEC_PER_CTA;
        const int plane = v / (NB * 4), n = (v / 4) % NB, q4 = v % 4;
        const uint4 val = *reinterpret_cast<const uint4*>(hstage + (plane * NB + n) * UNITS_PER_CTA + q4 * 8);
        __nv_bfloat16* dst_local = hb + ((nxt * 2 + plane) * NB + n) * HLD + rank * 32 + q4 * 8;
        __nv_bfloat16* dst = cluster.map_shared_rank(dst_local, d);
        *reinterpret_cast<uint4*>(dst) = val;
      }
    }
    cluster.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// backward (BPTT)
// ------------------------------------------------------------------------------------------------
template <int H, int NB>
struct BwdSmem {
  static constexpr int C = H / 32;
  static constexpr int DG_ELEMS = 2 * NB * DG_LD;               // bf16 [plane][n][local gate row]
  static constexpr int PS_LD = NB + 2;                          // 2-way instead of 16-way bank conflicts on the unit-strided reads
  static constexpr int PS_ELEMS = 2 * C * UNITS_PER_CTA * PS_LD; // fp32 [buf][src rank][unit][n]
  static constexpr int BYTES = DG_ELEMS * 2 + PS_ELEMS * 4;
};

template <int H, int NB>
__global__ void __launch_bounds__(SCAN_THREADS, 1) lstm_scan_bwd_kernel(ScanBwdParams p) {
  constexpr int C = H / 32, NT = NB / 8;
  constexpr int M_TILES = H / 16;
  constexpr int MT = (M_TILES + 7) / 8;  // m-tiles per warp
  constexpr int KS = ROWS_PER_CTA / 16;  // 8
  using SM = BwdSmem<H, NB>;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b0 = (blockIdx.x / C) * NB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int B = p.B, S = p.T * p.repeat;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* dgs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  float* ps = reinterpret_cast<float*>(smem_raw + SM::DG_ELEMS * 2);

  // ---- W_hh slice (transposed use): A(m = j output unit, k = local gate row r) = W_hh[grow(r)][j]
  uint32_t a_hi[MT][KS][4], a_lo[MT][KS][4];
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int mi = w * MT + i;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        float v0 = 0.f, v1 = 0.f;
        if (mi < M_TILES) {
          const int j = mi * 16 + g + (f & 1) * 8;
          const int r = ks * 16 + 2 * c + (f >> 1) * 8;  // r, r+1: local gate rows (gate = r/32, unit = r%32)
          const size_t row0 = (size_t)((r >> 5) * H + rank * 32 + (r & 31));
          v0 = __ldg(p.whh + row0 * H + j);
          v1 = __ldg(p.whh + (row0 + 1) * H + j);  // r even -> r+1 stays in the same gate block
        }
        split_pack2(v0, v1, a_hi[i][ks][f], a_lo[i][ks][f]);
      }
    }
  }

  const int ug = rank * 32 + lane;
  const size_t gstride = (size_t)4 * H;
  float dcn[NT], keep[NT][4];
#pragma unroll
  for (int e = 0; e < NT; ++e) {
    dcn[e] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) keep[e][q] = 0.f;
  }
  cluster.sync();

  for (int it = 0; it < S; ++it) {
    const int s = S - 1 - it;
    const int buf = it & 1;
    const int t = s / p.repeat;
    const int rel = s - p.head_first_step;
    const bool has_head = p.dh_head && rel >= 0 && (rel % p.repeat) == p.repeat - 1;

    // ---- pointwise backward of the cell (thread = (unit = lane, n = w + 8e))
#pragma unroll
    for (int e = 0; e < NT; ++e) {
      const int n = w + 8 * e, b = b0 + n;
      float dg[4] = {0.f, 0.f, 0.f, 0.f};
      if (b < B) {
        float dh = 0.f;
        if (it > 0) {
#pragma unroll
          for (int src = 0; src < C; ++src) dh += ps[((buf * C + src) * UNITS_PER_CTA + lane) * SM::PS_LD + n];
        }
        if (has_head) dh += __ldg(p.dh_head + ((size_t)(rel / p.repeat) * B + b) * H + ug);
        const float* gs = p.gates + ((size_t)s * B + b) * gstride + ug;
        const float ig = gs[0], fg = gs[H], gg = gs[2 * H], og = gs[3 * H];
        const float c_prev = __ldg(p.cs + ((size_t)s * B + b) * H + ug);
        const float tc = tanhf(__ldg(p.cs + ((size_t)(s + 1) * B + b) * H + ug));
        const float dc = dcn[e] + dh * og * (1.f - tc * tc);
        dg[3] = dh * tc * og * (1.f - og);
        dg[0] = dc * gg * ig * (1.f - ig);
        dg[1] = dc * c_prev * fg * (1.f - fg);
        dg[2] = dc * ig * (1.f - gg * gg);
        dcn[e] = dc * fg;
        float* go = p.dgates + ((size_t)s * B + b) * gstride + ug;
        go[0] = dg[0]; go[H] = dg[1]; go[2 * H] = dg[2]; go[3 * H] = dg[3];
        if (p.repeat > 1) {
#pragma unroll
          for (int q = 0; q < 4; ++q) keep[e][q] += dg[q];
          if (s % p.repeat == 0) {
            float* gi = p.dgin + ((size_t)t * B + b) * gstride + ug;
#pragma unroll
            for (int q = 0; q < 4; ++q) { gi[q * H] = keep[e][q]; keep[e][q] = 0.f; }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        __nv_bfloat16 hi, lo;
        split_bf16(dg[q], hi, lo);
        dgs[(0 * NB + n) * DG_LD + q * 32 + lane] = hi;
        dgs[(1 * NB + n) * DG_LD + q * 32 + lane] = lo;
      }
    }
    __syncthreads();

    if (s > 0) {
      // ---- partial dh_{s-1}[j, n] over this CTA's 128 gate rows
      float acc[MT][NT][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][nt][e] = 0.f;
      const __nv_bfloat16* d_hi = dgs;
      const __nv_bfloat16* d_lo = dgs + NB * DG_LD;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int off = (nt * 8 + g) * DG_LD + ks * 16 + 2 * c;
          uint32_t bh[2], bl[2];
          bh[0] = *reinterpret_cast<const uint32_t*>(d_hi + off);
          bh[1] = *reinterpret_cast<const uint32_t*>(d_hi + off + 8);
          bl[0] = *reinterpret_cast<const uint32_t*>(d_lo + off);
          bl[1] = *reinterpret_cast<const uint32_t*>(d_lo + off + 8);
#pragma unroll
          for (int i = 0; i < MT; ++i) {
            mma_bf16_16816(acc[i][nt], a_lo[i][ks], bh);
            mma_bf16_16816(acc[i][nt], a_hi[i][ks], bl);
            mma_bf16_16816(acc[i][nt], a_hi[i][ks], bh);
          }
        }
      }
      // ---- reduce-scatter: partial rows j go to the CTA that owns unit j (slot = my rank), fp32 via DSMEM
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int mi = w * MT + i;
        if (mi < M_TILES) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int j = mi * 16 + g + h2 * 8;
            const int owner = j >> 5, jl = j & 31;
            float* slot_local = ps + (((buf ^ 1) * C + rank) * UNITS_PER_CTA + jl) * SM::PS_LD;
            float* slot = cluster.map_shared_rank(slot_local, owner);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
              *reinterpret_cast<float2*>(slot + nt * 8 + 2 * c) = make_float2(acc[i][nt][2 * h2], acc[i][nt][2 * h2 + 1]);
          }
        }
      }
    }
    cluster.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// generic path (any H): one GEMM + one pointwise kernel per step.  Correct for every size; used
// only when the cluster kernels do not cover H (e.g. H = 512 until its kernel lands).
// ------------------------------------------------------------------------------------------------
__global__ void lstm_cell_fwd_pointwise(const float* __restrict__ gpre, const float* __restrict__ c_prev,
                                        float* __restrict__ gates, float* __restrict__ h_out,
                                        float* __restrict__ c_out, float* __restrict__ head_in, int B, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  const float* gr = gpre + (size_t)b * 4 * H + u;
  const float ig = accurate_sigmoid(gr[0]), fg = accurate_sigmoid(gr[H]);
  const float gg = tanhf(gr[2 * H]), og = accurate_sigmoid(gr[3 * H]);
  const float cn = fg * c_prev[idx] + ig * gg;
  const float hn = og * tanhf(cn);
  float* go = gates + (size_t)b * 4 * H + u;
  go[0] = ig; go[H] = fg; go[2 * H] = gg; go[3 * H] = og;
  h_out[idx] = hn;
  c_out[idx] = cn;
  if (head_in) head_in[idx] = tanhf(hn);
}

__global__ void lstm_cell_bwd_pointwise(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                        const float* __restrict__ c_new, const float* __restrict__ dh_head,
                                        const float* __restrict__ dh_rec, float* __restrict__ dc_state,
                                        float* __restrict__ dgates, float* __restrict__ dgin, int dgin_mode,
                                        int B, int H) {
  // dgin_mode: 0 = none, 1 = overwrite (first visit of this input row), 2 = accumulate
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, u = idx % H;
  const float* gs = gates + (size_t)b * 4 * H + u;
  const float ig = gs[0], fg = gs[H], gg = gs[2 * H], og = gs[3 * H];
  float dh = dh_rec ? dh_rec[idx] : 0.f;
  if (dh_head) dh += dh_head[idx];
  const float tc = tanhf(c_new[idx]);
  const float dc = dc_state[idx] + dh * og * (1.f - tc * tc);
  float dg[4];
  dg[3] = dh * tc * og * (1.f - og);
  dg[0] = dc * gg * ig * (1.f - ig);
  dg[1] = dc * c_prev[idx] * fg * (1.f - fg);
  dg[2] = dc * ig * (1.f - gg * gg);
  dc_state[idx] = dc * fg;
  float* go = dgates + (size_t)b * 4 * H + u;
  float* gi = dgin ? dgin + (size_t)b * 4 * H + u : nullptr;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    go[q * H] = dg[q];
    if (dgin_mode == 1) gi[q * H] = dg[q];
    else if (dgin_mode == 2) gi[q * H] += dg[q];
  }
}

int scan_forward_generic(const ScanFwdParams& p, float* scratch, cudaStream_t stream) {
  const int B = p.B, H = p.H, S = p.T * p.repeat;
  const size_t bh = (size_t)B * H;
  R2D2_REQUIRE(scratch != nullptr, "generic scan needs scratch");
  if (p.h0) R2D2_CUDA_TRY(cudaMemcpyAsync(p.hs, p.h0, bh * 4, cudaMemcpyDeviceToDevice, stream));
  else R2D2_CUDA_TRY(cudaMemsetAsync(p.hs, 0, bh * 4, stream));
  if (p.c0) R2D2_CUDA_TRY(cudaMemcpyAsync(p.cs, p.c0, bh * 4, cudaMemcpyDeviceToDevice, stream));
  else R2D2_CUDA_TRY(cudaMemsetAsync(p.cs, 0, bh * 4, stream));
  const int blocks = ceil_div((int)bh, 256);
  for (int s = 0; s < S; ++s) {
    const int t = s / p.repeat;
    GemmParams g;
    g.A = p.hs + (size_t)s * bh; g.lda = H;
    g.B = p.whh; g.ldb = H;
    g.C = scratch; g.ldc = 4 * H;
    g.M = B; g.N = 4 * H; g.K = H;
    g.Z = p.gin + (size_t)t * B * 4 * H; g.ldz = 4 * H;
    g.epilogue = EPI_ADD_Z;
    R2D2_TRY(gemm_f32(g, GEMM_NT, stream));
    float* head = (p.head_in && (s % p.repeat) == p.repeat - 1) ? p.head_in + (size_t)t * bh : nullptr;
    lstm_cell_fwd_pointwise<<<blocks, 256, 0, stream>>>(scratch, p.cs + (size_t)s * bh,
                                                         p.gates + (size_t)s * B * 4 * H, p.hs + (size_t)(s + 1) * bh,
                                                         p.cs + (size_t)(s + 1) * bh, head, B, H);
    count_launch();
  }
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

int scan_backward_generic(const ScanBwdParams& p, cudaStream_t stream) {
  const int B = p.B, H = p.H, S = p.T * p.repeat;
  const size_t bh = (size_t)B * H;
  R2D2_REQUIRE(p.scratch != nullptr, "generic scan needs scratch");
  float* dh_rec = p.scratch;
  float* dc_state = p.scratch + bh;
  R2D2_CUDA_TRY(cudaMemsetAsync(p.scratch, 0, 2 * bh * 4, stream));
  const int blocks = ceil_div((int)bh, 256);
  for (int s = S - 1; s >= 0; --s) {
    const int t = s / p.repeat;
    const int rel = s - p.head_first_step;
    const bool has_head = p.dh_head && rel >= 0 && (rel % p.repeat) == p.repeat - 1;
    const float* head = has_head ? p.dh_head + (size_t)(rel / p.repeat) * bh : nullptr;
    int mode = 0;
    if (p.repeat > 1) mode = ((s % p.repeat) == p.repeat - 1) ? 1 : 2;
    lstm_cell_bwd_pointwise<<<blocks, 256, 0, stream>>>(
        p.gates + (size_t)s * B * 4 * H, p.cs + (size_t)s * bh, p.cs + (size_t)(s + 1) * bh, head,
        (s == S - 1) ? nullptr : dh_rec, dc_state, p.dgates + (size_t)s * B * 4 * H,
        mode ? p.dgin + (size_t)t * B * 4 * H : nullptr, mode, B, H);
    count_launch();
    if (s > 0) {
      GemmParams g;
      g.A = p.dgates + (size_t)s * B * 4 * H; g.lda = 4 * H;
      g.B = p.whh; g.ldb = H;
      g.C = dh_rec; g.ldc = H;
      g.M = B; g.N = H; g.K = 4 * H;
      R2D2_TRY(gemm_f32(g, GEMM_NN, stream));
    }
  }
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

template <typename Kern, typename Params>
int launch_cluster(Kern kern, const Params& p, int cluster_size, int n_clusters, int smem_bytes, cudaStream_t stream) {
  R2D2_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cluster_size * n_clusters);
  cfg.blockDim = dim3(SCAN_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  R2D2_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  count_launch();
  return R2D2_OK;
}

// batch columns per cluster: the smallest tile that still fits one wave of clusters (per-step cost is
// dominated by fixed latencies, so more, narrower clusters win as long as they are co-resident)
int pick_nb(int B, int H) {
  const int avail = 148 / (H / 32);
  const int opts[3] = {8, 16, 32};
  for (int i = 0; i < 3; ++i)
    if (ceil_div(B, opts[i]) <= avail) return opts[i];
  return 32;
}

template <int H>
int fwd_dispatch(const ScanFwdParams& p, cudaStream_t stream) {
  const int nb = pick_nb(p.B, H);
  if (nb == 8)
    return launch_cluster(lstm_scan_fwd_kernel<H, 8>, p, H / 32, ceil_div(p.B, 8), FwdSmem<H, 8>::BYTES, stream);
  if (nb == 16)
    return launch_cluster(lstm_scan_fwd_kernel<H, 16>, p, H / 32, ceil_div(p.B, 16), FwdSmem<H, 16>::BYTES, stream);
  return launch_cluster(lstm_scan_fwd_kernel<H, 32>, p, H / 32, ceil_div(p.B, 32), FwdSmem<H, 32>::BYTES, stream);
}
template <int H>
int bwd_dispatch(const ScanBwdParams& p, cudaStream_t stream) {
  const int nb = pick_nb(p.B, H);
  if (nb == 8)
    return launch_cluster(lstm_scan_bwd_kernel<H, 8>, p, H / 32, ceil_div(p.B, 8), BwdSmem<H, 8>::BYTES, stream);
  if (nb == 16)
    return launch_cluster(lstm_scan_bwd_kernel<H, 16>, p, H / 32, ceil_div(p.B, 16), BwdSmem<H, 16>::BYTES, stream);
  return launch_cluster(lstm_scan_bwd_kernel<H, 32>, p, H / 32, ceil_div(p.B, 32), BwdSmem<H, 32>::BYTES, stream);
}

}  // namespace

// v1 (mma.sync) cluster kernels: weights as register fragments, up to 256 hidden units
static bool v1_cluster_supported(int H) { return H == 32 || H == 64 || H == 128 || H == 256; }
// tcgen05 kernels: up to 512 hidden units (cluster of 16 CTAs, W_hh lo plane partly in shared memory)
bool lstm_scan_cluster_supported(int H) { return v1_cluster_supported(H) || H == 512; }
bool lstm_scan_backward_emits_images(int H) { return lstm_scan_cluster_supported(H) && lstm_scan_get_impl() == 1; }

static int g_scan_impl = -1;
void lstm_scan_set_impl(int impl) { g_scan_impl = impl ? 1 : 0; }
int lstm_scan_get_impl() {
  if (g_scan_impl < 0) {
    const char* e = getenv("R2D2_SCAN_IMPL");
    g_scan_impl = (e && (e[0] == 'm' || e[0] == '0')) ? 0 : 1;
  }
  return g_scan_impl;
}

size_t lstm_scan_fwd_scratch_floats(int B, int H) {   // the per-step path is the A/B fallback wherever v1 has no kernel
  return v1_cluster_supported(H) ? 0 : (size_t)B * 4 * H;
}
size_t lstm_scan_bwd_scratch_floats(int B, int H) {
  return v1_cluster_supported(H) ? 0 : (size_t)2 * B * H;
}

int lstm_scan_forward(const ScanFwdParams& p, cudaStream_t stream) {
  R2D2_REQUIRE(p.gin && p.whh && p.gates && p.hs && p.cs, "null pointer");
  R2D2_REQUIRE(p.T > 0 && p.B > 0 && p.H > 0 && p.repeat >= 1, "shape");
  R2D2_REQUIRE(p.repeat == 1 || p.gates != p.gin, "gates must not alias gin when repeat > 1");
  if (lstm_scan_cluster_supported(p.H) && lstm_scan_get_impl() == 1) return lstm_scan_forward_tc(p, stream);
  switch (p.H) {
    case 32: return fwd_dispatch<32>(p, stream);
    case 64: return fwd_dispatch<64>(p, stream);
    case 128: return fwd_dispatch<128>(p, stream);
    case 256: return fwd_dispatch<256>(p, stream);
    default: break;
  }
  return scan_forward_generic(p, p.scratch, stream);
}

int lstm_scan_backward(const ScanBwdParams& p, cudaStream_t stream) {
  R2D2_REQUIRE(p.gates && p.hs && p.cs && p.whh && p.dgates, "null pointer");
  R2D2_REQUIRE(p.T > 0 && p.B > 0 && p.H > 0 && p.repeat >= 1, "shape");
  R2D2_REQUIRE(p.repeat == 1 || (p.dgin && p.dgin != p.dgates), "dgin buffer required when repeat > 1");
  if (lstm_scan_cluster_supported(p.H) && lstm_scan_get_impl() == 1) return lstm_scan_backward_tc(p, stream);  // bias sums fused
  int rc;
  switch (p.H) {
    case 32: rc = bwd_dispatch<32>(p, stream); break;
    case 64: rc = bwd_dispatch<64>(p, stream); break;
    case 128: rc = bwd_dispatch<128>(p, stream); break;
    case 256: rc = bwd_dispatch<256>(p, stream); break;
    default: rc = scan_backward_generic(p, stream); break;
  }
  R2D2_REQUIRE(!p.skip_fp32, "skip_fp32 needs the tcgen05 scan (lstm_scan_backward_emits_images)");
  R2D2_TRY(rc);
  if (p.dbias) {  // sum_t dgin_t == sum_s dgates_s
    const float* src = p.repeat > 1 ? p.dgin : p.dgates;
    R2D2_TRY(colsum(src, 4 * p.H, p.T * p.B, 4 * p.H, p.dbias, p.dbias2, stream));
  }
  return R2D2_OK;
}

}  // namespace r2d2
