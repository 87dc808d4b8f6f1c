// gemm_f32 on the 5th-gen tensor cores: fp32 operands in HBM are split into bf16 hi/lo planes while they are
// staged into shared memory (UMMA core-matrix layouts, K-major or MN-major so no transposes are needed for
// the dgrad / wgrad layouts); one elected thread issues tcgen05.mma (M=128, N<=128, K=16, three passes
// lo*hi + hi*lo + hi*hi) into a TMEM accumulator; warps 0-3 read it back with tcgen05.ld for the fused
// epilogue (bias / tanh / dtanh / add, or split-K reduction with vector red.global.add).
//
//   warps 0..7  producers: cp.async global fp32 -> raw smem ring (2 x 32 KB in flight per CTA: global/L2 latency is
//               hidden here; a register-staged prefetch exposed one full load latency per k-tile, 57% of samples)
//               -> ld.shared own slots -> split -> 16-byte st.shared into the UMMA operand stage
//   warp  8     MMA issuer: waits `full`, 6 UTCHMMA per 32-deep stage, tcgen05.commit -> `empty`
//   96 KB of shared memory and 128 TMEM columns per CTA: two CTAs per SM overlap each other's phases.
#include "gemm.cuh"
#include "tc05.cuh"

namespace r2d2 {
namespace {

constexpr int TBM = 128, TBN = 128, TBK = 32;
constexpr int TSTAGES = 1;                          // UMMA operand stages (bf16 hi/lo core matrices)
constexpr int RSTAGES = 2;                          // raw fp32 stages filled by cp.async (global latency hidden here)
constexpr int PRODUCER_WARPS = 8, TC_GEMM_THREADS = (PRODUCER_WARPS + 1) * 32;
constexpr int PLANE_BYTES = 128 * TBK * 2;          // one bf16 plane of a 128 x 32 operand tile = 8 KB
constexpr int STAGE_BYTES = 4 * PLANE_BYTES;        // A hi, A lo, B hi, B lo
constexpr int RAW_STAGE_BYTES = 2 * 128 * TBK * 4;  // A + B tiles in fp32 = 32 KB
constexpr int OFF_RAW = TSTAGES * STAGE_BYTES;
constexpr int OFF_BARS = OFF_RAW + RSTAGES * RAW_STAGE_BYTES;
constexpr int TC_GEMM_SMEM = OFF_BARS + 128;        // 96 KB + barriers: two CTAs per SM
constexpr int GROUPS = 128 * TBK / 8;               // 512 groups of 8 elements per operand tile

// one group = 8 consecutive elements along the operand's contiguous global dimension -> one 16-byte smem row
// of a core matrix.  The smem byte offset of group `id` is id*16 for both majors, and a warp's 32 groups are
// 8 global rows x 128 contiguous bytes (4 lanes share a cache line), written as 4 conflict-free 128-byte core matrices:
//   K-major  (global [row][k]):  row = 8*(id/32) + id%8, k = 8*((id/8)%4) .. +8      smem [row/8][k/8][row%8][16 B]
//                                 -> descriptor LBO (k-group stride) = 128, SBO (8-row-group stride) = 512
//   MN-major (global [k][mn]) :  k = 8*(id/128) + id%8, mn = 8*((id/8)%16) .. +8     smem [k/8][mn/8][k%8][16 B]
//                                 -> descriptor LBO (k-group stride) = 2048, SBO (8-mn-group stride) = 128
struct OperandSrc {
  const float* ptr;
  long long ld;
  int mn0;       // first row (M or N index) of this CTA's tile
  int mn_lim;    // M or N
  int k0;        // first k of the tile
  int k_lim;     // K of the segment
  int vec;       // 16-byte vector loads allowed
};

template <bool MN_MAJOR>
__device__ __forceinline__ void load_group(const OperandSrc& o, int id, float4& v0, float4& v1) {
  v0 = make_float4(0.f, 0.f, 0.f, 0.f);
  v1 = v0;
  int row, col, row_lim, col_lim;  // global [row][col], 8 consecutive cols
  if (MN_MAJOR) { row = o.k0 + 8 * (id >> 7) + (id & 7); col = o.mn0 + 8 * ((id >> 3) & 15); row_lim = o.k_lim; col_lim = o.mn_lim; }
  else          { row = o.mn0 + 8 * (id >> 5) + (id & 7); col = o.k0 + 8 * ((id >> 3) & 3);  row_lim = o.mn_lim; col_lim = o.k_lim; }
  if (row >= row_lim || col >= col_lim) return;
  const float* p = o.ptr + (long long)row * o.ld + col;
  if (o.vec && col + 7 < col_lim) {
    v0 = __ldg(reinterpret_cast<const float4*>(p));
    v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
  } else {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = (col + i < col_lim) ? __ldg(p + i) : 0.f;
    v0 = make_float4(t[0], t[1], t[2], t[3]);
    v1 = make_float4(t[4], t[5], t[6], t[7]);
  }
}

// global coordinates of group `id` of a tile: pointer to its first element (or nullptr when the whole group is
// out of range) and the number of valid elements (0..8)
template <bool MN_MAJOR>
__device__ __forceinline__ const float* group_src(const OperandSrc& o, int id, int& n_valid) {
  int row, col, row_lim, col_lim;
  if (MN_MAJOR) { row = o.k0 + 8 * (id >> 7) + (id & 7); col = o.mn0 + 8 * ((id >> 3) & 15); row_lim = o.k_lim; col_lim = o.mn_lim; }
  else          { row = o.mn0 + 8 * (id >> 5) + (id & 7); col = o.k0 + 8 * ((id >> 3) & 3);  row_lim = o.mn_lim; col_lim = o.k_lim; }
  if (row >= row_lim || col >= col_lim) { n_valid = 0; return nullptr; }
  n_valid = min(8, col_lim - col);
  return o.ptr + (long long)row * o.ld + col;
}

// 32 bytes of fp32 -> this thread's raw slot: two 16-byte cp.async with zero fill past `n_valid` elements; operands
// whose rows are not 16-byte aligned (ld % 4 != 0: obs = 17 columns ...) go through registers instead
template <bool MN_MAJOR>
__device__ __forceinline__ void fetch_group(const OperandSrc& o, int id, unsigned char* raw_slot) {
  int nv;
  const float* src = group_src<MN_MAJOR>(o, id, nv);
  if (o.vec) {
    const uint32_t dst = tc::smem_u32(raw_slot);
    const float* s0 = src ? src : o.ptr;
    const int b0 = min(nv, 4) * 4, b1 = max(nv - 4, 0) * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(s0), "r"(b0) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16), "l"(b1 ? s0 + 4 : o.ptr), "r"(b1) : "memory");
  } else {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = (i < nv) ? __ldg(src + i) : 0.f;
    *reinterpret_cast<float4*>(raw_slot) = make_float4(t[0], t[1], t[2], t[3]);
    *reinterpret_cast<float4*>(raw_slot + 16) = make_float4(t[4], t[5], t[6], t[7]);
  }
}

__device__ __forceinline__ void store_group(unsigned char* hi_plane, unsigned char* lo_plane, int id, const float4& v0,
                                            const float4& v1) {
  uint4 h, l;
  split_pack2(v0.x, v0.y, h.x, l.x);
  split_pack2(v0.z, v0.w, h.y, l.y);
  split_pack2(v1.x, v1.y, h.z, l.z);
  split_pack2(v1.z, v1.w, h.w, l.w);
  *reinterpret_cast<uint4*>(hi_plane + id * 16) = h;
  *reinterpret_cast<uint4*>(lo_plane + id * 16) = l;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!tc::mbar_try_wait(bar, parity)) {}
}

template <int LAYOUT>
__global__ void __launch_bounds__(TC_GEMM_THREADS, 2) gemm_tc_kernel(GemmParams p, int vecA, int vecB, int vecA2, int vecB2) {
  constexpr bool A_MN = (LAYOUT == GEMM_TN);   // A given as [K][M]
  constexpr bool B_MN = (LAYOUT != GEMM_NT);   // B given as [K][N]
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BARS);               // [TSTAGES]
  uint64_t* empty = full + TSTAGES;                                            // [TSTAGES]
  uint64_t* accum_full = empty + TSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int w_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;

  const int nk1 = (p.K + TBK - 1) / TBK, nk2 = (p.K2 + TBK - 1) / TBK;
  const int nk_total = nk1 + nk2;
  const int per_split = (nk_total + p.split_k - 1) / p.split_k;
  const int t_begin = blockIdx.z * per_split;
  const int t_end = min(nk_total, t_begin + per_split);
  if (t_begin >= t_end) return;
  const int n_tiles = t_end - t_begin;
  int n_eff = min(TBN, p.N - n0);
  n_eff = (n_eff + 15) & ~15;                   // MMA N: multiple of 16 (rows >= N are zero-filled)

  if (tid == 0) {
    for (int s = 0; s < TSTAGES; ++s) { tc::mbar_init(&full[s], PRODUCER_WARPS); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(accum_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w_u == PRODUCER_WARPS) { __syncwarp(); tc::tmem_alloc(tmem_slot, TBN); }
  tc::fence_before_thread_sync();
  __syncthreads();
  tc::fence_after_thread_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (w_u < PRODUCER_WARPS) {
    // ================= producers =================
    auto src_of = [&](int t, OperandSrc& a, OperandSrc& b) {
      if (t < nk1) {
        a = OperandSrc{p.A, p.lda, m0, p.M, t * TBK, p.K, vecA};
        b = OperandSrc{p.B, p.ldb, n0, p.N, t * TBK, p.K, vecB};
      } else {
        a = OperandSrc{p.A2, p.lda2, m0, p.M, (t - nk1) * TBK, p.K2, vecA2};
        b = OperandSrc{p.B2, p.ldb2, n0, p.N, (t - nk1) * TBK, p.K2, vecB2};
      }
    };
    // raw slot of (stage r, operand op, group id): each thread only ever touches its own four slots, so the raw ring
    // needs no block-level synchronisation - cp.async.wait_group is enough
    auto raw_slot = [&](int r, int op, int id) { return smem + OFF_RAW + r * RAW_STAGE_BYTES + op * (RAW_STAGE_BYTES / 2) + id * 32; };
    auto fetch_tile = [&](int t, int r) {
      OperandSrc a, b;
      src_of(t, a, b);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        fetch_group<A_MN>(a, tid + g * 256, raw_slot(r, 0, tid + g * 256));
        fetch_group<B_MN>(b, tid + g * 256, raw_slot(r, 1, tid + g * 256));
      }
    };
    fetch_tile(t_begin, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (n_tiles > 1) fetch_tile(t_begin + 1, 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i % TSTAGES, r = i % RSTAGES;
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // tile i has landed (tile i+1 may still be in flight)
      float4 ca[2][2], cb[2][2];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const unsigned char* sa = raw_slot(r, 0, tid + g * 256);
        const unsigned char* sb = raw_slot(r, 1, tid + g * 256);
        ca[g][0] = *reinterpret_cast<const float4*>(sa); ca[g][1] = *reinterpret_cast<const float4*>(sa + 16);
        cb[g][0] = *reinterpret_cast<const float4*>(sb); cb[g][1] = *reinterpret_cast<const float4*>(sb + 16);
      }
      if (i + RSTAGES < n_tiles) fetch_tile(t_begin + i + RSTAGES, r);   // the slot is free again: refill it
      asm volatile("cp.async.commit_group;" ::: "memory");
      mbar_wait_spin(&empty[s], ((i / TSTAGES) & 1) ^ 1);    // the MMAs that read this operand stage have retired
      unsigned char* st = smem + s * STAGE_BYTES;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        store_group(st, st + PLANE_BYTES, tid + g * 256, ca[g][0], ca[g][1]);
        store_group(st + 2 * PLANE_BYTES, st + 3 * PLANE_BYTES, tid + g * 256, cb[g][0], cb[g][1]);
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
    // ================= MMA issuer =================
    const uint32_t idesc = tc::make_idesc_bf16_f32(TBM, n_eff) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16);
    // per-operand descriptors (LBO = k-group stride, SBO = 8-row-group stride; see the layout note above)
    const uint64_t da0 = A_MN ? tc::make_smem_desc(tc::smem_u32(smem), 2048, 128) : tc::make_smem_desc(tc::smem_u32(smem), 128, 512);
    const uint64_t db0 = B_MN ? tc::make_smem_desc(tc::smem_u32(smem), 2048, 128) : tc::make_smem_desc(tc::smem_u32(smem), 128, 512);
    constexpr int A_KS = A_MN ? 4096 : 256, B_KS = B_MN ? 4096 : 256;   // byte advance per K=16 step
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i % TSTAGES;
      mbar_wait_spin(&full[s], (i / TSTAGES) & 1);
      __syncwarp();
      tc::fence_after_thread_sync();
      if (tc::elect_one()) {
        const uint64_t dsa = da0 + (uint64_t)((s * STAGE_BYTES) >> 4);
        const uint64_t dsb = db0 + (uint64_t)((s * STAGE_BYTES + 2 * PLANE_BYTES) >> 4);
#pragma unroll
        for (int ks = 0; ks < TBK / 16; ++ks) {
          const uint64_t a_hi = dsa + (uint64_t)((ks * A_KS) >> 4);
          const uint64_t a_lo = a_hi + (uint64_t)(PLANE_BYTES >> 4);
          const uint64_t b_hi = dsb + (uint64_t)((ks * B_KS) >> 4);
          const uint64_t b_lo = b_hi + (uint64_t)(PLANE_BYTES >> 4);
          tc::mma_bf16_ss(tmem_base, a_lo, b_hi, idesc, (i | ks) != 0);
          tc::mma_bf16_ss(tmem_base, a_hi, b_lo, idesc, true);
          tc::mma_bf16_ss(tmem_base, a_hi, b_hi, idesc, true);
        }
        tc::mma_commit(&empty[s]);
        if (i + 1 == n_tiles) tc::mma_commit(accum_full);
      }
      __syncwarp();
    }
  }

  // ================= epilogue: all 8 producer warps; warp w owns TMEM lane quarter w%4 and column half w/4 ========
  if (w_u < PRODUCER_WARPS) {
    // bias tile -> shared (reuses the raw ring, which is idle now), so the per-column adds do not wait on global loads
    float* s_bias = reinterpret_cast<float*>(smem + OFF_RAW);
    if (p.bias) {
      asm volatile("bar.sync 1, 256;" ::: "memory");                  // every producer is done with the raw ring
      if (tid < TBN) s_bias[tid] = (n0 + tid < p.N) ? __ldg(p.bias + n0 + tid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    mbar_wait_spin(accum_full, 0);
    __syncwarp();
    tc::fence_after_thread_sync();
    const int q = w_u & 3, half = w_u >> 2;
    const int row = m0 + q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec_c = ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (p.ldc % 4 == 0);
    const bool vec_z = p.Z && ((reinterpret_cast<uintptr_t>(p.Z) & 15) == 0) && (p.ldz % 4 == 0);
    const int c_begin = half * (TBN / 2), c_end = min(n_eff, c_begin + TBN / 2);
    for (int c0 = c_begin; c0 < c_end; c0 += 8) {
      float v[8];
      __syncwarp();                                          // tcgen05.ld is warp-collective: reconverge first
      tc::tmem_ld_32x32b_x8(lane_base + (uint32_t)c0, v);
      const int col = n0 + c0;
      if (row >= p.M || col >= p.N) continue;
      const int nv = min(8, p.N - col);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += s_bias[c0 + j];
      }
      if (p.epilogue == EPI_TANH) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = tanhf(v[j]);
      } else if (p.epilogue == EPI_MUL_DTANH || p.epilogue == EPI_ADD_Z) {
        const float* z = p.Z + (long long)row * p.ldz + col;
        float zz[8];
        if (vec_z && nv == 8) {
          const float4 z0 = *reinterpret_cast<const float4*>(z), z1 = *reinterpret_cast<const float4*>(z + 4);
          zz[0] = z0.x; zz[1] = z0.y; zz[2] = z0.z; zz[3] = z0.w; zz[4] = z1.x; zz[5] = z1.y; zz[6] = z1.z; zz[7] = z1.w;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) zz[j] = (j < nv) ? z[j] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (p.epilogue == EPI_MUL_DTANH) ? v[j] * (1.f - zz[j] * zz[j]) : v[j] + zz[j];
      }
      float* cp = p.C + (long long)row * p.ldc + col;
      if (p.split_k > 1) {
        if (vec_c && nv == 8) {
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
          asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(cp + 4), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) if (j < nv) atomicAdd(cp + j, v[j]);
        }
      } else if (vec_c && nv == 8) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (j < nv) cp[j] = v[j];
      }
    }
    tc::fence_before_thread_sync();
  }
  __syncthreads();
  if (w_u == PRODUCER_WARPS) { __syncwarp(); tc::tmem_dealloc(tmem_base, TBN); }
}

template <int LAYOUT>
int launch_gemm_tc(const GemmParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    R2D2_CUDA_TRY(cudaFuncSetAttribute(gemm_tc_kernel<LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_GEMM_SMEM));
    attr_set = true;
  }
  auto aligned = [](const float* ptr, long long ld) {
    return ptr != nullptr && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 4 == 0);
  };
  dim3 grid(ceil_div(p.N, TBN), ceil_div(p.M, TBM), p.split_k);
  gemm_tc_kernel<LAYOUT><<<grid, TC_GEMM_THREADS, TC_GEMM_SMEM, stream>>>(
      p, aligned(p.A, p.lda), aligned(p.B, p.ldb), aligned(p.A2, p.lda2), aligned(p.B2, p.ldb2));
  count_launch();
  R2D2_CUDA_TRY(cudaGetLastError());
  return R2D2_OK;
}

}  // namespace

int gemm_tc_suggest_split_k(int M, int N, int K) {
  long long tiles = (long long)ceil_div(M, TBM) * ceil_div(N, TBN);
  int k_tiles = ceil_div(K, TBK);
  if (tiles >= 2 * 148 || k_tiles < 16) return 1;
  int want = (int)ceil_div_ll(2 * 148, tiles);
  int max_by_k = k_tiles / 8;
  int s = want < max_by_k ? want : max_by_k;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return s;
}

int gemm_f32_tc(const GemmParams& p, GemmLayout layout, cudaStream_t stream) {
  switch (layout) {
    case GEMM_NT: return launch_gemm_tc<GEMM_NT>(p, stream);
    case GEMM_NN: return launch_gemm_tc<GEMM_NN>(p, stream);
    case GEMM_TN: return launch_gemm_tc<GEMM_TN>(p, stream);
  }
  set_last_error("bad gemm layout");
  return R2D2_ERR_ARG;
}

}  // namespace r2d2
