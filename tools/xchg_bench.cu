// Micro-benchmark (dev tool, not product code): cost of the per-step h_t all-gather inside a cluster of 16 CTAs,
// (A) as bulk shared->remote-shared copies (DSMEM, what the scan kernels do) versus (B) through L2: bulk store of the
// CTA's slice to global memory, wait for the writes, then ONE multicast bulk load that lands in all 16 CTAs.
// Prints cycles per round for RG = 4 and 8 row groups (32 / 64 batch rows: 64 / 128 KB gathered per CTA per round)
// and checks the payload in mode B (visibility of the bulk store to the multicast load that follows it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/xchg_bench tools/xchg_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t par) {
  const long long t0 = clock64();
  while (!mbar_try(b, par)) if (clock64() - t0 > 2000000000LL) return false;
  return true;
}

constexpr int C = 16, SLICE = 1024, THREADS = 512;

struct Params { int mode, rg, rounds; unsigned char* scratch; long long* cycles; int* errors; };

// smem: stage [RG][1 KB], recv [2][RG][C][1 KB] (mode A double-buffered like the scan), bars
__global__ void __launch_bounds__(THREADS, 1) xchg_kernel(Params p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), cl = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  extern __shared__ __align__(128) unsigned char smem[];
  const int RG = p.rg;
  unsigned char* stage = smem;                     // RG KB
  unsigned char* recv = smem + 8 * SLICE;          // [RG][C][1 KB] (single buffer; payload races do not matter for mode A timing)
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 8 * SLICE + 8 * C * SLICE);   // [2]
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  cluster.sync();
  const uint32_t tx = p.mode == 3 ? (uint32_t)(C * 16) : (uint32_t)(RG * C * SLICE);
  long long t_begin = 0;
  int bad = 0;
  for (int r = 0; r < p.rounds; ++r) {
    if (r == 8 && tid == 0) t_begin = clock64();
    const int par = r & 1;
    // "cell math": every thread writes its part of the slice (payload = round, source rank, word index)
    for (int i = tid; i < RG * SLICE / 4; i += THREADS) reinterpret_cast<uint32_t*>(stage)[i] = ((uint32_t)r << 16) | ((uint32_t)rank << 12) | (uint32_t)(i & 0xFFF);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) mbar_expect(&full[par], tx);
    __syncthreads();
    if (p.mode == 3) {   // fixed part of a round: 16-byte copies only
      if (lane == 0 && w < C)
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(smem_u32(recv + rank * 16), w)),
                     "r"(smem_u32(stage)), "r"(16u), "r"(mapa(smem_u32(&full[par]), w)) : "memory");
    } else if (p.mode == 0) {
      // (A) DSMEM: warp w -> destination w, one 1 KB copy per row group
      if (lane == 0 && w < C) {
        for (int g = 0; g < RG; ++g) {
          const uint32_t dst_local = smem_u32(recv + (g * C + rank) * SLICE);
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mapa(dst_local, w)),
                       "r"(smem_u32(stage + g * SLICE)), "r"((uint32_t)SLICE), "r"(mapa(smem_u32(&full[par]), w)) : "memory");
        }
      }
    } else {
      // (B) through L2: warp g (< RG) stores row group g (1 KB) to global, waits for the write, multicasts it back
      unsigned char* gbase = p.scratch + ((size_t)(cl * 2 + par) * C + rank) * 8 * SLICE;
      if (p.mode == 1) {
        if (lane == 0 && w < RG) {
          unsigned char* gdst = gbase + w * SLICE;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(stage + w * SLICE)), "r"((uint32_t)SLICE) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          const uint32_t dst_local = smem_u32(recv + (w * C + rank) * SLICE);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_local),
                       "l"(gdst), "r"((uint32_t)SLICE), "r"(smem_u32(&full[par])), "h"((unsigned short)0xFFFF) : "memory");
        }
      } else {   // mode 2: one thread moves the whole slice (RG KB) in one store + one multicast load; recv layout [C][RG KB]
        if (tid == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gbase), "r"(smem_u32(stage)), "r"((uint32_t)(RG * SLICE)) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          const uint32_t dst_local = smem_u32(recv + rank * RG * SLICE);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst_local),
                       "l"(gbase), "r"((uint32_t)(RG * SLICE)), "r"(smem_u32(&full[par])), "h"((unsigned short)0xFFFF) : "memory");
        }
      }
    }
    if (!mbar_wait(&full[par], (r >> 1) & 1)) { if (tid == 0) atomicAdd(p.errors, 1000000); break; }
    if (p.mode == 1 || p.mode == 2) {   // payload check: a few words of every source's slice
      for (int i = tid; i < RG * C; i += THREADS) {
        const int g = i / C, src = i % C;
        const uint32_t* wv = reinterpret_cast<const uint32_t*>(p.mode == 1 ? recv + (g * C + src) * SLICE : recv + src * RG * SLICE + g * SLICE);
        const uint32_t idx = (uint32_t)(g * SLICE / 4 + (r % 200));
        const uint32_t want = ((uint32_t)r << 16) | ((uint32_t)src << 12) | (idx & 0xFFF);
        if (wv[r % 200] != want) ++bad;
      }
    }
    __syncthreads();   // recv has been read; in lock step with the other CTAs through the next round's data dependency
    cluster.sync();    // keep the payload check honest (no overwrite while a slow CTA still reads)
  }
  if (tid == 0) p.cycles[blockIdx.x] = clock64() - t_begin;
  if (bad) atomicAdd(p.errors, bad);
  cluster.sync();
}

int main() {
  const int n_clusters = 8, rounds = 208;
  unsigned char* scratch; long long* cycles; int* errors;
  CK(cudaMalloc(&scratch, (size_t)n_clusters * 2 * C * 8 * SLICE));
  CK(cudaMalloc(&cycles, sizeof(long long) * n_clusters * C));
  CK(cudaMalloc(&errors, sizeof(int)));
  const int smem_bytes = 8 * SLICE + 8 * C * SLICE + 64;
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C * n_clusters); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&max_clusters, xchg_kernel, &cfg));
  printf("resident clusters of 16 (%d B smem): %d\n", smem_bytes, max_clusters);
  for (int mode = 0; mode < 4; ++mode)
    for (int rg = 4; rg <= 8; rg += 4) {
      Params p = {mode, rg, rounds, scratch, cycles, errors};
      CK(cudaMemset(errors, 0, sizeof(int)));
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaLaunchKernelEx(&cfg, xchg_kernel, p));
        CK(cudaDeviceSynchronize());
      }
      long long h[n_clusters * C]; int herr = 0;
      CK(cudaMemcpy(h, cycles, sizeof(h), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&herr, errors, sizeof(int), cudaMemcpyDeviceToHost));
      long long mx = 0, mn = 1LL << 62;
      for (int i = 0; i < n_clusters * C; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
      const double per = (double)mx / (rounds - 8);
      printf("mode %d (%s) rows %d: %.0f cycles/round (min CTA %.0f) incl. ~cluster.sync + fill; %d KB gathered per CTA -> %.1f B/clk/SM; errors %d\n", mode,
             mode == 0 ? "DSMEM bulk copies" : (mode == 1 ? "L2: 1 KB store+multicast per row group" : (mode == 2 ? "L2: one store+multicast per CTA" : "fixed part: 16 B copies")), rg * 8, per,
             (double)mn / (rounds - 8), rg * C, rg * C * 1024.0 / per, herr);
    }
  // baseline: the same loop with no exchange cannot be expressed (the barrier would never complete); report the fixed part
  return 0;
}
