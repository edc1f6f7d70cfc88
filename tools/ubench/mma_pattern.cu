// Microbenchmark (GPU box): why do the UMMAs of the real kernels run slower than a bare dependent chain (r02x: 128 cycles
// for N = 256, 74 for N <= 128)?  Replays the issue PATTERNS of the kernels on resident operands:
//   pair   : per K-slab 8 UMMAs (A_lo.W_hi, A_hi.W_hi alternating) + commit, 4 UMMAs (A_hi.W_lo) + 2 commits  (cta_group::2, N = 256)
//   conv   : per tap 8 UMMAs (A in TMEM, B alternating lo / hi) + commit                                       (cta_group::1, N = 128)
// knobs: commits on/off, multicast commits, an mbarrier try_wait + tcgen05.fence::after_thread_sync before every group
// (what the issuer does when it checks W_FULL), whole-warp loop with lane 0 issuing (as the kernels) vs a single thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../ciaosr_b200/csrc -o mma_pattern mma_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ciaosr::tc;

constexpr uint32_t DESC_HI_ = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t dlo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma_ss1(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
}
__device__ __forceinline__ void mma_ts1(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
}
__device__ __forceinline__ void commit1(uint32_t bar) { umma_commit(bar); }
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit2mc(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// flags: 1 commits, 2 multicast commits (pair only), 4 try_wait + fence before every group, 8 whole warp walks the loop
__device__ int g_random_data;
template <int PAT>     // 0 pair (cta_group::2 N=256 SS), 1 conv (cta_group::1 N=128 TS), 2 pair N-split (cta_group::2 N=128 SS)
__global__ void __launch_bounds__(128, 1) bench(int flags, int slabs, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_done, bar_sink[8], bar_ok;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) {
    uint32_t v = 0x3c003c00u;                     // 1.0, 1.0
    if (g_random_data) {                          // two random fp16 in +-[2^-4, 1): random sign, exponent 11..14, random mantissa
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
      const uint32_t lo = (h & 0x83FFu) | ((11u + ((h >> 10) & 3u)) << 10), hi = ((h >> 16) & 0x83FFu) | ((11u + ((h >> 26) & 3u)) << 10);
      v = lo | (hi << 16);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_ok), 1);
    for (int q = 0; q < 8; ++q) mbar_init(smem_u32(&bar_sink[q]), 1 << 20);
    fence_mbar_init();
    mbar_arrive(smem_u32(&bar_ok));            // phase 0 of bar_ok is complete: try_wait(parity 0) succeeds at once
  }
  constexpr int CG = PAT == 1 ? 1 : 2;
  if (warp == 0) {
    if (CG == 1) tmem_alloc(smem_u32(&slot), 512);
    else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = slot;
  const bool leader_cta = CG == 1 || cluster_ctarank() == 0;
  const bool whole_warp = flags & 8;
  if (warp == 0 && leader_cta && (whole_warp || lane == 0)) {
    const bool issue = lane == 0;
    const uint32_t sb = smem_u32(smem);
    const uint32_t idesc = make_idesc_split(128 * CG, PAT == 0 ? 256 : 128);
    const long long t0 = clock64();
    for (int s = 0; s < slabs; ++s) {
      const uint32_t slot4 = s & 3, d = tb + ((s >> 2) & 1) * 256;      // same accumulator for 4 slabs, then the other
      const uint32_t a_hi = dlo(sb + slot4 * 16384), a_lo = dlo(sb + 65536 + slot4 * 16384);
      const uint32_t b_hi = dlo(sb + 131072 + (s & 1) * 32768), b_lo = dlo(sb + 131072 + (s & 1) * 32768 + 16384);
      const uint32_t acc0 = (s & 3) ? 1u : 0u;
      if (flags & 4) { while (!mbar_try_wait(smem_u32(&bar_ok), 0)) {} tc_fence_after(); }
      if (PAT == 0 || PAT == 2) {
        if (issue) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ss2(d, a_lo + 2 * ks, b_hi + 2 * ks, idesc, acc0 | ks);
            mma_ss2(d, a_hi + 2 * ks, b_hi + 2 * ks, idesc, 1u);
          }
          if (flags & 1) { if (flags & 2) commit2mc(smem_u32(&bar_sink[0])); else commit2(smem_u32(&bar_sink[0])); }
        }
        if (whole_warp) __syncwarp();
        if (flags & 4) { while (!mbar_try_wait(smem_u32(&bar_ok), 0)) {} tc_fence_after(); }
        if (issue) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) mma_ss2(d, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
          if (flags & 1) {
            if (flags & 2) { commit2mc(smem_u32(&bar_sink[1])); commit2mc(smem_u32(&bar_sink[2])); }
            else { commit2(smem_u32(&bar_sink[1])); commit2(smem_u32(&bar_sink[2])); }
          }
        }
        if (whole_warp) __syncwarp();
      } else {
        const uint32_t a = tb + 256 + (s & 7) * 32;
        if (issue) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_ts1(d, a + 8 * ks, b_lo + 2 * ks, idesc, acc0 | ks);
            mma_ts1(d, a + 8 * ks, b_hi + 2 * ks, idesc, 1u);
          }
          if (flags & 1) commit1(smem_u32(&bar_sink[0]));
        }
        if (whole_warp) __syncwarp();
      }
    }
    if (issue) {
      if (CG == 1) commit1(smem_u32(&bar_done)); else commit2(smem_u32(&bar_done));
      mbar_wait(smem_u32(&bar_done), 0, 1);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 1) tmem_dealloc(tb, 512);
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

template <int PAT>
void run(const char* name, int flags) {
  const int grid = 148, slabs = 4096 * 16;
  constexpr int CG = PAT == 1 ? 1 : 2;
  long long* out;
  cudaMalloc(&out, grid * 8);
  cudaMemset(out, 0, grid * 8);
  auto k = bench<PAT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, flags, slabs, out);
    if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  double sum = 0; int n = 0;
  for (int i = 0; i < grid; ++i) if (h[i]) { sum += h[i]; ++n; }
  const int per_slab = PAT == 1 ? 8 : 12;
  printf("%-28s flags=%2d (%s%s%s%s): %7.1f cycles/MMA  (%7.0f per K-slab / tap)\n", name, flags, flags & 1 ? "commits " : "",
         flags & 2 ? "multicast " : "", flags & 4 ? "wait+fence " : "", flags & 8 ? "whole-warp" : "", sum / n / slabs / per_slab, sum / n / slabs);
  cudaFree(out);
}

int main() {
  for (int rnd = 0; rnd < 2; ++rnd) {
    cudaMemcpyToSymbol(g_random_data, &rnd, 4);
    printf("== operands in shared memory: %s\n", rnd ? "random fp16 in +-[2^-4, 1)" : "constant 1.0");
    for (int flags : {0, 15}) run<0>("pair N=256 (12 per slab)", flags);
    for (int flags : {0, 15}) run<2>("pair N=128 (12 per unit)", flags);
    for (int flags : {0, 13}) run<1>("conv N=128 A=tmem (8 per tap)", flags);
  }
  return 0;
}
