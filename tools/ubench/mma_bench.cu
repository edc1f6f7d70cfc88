// Microbenchmark (GPU box): cycles per tcgen05.mma for the shapes the kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../ciaosr_b200/csrc -o mma_bench mma_bench.cu
// Every SM runs one CTA (or CTA pair) that issues a dependent chain of MMAs on resident operands.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ciaosr::tc;

constexpr uint32_t DESC_HI_ = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t dlo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }

template <int CG, bool ATMEM>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1) {
    if (ATMEM)
      asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
    else
      asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
  } else {
    if (ATMEM)
      asm volatile("{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 db, {%2, %5};\n"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
    else
      asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
  }
}

template <int CG, bool ATMEM, int CE>
__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int nd, long long* out) {
  constexpr int commit_every = CE;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int q = 0; q < 4; ++q) mbar_init(smem_u32(&bar2[q]), 1); fence_mbar_init(); }
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
  const bool leader = CG == 1 || cluster_ctarank() == 0;
  long long dt = 0;
  if (threadIdx.x == 0 && leader) {
    const uint32_t idesc = make_idesc_split(128 * CG, N);
    const uint32_t sb = smem_u32(smem);
    const long long t0 = clock64();
    const uint32_t a0 = ATMEM ? tb + 256 : dlo(sb), b0 = dlo(sb + 65536);
    const uint32_t astep_ks = ATMEM ? 8 : 2, astep_sl = ATMEM ? 32 : (16384 >> 4);
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t ks = j & 3, sl = j >> 2;
        mma<CG, ATMEM>(tb + (nd > 1 ? (uint32_t)(j % nd) * (uint32_t)(ATMEM ? 128 : N) : 0u), a0 + sl * astep_sl + ks * astep_ks,
                       b0 + sl * (16384 >> 4) + ks * 2, idesc, (i | j) >= nd ? 1u : 0u);
        if (commit_every && ((j + 1) % commit_every) == 0) {
          if (CG == 1) umma_commit(smem_u32(&bar2[(j / commit_every) & 3]));
          else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[(j / commit_every) & 3])) : "memory");
        }
      }
    }
    if (CG == 1) umma_commit(smem_u32(&bar));
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0, 1);
    dt = clock64() - t0;
    out[blockIdx.x] = dt;
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

template <int CG, bool ATMEM, int CE = 0>
void run(const char* name, int N, int nd, int iters = 4096) {
  const int commit_every = CE;
  const int grid = 148;
  long long* out;
  cudaMalloc(&out, grid * 8);
  cudaMemset(out, 0, grid * 8);
  auto k = bench<CG, ATMEM, CE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, N, iters, nd, out);
    cudaEventRecord(e1);
    if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
  }
  long long h[148];
  cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; double sum = 0; int n = 0;
  for (int i = 0; i < grid; ++i) if (h[i]) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; ++n; }
  const double per = sum / n / iters;
  const double macs = 128.0 * CG * N * 16 / per / CG;     // per SM per cycle
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  printf("[%8.3f ms, %5.0f MHz] ", ms, sum / n / ms / 1e3);
  printf("%-34s N=%3d nd=%d ce=%d: %7.1f cycles/MMA (max %7.1f)  %6.0f MAC/cycle/SM\n", name, N, nd, commit_every, per, (double)mx / iters, macs);
  cudaFree(out);
}

int main() {
  for (int N : {64, 128, 256}) {
    run<1, false>("cta_group::1 A=smem", N, 1);
    run<1, true>("cta_group::1 A=tmem", N, 1);
    run<2, false>("cta_group::2 A=smem (M=256)", N, 1);
    run<2, true>("cta_group::2 A=tmem (M=256)", N, 1);
  }
  run<1, false, 16>("cta_group::1 A=smem commits", 256, 1);
  run<1, false, 8>("cta_group::1 A=smem commits", 256, 1);
  run<1, false, 4>("cta_group::1 A=smem commits", 256, 1);
  run<1, false, 2>("cta_group::1 A=smem commits", 256, 1);
  run<1, false, 1>("cta_group::1 A=smem commits", 256, 1);
  run<1, true, 8>("cta_group::1 A=tmem commits", 128, 1);
  run<1, true, 4>("cta_group::1 A=tmem commits", 128, 1);
  run<1, true, 2>("cta_group::1 A=tmem commits", 128, 1);
  run<1, false>("LONG cta_group::1 A=smem", 256, 1, 4096 * 1000);
  run<1, false>("LONG cta_group::1 A=smem", 256, 1, 4096 * 1000);
  run<1, true>("LONG cta_group::1 A=tmem", 128, 1, 4096 * 2000);
  run<1, false>("cta_group::1 A=smem alt-D", 128, 2);
  run<1, false>("cta_group::1 A=smem alt-D", 64, 4);
  run<1, false>("cta_group::1 A=smem alt-D", 256, 2);
  run<1, false>("cta_group::1 A=smem alt-D", 128, 4);
  run<1, true>("cta_group::1 A=tmem alt-D", 128, 2);
  run<2, false>("cta_group::2 A=smem alt-D", 256, 2);
  run<2, false>("cta_group::2 A=smem alt-D", 128, 2);
  run<2, false>("cta_group::2 A=smem alt-D", 128, 4);
  run<2, false>("cta_group::2 A=smem alt-D", 64, 4);
  for (int N : {32, 64, 96, 128, 160, 192, 224, 256}) run<2, false>("cta_group::2 A=smem sweep", N, 1);
  for (int N : {32, 64, 96, 128, 192, 256}) run<1, false>("cta_group::1 A=smem sweep", N, 1);
  return 0;
}
