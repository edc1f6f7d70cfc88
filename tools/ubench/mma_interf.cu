// Microbenchmark (GPU box): does other SM activity slow tcgen05.mma down?  One CTA per SM, warp 0 lane 0 issues a
// chain of M128 N256 K16 MMAs; warps 4-7 run an interference loop until the issuer is done.
//   modes: 0 none | 1 tcgen05.ld from other TMEM columns | 2 st.shared.v4 to a scratch region
//          3 bulk copies global -> smem scratch (TMA engine) | 4 mbarrier try_wait polling | 5 ld.global streaming
//          6 tcgen05.ld + cvt + st.shared + fence.proxy.async (an epilogue)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace ciaosr::tc;
constexpr uint32_t DESC_HI_ = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t dlo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void mma(uint32_t d, uint32_t a, uint32_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(d), "r"(a), "r"(b), "r"(idesc), "r"(acc), "r"(DESC_HI_) : "memory");
}

__global__ void __launch_bounds__(256, 1) bench(int mode, int iters, const uint4* gsrc, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar_never, bar_tx;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar_never), 1); mbar_init(smem_u32(&bar_tx), 1); done = 0; fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t sb = smem_u32(smem);
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_split(128, 256);
    const uint32_t a0 = dlo(sb), b0 = dlo(sb + 65536);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t ks = j & 3, sl = j >> 2;
        mma(tb, a0 + sl * (16384 >> 4) + ks * 2, b0 + sl * (16384 >> 4) + ks * 2, idesc, (i | j) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 1);
    out[blockIdx.x] = clock64() - t0;
    done = 1;
  } else if (warp >= 4) {
    const uint32_t lane_taddr = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    const uint32_t scratch = sb + 160 * 1024 + (threadIdx.x - 128) * 16;       // 2 KB per pass, 32 KB region
    float acc = 0.f;
    uint32_t ph = 0;
    long long n = 0;
    while (!done) {
      if (mode == 1) {
        float v[32];
        tmem_ld32(lane_taddr + (n & 7) * 32, v);
        acc += v[0] + v[31];
      } else if (mode == 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scratch + q * 2048), "r"(q), "r"(q), "r"(q), "r"(q) : "memory");
      } else if (mode == 3) {
        if (warp == 4) {
          if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(&bar_tx), 4 * 16384);
            for (int q = 0; q < 4; ++q) bulk_g2s(sb + 128 * 1024 + q * 16384, gsrc + (size_t)((n * 4 + q) & 63) * 1024, 16384, smem_u32(&bar_tx));
          }
          mbar_wait(smem_u32(&bar_tx), ph, 2);
          ph ^= 1;
        }
      } else if (mode == 4) {
        acc += mbar_try_wait(smem_u32(&bar_never), 0) ? 1.f : 0.f;
      } else if (mode == 5) {
        const uint4 q = __ldg(gsrc + ((n * 128 + (threadIdx.x - 128)) & 65535));
        acc += __uint_as_float(q.x);
      } else if (mode == 6) {
        float v[32];
        tmem_ld32(lane_taddr + (n & 7) * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + 1.0f, 0.0f);
        a_store32(sb + 128 * 1024, sb + 144 * 1024, threadIdx.x - 128, (int)(n & 1) * 32, v);
        fence_proxy_async();
      }
      ++n;
    }
    if (acc == 123.456f) sink[0] = acc;
    if (lane == 0 && warp == 4) out[148 + blockIdx.x] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  const int iters = 4096 * 8, grid = 148;
  long long* out; float* sink; uint4* gsrc;
  cudaMalloc(&out, 2 * grid * 8); cudaMalloc(&sink, 4); cudaMalloc(&gsrc, 65536 * 16); cudaMemset(gsrc, 1, 65536 * 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"none", "tcgen05.ld (4 warps)", "st.shared.v4 (4 warps)", "bulk copy g->s 64 KB batches", "mbarrier polling (4 warps)",
                         "ld.global (4 warps)", "epilogue: ld+cvt+st.shared+fence.proxy"};
  for (int mode = 0; mode < 7; ++mode) {
    cudaMemset(out, 0, 2 * grid * 8);
    bench<<<grid, 256, 200 * 1024>>>(mode, iters, gsrc, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h[296];
    cudaMemcpy(h, out, 2 * grid * 8, cudaMemcpyDeviceToHost);
    double sum = 0, its = 0;
    for (int i = 0; i < grid; ++i) { sum += h[i]; its += h[148 + i]; }
    printf("mode %d %-42s %7.1f cycles/MMA   interference iterations per MMA %.2f\n", mode, names[mode], sum / grid / iters, its / grid / iters);
  }
  return 0;
}
