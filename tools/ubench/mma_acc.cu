// Microbenchmark (GPU box): how tcgen05.mma accumulates into its fp32 TMEM accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../ciaosr_b200/csrc -o mma_acc mma_acc.cu
// Row r of A carries the probe x_r = f_r * 2^-11, B column 0 carries y = 2^-12, so one product is f_r ulp(1)
// (ulp(1) = 2^-23).  Test 1: D = 1, then T separate MMAs each adding one product  -> rounding of the
// accumulate step between instructions.  Test 2: ONE MMA whose K = 16 terms are 1 and 15 products -> rounding
// inside an instruction.  Printed in units of ulp(1) above 1.0 next to the exact value.
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

__global__ void __launch_bounds__(128, 1) probe(int T, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  // slabs: A1 (ones in k=0), A2 (x_r in k=0), A3 (1 in k=0, x_r in k=1..15), B (y.. see below)
  __half* A1 = reinterpret_cast<__half*>(smem);
  __half* A2 = reinterpret_cast<__half*>(smem + SLAB_BYTES);
  __half* A3 = reinterpret_cast<__half*>(smem + 2 * SLAB_BYTES);
  __half* B1 = reinterpret_cast<__half*>(smem + 3 * SLAB_BYTES);     // column n=0: [1, 0, ...]
  __half* B2 = reinterpret_cast<__half*>(smem + 4 * SLAB_BYTES);     // column n=0: [y, 0, ...]
  __half* B3 = reinterpret_cast<__half*>(smem + 5 * SLAB_BYTES);     // column n=0: [1, y, y, ...]
  __half* B4 = reinterpret_cast<__half*>(smem + 6 * SLAB_BYTES);     // column n=0: [0, y, y, ...]
  for (int i = threadIdx.x; i < 7 * SLAB_BYTES / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  const int r = threadIdx.x;
  const float f = 0.0625f * (float)(r % 16 + 1);           // 1/16 .. 1 ulp per product
  const float x = f * exp2f(-11.0f), y = exp2f(-12.0f);
  auto at = [](__half* s, int n, int k) -> __half& { return *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(s) + sw128_offset(n, k)); };
  at(A1, r, 0) = __float2half(1.0f);
  at(A2, r, 0) = __float2half(x);
  at(A3, r, 0) = __float2half(1.0f);
  for (int k = 1; k < 16; ++k) at(A3, r, k) = __float2half(x);
  if (r == 0) {
    at(B1, 0, 0) = __float2half(1.0f);
    at(B2, 0, 0) = __float2half(y);
    at(B3, 0, 0) = __float2half(1.0f);
    for (int k = 1; k < 16; ++k) { at(B3, 0, k) = __float2half(y); at(B4, 0, k) = __float2half(y); }
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if ((threadIdx.x >> 5) == 0) tmem_alloc(smem_u32(&slot), 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_split(128, 16);
    const uint32_t sb = smem_u32(smem);
    // test 1 -> columns [0,16)
    mma(tb, dlo(sb), dlo(sb + 3 * SLAB_BYTES), idesc, 0);
    for (int t = 0; t < T; ++t) mma(tb, dlo(sb + SLAB_BYTES), dlo(sb + 4 * SLAB_BYTES), idesc, 1);
    // test 2 -> columns [16,32)
    mma(tb + 16, dlo(sb + 2 * SLAB_BYTES), dlo(sb + 5 * SLAB_BYTES), idesc, 0);
    // test 3 -> columns [32,48): the 15 small products first (one MMA), then +1 in a second MMA
    mma(tb + 32, dlo(sb + 2 * SLAB_BYTES), dlo(sb + 6 * SLAB_BYTES), idesc, 0);
    mma(tb + 32, dlo(sb), dlo(sb + 3 * SLAB_BYTES), idesc, 1);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0, 1);
  tc_fence_after();
  float v[32];
  tmem_ld32(tb + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
  out[r * 4 + 0] = v[0]; out[r * 4 + 1] = v[16];
  float w[32];
  tmem_ld32(tb + 32 + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), w);
  out[r * 4 + 2] = w[0]; out[r * 4 + 3] = f;
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc(tb, 64);
}

int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 16;
  float* out;
  cudaMalloc(&out, 128 * 4 * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * SLAB_BYTES);
  probe<<<1, 128, 7 * SLAB_BYTES>>>(T, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  float h[128 * 4];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const double ulp = 1.1920928955078125e-07;
  printf("f(ulp/product)  T=%d adds: got / exact   one K16 MMA (1 + 15 products): got / exact   (15 products) then +1: got / exact 2-step\n", T);
  for (int r = 0; r < 16; ++r) {
    const double f = h[r * 4 + 3];
    printf("%6.4f   %8.3f / %8.3f     %8.3f / %8.3f     %8.3f / %8.3f\n", f, (h[r * 4] - 1.0) / ulp, T * f,
           (h[r * 4 + 1] - 1.0) / ulp, 15 * f, (h[r * 4 + 2] - 1.0) / ulp, 15 * f + 0.0);
  }
  return 0;
}
