// Generic fp32 CUDA-core GEMM with functor operands: C(m,n) = epi(sum_k A(m,k) B(k,n)).
//
// This is the any-shape path (CIAOSR_ENGINE_SIMT) and the on-device fp32
// cross-check of the tcgen05 path.  Operands are functors so that the
// reference's unfold / patch-extraction / gather steps (ciaosr_net.py:131-139,
// arch_csnln.py:59-87) are never materialised: the loaders index the NHWC
// feature maps directly (implicit im2col).
#pragma once
#include "common.cuh"

namespace ciaosr {

constexpr int GBM = 128, GBN = 128, GBK = 16, GTHREADS = 256;

struct RowMajorA {            // A[m, k] = p[m * ld + k]
  const float* p; long long ld;
  __device__ __forceinline__ float operator()(int m, int k) const { return p[(long long)m * ld + k]; }
};
struct RowMajorB {            // B[k, n] = p[k * ld + n]
  const float* p; long long ld;
  __device__ __forceinline__ float operator()(int k, int n) const { return p[(long long)k * ld + n]; }
};

struct EpiBiasAct {           // C[m, n] = act(acc + bias[n])
  float* c; long long ldc; const float* bias; int relu;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    float v = acc + (bias ? bias[n] : 0.0f);
    if (relu) v = fmaxf(v, 0.0f);
    c[(long long)m * ldc + n] = v;
  }
};

template <class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(GTHREADS)
gemm_simt_kernel(int M, int N, int K, ALoad A, BLoad Bm, Epi epi) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int ty = tid / 16, tx = tid % 16;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += GBK) {
    // A tile: thread -> (k = tid % 16, m = tid / 16 + 16 j)
    {
      const int k = tid % GBK, kk = k0 + k;
#pragma unroll
      for (int j = 0; j < GBM / 16; ++j) {
        const int m = tid / GBK + 16 * j, mm = m0 + m;
        As[k][m] = (mm < M && kk < K) ? A(mm, kk) : 0.0f;
      }
    }
    // B tile: thread -> (n = tid % 128, k = tid / 128 + 2 j)
    {
      const int n = tid % GBN, nn = n0 + n;
#pragma unroll
      for (int j = 0; j < GBK / 2; ++j) {
        const int k = tid / GBN + 2 * j, kk = k0 + k;
        Bs[k][n] = (nn < N && kk < K) ? Bm(kk, nn) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) epi(m, n, acc[i][j]);
    }
  }
}

template <class ALoad, class BLoad, class Epi>
static int gemm_simt(int M, int N, int K, ALoad A, BLoad Bm, Epi epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return CIAOSR_OK;
  dim3 grid(cdiv(M, GBM), cdiv(N, GBN));
  CIAOSR_LAUNCH((gemm_simt_kernel<ALoad, BLoad, Epi>), grid, GTHREADS, 0, st, M, N, K, A, Bm, epi);
  return CIAOSR_OK;
}

}  // namespace ciaosr
