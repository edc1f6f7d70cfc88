// Cross-scale non-local attention (arch_csnln.py:430-532), CUDA-core fp32 path.
//
// Attention form of the reference's conv2d / softmax / conv_transpose2d chain
// (multi_scale = [2]):
//   E  = PReLU(conv1x1_assembly(xp))   [Hp*Wp, C]      xp = reflect-pad(x) to even size
//   Mi = PReLU(conv1x1_match_1(xp))    [Hp*Wp, C/2]
//   R  = PReLU(conv1x1_match_2(avgpool2(xp)))  [L, C/2],  L = Hp*Wp/4
//   S  = 10 * Q K^T,   Q = 3x3 zero-padded patches of Mi, K = 3x3 patches of R / max(|.|, 1e-4)
//   P  = softmax_L(S)
//   O  = P V,          V[l] = 6x6 stride-2 patch of E around (2ly, 2lx), zero padded by 2
//   canvas = overlap-add of O at (2y-2, 2x-2)          [2Hp, 2Wp, C]
//   out = (conv3x3 stride 2 (canvas) + b) / 6, cropped to H x W
// None of the patch tensors is materialised: the GEMM loaders index the NHWC
// maps directly.  S (per row chunk) and O (per image) do live in the workspace.
#include "gemm_simt.cuh"
#include "plan.cuh"
#include "kernels.cuh"

namespace ciaosr {

// ---- layout kernels ---------------------------------------------------------
// dst[b][c][r] = src[b][r][c]  (rows x cols -> cols x rows), 32x32 smem tiles
__global__ void transpose_batched_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                         int rows, int cols) {
  __shared__ float tile[32][33];
  const long long base = (long long)blockIdx.z * rows * cols;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[base + (long long)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[base + (long long)c * rows + r] = tile[threadIdx.x][i];
  }
}

int transpose_batched(const float* src, float* dst, int batch, int rows, int cols, cudaStream_t st) {
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32), batch), block(32, 8);
  CIAOSR_LAUNCH(transpose_batched_kernel, grid, block, 0, st, src, dst, rows, cols);
  return CIAOSR_OK;
}

// ---- loaders ------------------------------------------------------------------
struct PadFeatA {      // A[(y,x) in padded image, ci] of one image, reflect pad bottom/right
  const float* f; int H, W, Wp, C;
  __device__ __forceinline__ float operator()(int m, int k) const {
    int y = m / Wp, x = m % Wp;
    if (y >= H) y = 2 * (H - 1) - y;
    if (x >= W) x = 2 * (W - 1) - x;
    return f[((long long)y * W + x) * C + k];
  }
};
struct PoolFeatA {     // 2x2 average of the padded image (bilinear x0.5 on an even size)
  const float* f; int H, W, Wl, C;
  __device__ __forceinline__ float at(int y, int x, int k) const {
    if (y >= H) y = 2 * (H - 1) - y;
    if (x >= W) x = 2 * (W - 1) - x;
    return f[((long long)y * W + x) * C + k];
  }
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int y = 2 * (m / Wl), x = 2 * (m % Wl);
    const float r0 = at(y, x, k) * 0.5f + at(y + 1, x, k) * 0.5f;
    const float r1 = at(y, x + 1, k) * 0.5f + at(y + 1, x + 1, k) * 0.5f;
    return r0 * 0.5f + r1 * 0.5f;
  }
};
struct EpiBiasPrelu {
  float* c; int ldc; const float* bias; const float* slope;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const float v = acc + bias[n];
    c[(long long)m * ldc + n] = v >= 0.0f ? v : v * slope[0];
  }
};
struct QPatchA {       // 3x3 zero-padded patch of Mi; k = t*Ch + c
  const float* mi; int Hp, Wp, Ch, row0;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int p = row0 + m, t = k / Ch, c = k % Ch;
    const int y = p / Wp + t / 3 - 1, x = p % Wp + t % 3 - 1;
    return (y >= 0 && y < Hp && x >= 0 && x < Wp) ? mi[((long long)y * Wp + x) * Ch + c] : 0.0f;
  }
};
struct KPatchB {       // normalised 3x3 patch of R; k = t*Ch + c, n = l
  const float* r; const float* nrm; int Hl, Wl, Ch;
  __device__ __forceinline__ float operator()(int k, int n) const {
    const int t = k / Ch, c = k % Ch;
    const int y = n / Wl + t / 3 - 1, x = n % Wl + t % 3 - 1;
    return (y >= 0 && y < Hl && x >= 0 && x < Wl)
               ? __fdiv_rn(r[((long long)y * Wl + x) * Ch + c], nrm[n]) : 0.0f;
  }
};
struct EpiScale {
  float* c; long long ldc; float s;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    c[(long long)m * ldc + n] = acc * s;
  }
};
struct VPatchB {       // 6x6 stride-2 patch of E, zero padded by 2; k = l, n = (i*6+j)*C + c
  const float* e; int Hp, Wp, Wl, C;
  __device__ __forceinline__ float operator()(int k, int n) const {
    const int ij = n / C, c = n % C;
    const int y = 2 * (k / Wl) - 2 + ij / 6, x = 2 * (k % Wl) - 2 + ij % 6;
    return (y >= 0 && y < Hp && x >= 0 && x < Wp) ? e[((long long)y * Wp + x) * C + c] : 0.0f;
  }
};
struct DownA {         // 3x3 stride-2 pad-1 patch of the canvas; m = y*W + x (cropped), k = (u*3+v)*C + ci
  const float* cv; int W, H2, W2, C;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int uv = k / C, ci = k % C;
    const int y = 2 * (m / W) - 1 + uv / 3, x = 2 * (m % W) - 1 + uv % 3;
    return (y >= 0 && y < H2 && x >= 0 && x < W2) ? cv[((long long)y * W2 + x) * C + ci] : 0.0f;
  }
};
struct EpiDown {       // (acc + b) / 6 -> NHWC slice (stride ldo) and optionally NCHW
  float* o_nhwc; int ldo; float* o_nchw; int HW; const float* bias;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const float v = __fdiv_rn(acc + bias[n], 6.0f);
    if (o_nhwc) o_nhwc[(long long)m * ldo + n] = v;
    if (o_nchw) o_nchw[(long long)n * HW + m] = v;
  }
};

// ---- small kernels ------------------------------------------------------------
__global__ void csa_knorm_kernel(const float* __restrict__ r, float* __restrict__ nrm, int Hl,
                                 int Wl, int Ch, const float* __restrict__ scalars) {
  const int l = blockIdx.x;
  float ss = 0.0f;
  for (int i = threadIdx.x; i < 9 * Ch; i += blockDim.x) {
    const int t = i / Ch, c = i % Ch;
    const int y = l / Wl + t / 3 - 1, x = l % Wl + t % 3 - 1;
    if (y >= 0 && y < Hl && x >= 0 && x < Wl) {
      const float v = r[((long long)y * Wl + x) * Ch + c];
      ss = fmaf(v, v, ss);
    }
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    ss = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (threadIdx.x == 0) nrm[l] = fmaxf(sqrtf(ss), scalars[3]);
  }
}

__global__ void softmax_rows_kernel(float* __restrict__ s, int L) {
  float* row = s + (long long)blockIdx.x * L;
  __shared__ float red[32];
  __shared__ float bc;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) mx = fmaxf(mx, row[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x < 32) {
    mx = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) bc = mx;
  }
  __syncthreads();
  mx = bc;
  float sum = 0.0f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float e = expf(row[i] - mx);
    row[i] = e;
    sum += e;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x < 32) {
    sum = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (threadIdx.x == 0) bc = sum;
  }
  __syncthreads();
  sum = bc;
  for (int i = threadIdx.x; i < L; i += blockDim.x) row[i] = __fdiv_rn(row[i], sum);
}

// canvas[(Y,X), c] = sum_{a,b} O[(Y/2+1-a, X/2+1-b), ((Y%2+2a)*6 + X%2+2b)*C + c]
__global__ void csa_fold_kernel(const float* __restrict__ o, float* __restrict__ cv, int Hp, int Wp,
                                int C) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = 4LL * Hp * Wp * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int W2 = 2 * Wp;
  const int Y = (int)(p / W2), X = (int)(p % W2);
  float acc = 0.0f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int y = Y / 2 + 1 - a;
    if (y < 0 || y >= Hp) continue;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int x = X / 2 + 1 - b;
      if (x < 0 || x >= Wp) continue;
      const int ij = (Y % 2 + 2 * a) * 6 + (X % 2 + 2 * b);
      acc += o[((long long)y * Wp + x) * (36LL * C) + (long long)ij * C + c];
    }
  }
  cv[i] = acc;
}

// ---- host orchestration ---------------------------------------------------------
struct CsaSizes { int Hp, Wp, Hl, Wl, L, rows_chunk; };
static CsaSizes csa_sizes(int H, int W) {
  CsaSizes s;
  s.Hp = H + (H & 1); s.Wp = W + (W & 1);
  s.Hl = s.Hp / 2; s.Wl = s.Wp / 2; s.L = s.Hl * s.Wl;
  long long rows = (64LL << 20) / s.L;              // <= 256 MB of scores at a time
  if (rows < 128) rows = 128;
  if (rows > (long long)s.Hp * s.Wp) rows = (long long)s.Hp * s.Wp;
  s.rows_chunk = (int)rows;
  return s;
}

size_t cs_attn_workspace(const PlanLayout& L, int H, int W) {
  const CsaSizes s = csa_sizes(H, W);
  Arena a(nullptr, 0);
  const int C = L.C, Ch = C / 2;
  a.take<float>((size_t)s.Hp * s.Wp * C);        // E
  a.take<float>((size_t)s.Hp * s.Wp * Ch);       // Mi
  a.take<float>((size_t)s.L * Ch);               // R
  a.take<float>((size_t)s.L);                    // norms
  a.take<float>((size_t)s.rows_chunk * s.L);     // S chunk
  a.take<float>((size_t)s.Hp * s.Wp * 36 * C);   // O
  a.take<float>((size_t)4 * s.Hp * s.Wp * C);    // canvas
  return a.used();
}

// featT [B,H,W,C] NHWC -> out_nhwc [B,H,W,ldo] (first C channels written) and/or out_nchw [B,C,H,W]
int run_cs_attn(const PlanLayout& L, const float* plan, const float* featT, int B, int H, int W,
                float* out_nhwc, int ldo, float* out_nchw, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  CIAOSR_REQUIRE(H >= 2 && W >= 2, CIAOSR_E_INVALID,
                 "cross-scale attention needs H, W >= 2 (reflect padding), got %dx%d", H, W);
  const CsaSizes s = csa_sizes(H, W);
  const int C = L.C, Ch = C / 2;
  Arena a(ws, ws_bytes);
  float* E = a.take<float>((size_t)s.Hp * s.Wp * C);
  float* Mi = a.take<float>((size_t)s.Hp * s.Wp * Ch);
  float* R = a.take<float>((size_t)s.L * Ch);
  float* nrm = a.take<float>((size_t)s.L);
  float* S = a.take<float>((size_t)s.rows_chunk * s.L);
  float* O = a.take<float>((size_t)s.Hp * s.Wp * 36 * C);
  float* cv = a.take<float>((size_t)4 * s.Hp * s.Wp * C);
  CIAOSR_REQUIRE(a.ok, CIAOSR_E_WORKSPACE, "cross-scale attention workspace too small: need %zu, have %zu",
                 a.used(), ws_bytes);
  const float* scal = plan + L.scalars;
  const int HWp = s.Hp * s.Wp;
  int rc;
  for (int b = 0; b < B; ++b) {
    const float* f = featT + (size_t)b * H * W * C;
    PadFeatA pa{f, H, W, s.Wp, C};
    if ((rc = gemm_simt(HWp, C, C, pa, RowMajorB{plan + L.as_wt, C},
                        EpiBiasPrelu{E, C, plan + L.as_b, scal + 2}, st))) return rc;
    if ((rc = gemm_simt(HWp, Ch, C, pa, RowMajorB{plan + L.m1_wt, Ch},
                        EpiBiasPrelu{Mi, Ch, plan + L.m1_b, scal + 0}, st))) return rc;
    if ((rc = gemm_simt(s.L, Ch, C, PoolFeatA{f, H, W, s.Wl, C}, RowMajorB{plan + L.m2_wt, Ch},
                        EpiBiasPrelu{R, Ch, plan + L.m2_b, scal + 1}, st))) return rc;
    CIAOSR_LAUNCH(csa_knorm_kernel, s.L, 128, 0, st, R, nrm, s.Hl, s.Wl, Ch, scal);
    for (int r0 = 0; r0 < HWp; r0 += s.rows_chunk) {
      const int rows = min(s.rows_chunk, HWp - r0);
      if ((rc = gemm_simt(rows, s.L, 9 * Ch, QPatchA{Mi, s.Hp, s.Wp, Ch, r0},
                          KPatchB{R, nrm, s.Hl, s.Wl, Ch}, EpiScale{S, s.L, L.cs_softmax_scale}, st)))
        return rc;
      CIAOSR_LAUNCH(softmax_rows_kernel, rows, 256, 0, st, S, s.L);
      if ((rc = gemm_simt(rows, 36 * C, s.L, RowMajorA{S, s.L}, VPatchB{E, s.Hp, s.Wp, s.Wl, C},
                          EpiBiasAct{O + (size_t)r0 * 36 * C, 36LL * C, nullptr, 0}, st)))
        return rc;
    }
    CIAOSR_LAUNCH(csa_fold_kernel, cdiv(4LL * HWp * C, 256), 256, 0, st, O, cv, s.Hp, s.Wp, C);
    EpiDown ed{out_nhwc ? out_nhwc + (size_t)b * H * W * ldo : nullptr, ldo,
               out_nchw ? out_nchw + (size_t)b * C * H * W : nullptr, H * W, plan + L.down_b};
    if ((rc = gemm_simt(H * W, C, 9 * C, DownA{cv, W, 2 * s.Hp, 2 * s.Wp, C},
                        RowMajorB{plan + L.down_wt, C}, ed, st))) return rc;
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr
