// Cross-scale non-local attention on the tensor cores (arch_csnln.py:430-532).
//
// Same attention form as cs_attn.cu (see its header), with the three large contractions on
// tcgen05 through the generic functor GEMM (gemm_tc.cuh, fp16 hi/lo split x3 = fp32-grade):
//   S = 10 * Q K^T        A = 3x3 patches of Mi (implicit), B = normalised 3x3 patches of R (packed per image)
//   O = P V               A = softmax rows,                 B = V^T: 6x6 stride-2 patches of E (packed per image)
//   out = down(canvas)/6  A = 3x3 stride-2 patches of the folded canvas, B = down-conv weights
// The 1x1 embeddings, the row softmax and the fold stay on CUDA cores (they are bandwidth-trivial),
// batched over all images of the call.
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ciaosr {

// ---- batched CUDA-core pieces --------------------------------------------------------------------
struct PadFeatBatchA {      // A[(img, y, x) in padded coords, ci], reflect pad bottom/right
  const float* f; int H, W, Hp, Wp, C;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int img = m / (Hp * Wp), r = m % (Hp * Wp);
    int y = r / Wp, x = r % Wp;
    if (y >= H) y = 2 * (H - 1) - y;
    if (x >= W) x = 2 * (W - 1) - x;
    return f[(((long long)img * H + y) * W + x) * C + k];
  }
};
struct PoolFeatBatchA {     // 2x2 average of the padded image
  const float* f; int H, W, Hl, Wl, C;
  __device__ __forceinline__ float at(int img, int y, int x, int k) const {
    if (y >= H) y = 2 * (H - 1) - y;
    if (x >= W) x = 2 * (W - 1) - x;
    return f[(((long long)img * H + y) * W + x) * C + k];
  }
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int img = m / (Hl * Wl), r = m % (Hl * Wl);
    const int y = 2 * (r / Wl), x = 2 * (r % Wl);
    const float r0 = at(img, y, x, k) * 0.5f + at(img, y + 1, x, k) * 0.5f;
    const float r1 = at(img, y, x + 1, k) * 0.5f + at(img, y + 1, x + 1, k) * 0.5f;
    return r0 * 0.5f + r1 * 0.5f;
  }
};
struct EpiPrelu {
  float* c; int ldc; const float* bias; const float* slope;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const float v = acc + bias[n];
    c[(long long)m * ldc + n] = v >= 0.0f ? v : v * slope[0];
  }
};

__global__ void csa_knorm_batch_kernel(const float* __restrict__ r, float* __restrict__ nrm, int Hl, int Wl,
                                       int Ch, const float* __restrict__ scalars) {
  const int gl = blockIdx.x, img = gl / (Hl * Wl), l = gl % (Hl * Wl);
  const float* ri = r + (long long)img * Hl * Wl * Ch;
  float ss = 0.0f;
  for (int i = threadIdx.x; i < 9 * Ch; i += blockDim.x) {
    const int t = i / Ch, c = i % Ch;
    const int y = l / Wl + t / 3 - 1, x = l % Wl + t % 3 - 1;
    if (y >= 0 && y < Hl && x >= 0 && x < Wl) {
      const float v = ri[((long long)y * Wl + x) * Ch + c];
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
    if (threadIdx.x == 0) nrm[gl] = fmaxf(sqrtf(ss), scalars[3]);
  }
}

// Softmax of the scores, written as the two 16-bit halves of P * 2^11 (row-major [rows, ldp], columns L..ldp
// zero) that the P.V GEMM loads with TMA.  The power-of-two scale keeps small probabilities out
// of fp16's subnormal range (absolute resolution 3e-11 instead of 6e-8); the P.V epilogue multiplies by 2^-11.
constexpr float CSA_P_SCALE = 2048.0f;
// One 256-thread block per row, the row cached in registers (NE values per thread): S is read once.
template <int NE>
__global__ void __launch_bounds__(256) softmax_rows_split_kernel(const float* __restrict__ s, split_t* __restrict__ p_hi,
                                                                 split_t* __restrict__ p_lo, int L, int ld, int ldp) {
  const long long row = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const float2* p = reinterpret_cast<const float2*>(s + row * ld);      // ld % 4 == 0
  __shared__ float red[8];
  float2 v[NE / 2];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NE / 2; ++j) {
    const int i = 2 * (t + 256 * j);
    v[j] = make_float2(-INFINITY, -INFINITY);
    if (i + 1 < L) v[j] = p[t + 256 * j];
    else if (i < L) v[j].x = s[row * ld + i];
    mx = fmaxf(mx, fmaxf(v[j].x, v[j].y));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.0f;
#pragma unroll
  for (int j = 0; j < NE / 2; ++j) {
    v[j].x = expf(v[j].x - mx);          // exp(-inf) = 0 for the padding
    v[j].y = expf(v[j].y - mx);
    sum += v[j].x + v[j].y;
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) sum += red[w];
  uint32_t* hi = reinterpret_cast<uint32_t*>(p_hi + row * ldp);
  uint32_t* lo = reinterpret_cast<uint32_t*>(p_lo + row * ldp);
#pragma unroll
  for (int j = 0; j < NE / 2; ++j) {
    const int i = 2 * (t + 256 * j);
    if (i < ldp) {                                            // ldp % 8 == 0; columns L..ldp are written as zero
      uint32_t h, l;
      split2(__fdiv_rn(v[j].x, sum) * CSA_P_SCALE, __fdiv_rn(v[j].y, sum) * CSA_P_SCALE, h, l);
      hi[i >> 1] = h;
      lo[i >> 1] = l;
    }
  }
}
static int softmax_rows_split(const float* s, split_t* p_hi, split_t* p_lo, long long rows, int L, int ld, int ldp,
                              cudaStream_t st) {
  CIAOSR_REQUIRE(L <= 256 * 128, CIAOSR_E_INVALID,
                 "cross-scale attention (tcgen05): %d keys per image exceed the softmax kernel's 32768; tile the input "
                 "(test_cfg.tile) or use the fp32 engine", L);
  if (L <= 256 * 8) CIAOSR_LAUNCH(softmax_rows_split_kernel<8>, (unsigned)rows, 256, 0, st, s, p_hi, p_lo, L, ld, ldp);
  else if (L <= 256 * 40) CIAOSR_LAUNCH(softmax_rows_split_kernel<40>, (unsigned)rows, 256, 0, st, s, p_hi, p_lo, L, ld, ldp);
  else CIAOSR_LAUNCH(softmax_rows_split_kernel<128>, (unsigned)rows, 256, 0, st, s, p_hi, p_lo, L, ld, ldp);
  return CIAOSR_OK;
}

__global__ void csa_fold_batch_kernel(const float* __restrict__ o, float* __restrict__ cv, int Hp, int Wp,
                                      int C, long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long p = i / C;
  const int W2 = 2 * Wp, H2 = 2 * Hp;
  const int X = (int)(p % W2), Y = (int)((p / W2) % H2);
  const long long img = p / ((long long)W2 * H2);
  const float* oi = o + img * Hp * Wp * 36LL * C;
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
      acc += oi[((long long)y * Wp + x) * (36LL * C) + (long long)ij * C + c];
    }
  }
  cv[i] = acc;
}

// ---- operand sources for the per-image blobs ----------------------------------------------------
struct KhatSrc {            // B[n = l, k = t*Chp + c] = R_pad[l + tap t, c] / max(|patch l|, eps); 0 for c >= Ch
  const float* r; const float* nrm; int Hl, Wl, Ch, Chp, img0;
  __device__ __forceinline__ float operator()(int image, int n, int k) const {
    const int img = img0 + image, t = k / Chp, c = k % Chp;
    const int y = n / Wl + t / 3 - 1, x = n % Wl + t % 3 - 1;
    if (c >= Ch || y < 0 || y >= Hl || x < 0 || x >= Wl) return 0.0f;
    return __fdiv_rn(r[(((long long)img * Hl + y) * Wl + x) * Ch + c], nrm[(long long)img * Hl * Wl + n]);
  }
};
struct VtSrc {              // B[n = (i*6+j)*C + c, k = l] = E_pad[c, 2ly-2+i, 2lx-2+j]
  const float* e; int Hp, Wp, Wl, C, img0;
  __device__ __forceinline__ float operator()(int image, int n, int k) const {
    const int img = img0 + image, ij = n / C, c = n % C;
    const int y = 2 * (k / Wl) - 2 + ij / 6, x = 2 * (k % Wl) - 2 + ij % 6;
    if (y < 0 || y >= Hp || x < 0 || x >= Wp) return 0.0f;
    return e[(((long long)img * Hp + y) * Wp + x) * C + c];
  }
};
struct DownSrc {            // B[n = co, k = (u*3+v)*C + ci] = down_wt[k, co]
  const float* w; int C;
  __device__ __forceinline__ float operator()(int, int n, int k) const { return w[(long long)k * C + n]; }
};

// ---- A generators / epilogues ----------------------------------------------------------------------
// Q operand of the score GEMM: 3x3 zero-padded patches of Mi (k = t*Chp + c), materialised ONCE per pass as the
// two 16-bit halves [rows, ldq] that the GEMM then loads with TMA for each of its L/256 column chunks (generating
// the patches in the GEMM's row threads re-gathered and re-split them per chunk: 36x on a 192x192 tile).
__global__ void csa_qpatch_split_kernel(const float* __restrict__ mi, split_t* __restrict__ q_hi,
                                        split_t* __restrict__ q_lo, int Hp, int Wp, int Chp, int ldq, long long pix0,
                                        long long total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;      // one thread per 4 columns
  if (i >= total) return;
  const int per_row = ldq / 4;
  const long long m = i / per_row;
  const int k = (int)(i % per_row) * 4;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < 9 * Chp) {
    const long long pix = pix0 + m;
    const int r = (int)(pix % ((long long)Hp * Wp)), y = r / Wp, x = r % Wp;
    const int t = k / Chp, c = k - t * Chp;
    const int dy = t / 3 - 1, dx = t % 3 - 1;
    if (y + dy >= 0 && y + dy < Hp && x + dx >= 0 && x + dx < Wp)
      q = __ldg(reinterpret_cast<const float4*>(mi + (pix + dy * Wp + dx) * Chp + c));
  }
  uint2 h, l;
  split2(q.x, q.y, h.x, l.x);
  split2(q.z, q.w, h.y, l.y);
  *reinterpret_cast<uint2*>(q_hi + m * ldq + k) = h;
  *reinterpret_cast<uint2*>(q_lo + m * ldq + k) = l;
}
struct ScoreEpi {           // S[m, n] = scale * acc, n < L
  float* s; int L, ld; float scale;
  __device__ __forceinline__ void store(const TmaRowsGen::Row&, long long m, int n0, const float (&v)[32]) const {
    float* dst = s + m * ld + n0;
    if (n0 + 32 <= L) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst)[j] =
            make_float4(v[4 * j] * scale, v[4 * j + 1] * scale, v[4 * j + 2] * scale, v[4 * j + 3] * scale);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (n0 + i < L) dst[i] = v[i] * scale;
    }
  }
};
struct OutEpi {             // O[m, n] = sc * acc, n < N (N % 4 == 0); accumulate: O += sc * acc (K chunks, see gemm_tc.cuh)
  float* o; int N; float sc;
  __device__ __forceinline__ void store(const TmaRowsGen::Row&, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(o + m * N + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + 4 * j < N) dst[j] = make_float4(sc * v[4 * j], sc * v[4 * j + 1], sc * v[4 * j + 2], sc * v[4 * j + 3]);
  }
  __device__ __forceinline__ void accumulate(const TmaRowsGen::Row&, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(o + m * N + n0);
    float4 old[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + 4 * j < N) old[j] = dst[j];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + 4 * j < N)
        dst[j] = make_float4(fmaf(sc, v[4 * j], old[j].x), fmaf(sc, v[4 * j + 1], old[j].y),
                             fmaf(sc, v[4 * j + 2], old[j].z), fmaf(sc, v[4 * j + 3], old[j].w));
  }
};
struct DownGen {            // rows = cropped output pixels (img, y, x); k = (u*3+v)*C + ci
  const float* cv; int H, W, H2, W2, C, K; long long img0;
  struct Row { int y, x, hw; long long img; };
  __device__ __forceinline__ Row row(long long m) const {
    const int hw = (int)(m % ((long long)H * W));
    return Row{hw / W, hw % W, hw, m / ((long long)H * W)};
  }
  __device__ __forceinline__ void fill(Row& r, long long, int k0, float (&v)[32]) const {
    const float* centre = cv + (((r.img * H2) + 2 * r.y) * W2 + 2 * r.x) * C;    // always inside the canvas
    const float* src[8];
    bool ok[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      src[g] = centre; ok[g] = false;
      if (k < K) {
        const int uv = k / C, ci = k - uv * C;
        const int u = uv / 3, w = uv - u * 3;
        const int Y = 2 * r.y - 1 + u, X = 2 * r.x - 1 + w;
        ok[g] = Y >= 0 && Y < H2 && X >= 0 && X < W2;
        if (ok[g]) src[g] = centre + ((u - 1) * W2 + (w - 1)) * C + ci;
      }
    }
    float4 q[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) q[g] = __ldg(reinterpret_cast<const float4*>(src[g]));
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      v[4 * g] = ok[g] ? q[g].x : 0.f; v[4 * g + 1] = ok[g] ? q[g].y : 0.f;
      v[4 * g + 2] = ok[g] ? q[g].z : 0.f; v[4 * g + 3] = ok[g] ? q[g].w : 0.f;
    }
  }
};
struct DownEpi {            // (acc + b) / 6 -> NHWC slice and / or NCHW
  float* o_nhwc; int ldo; float* o_nchw; int HW, C; const float* bias; long long img0;
  __device__ __forceinline__ void store(const DownGen::Row& r, long long m, int n0, const float (&v)[32]) const {
    const long long gm = img0 * HW + m;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int n = n0 + i;
      if (n < C) {
        const float val = __fdiv_rn(v[i] + bias[n], 6.0f);
        if (o_nhwc) o_nhwc[gm * ldo + n] = val;
        if (o_nchw) o_nchw[((img0 + r.img) * C + n) * HW + r.hw] = val;
      }
    }
  }
};

// ---- host orchestration ------------------------------------------------------------------------------
struct CsaTcSizes {
  int Hp, Wp, Hl, Wl, L, ldS, HWp, group;      // group = images processed together
  int Chp, ldP, ldQ;                                // ldP: row stride of the split P matrices (16-bit elements, 16-byte rows)
  int kq_slabs, kq_units, vt_slabs, vt_units, dn_slabs, dn_units;
};
static CsaTcSizes csa_tc_sizes(int B, int H, int W, int C) {
  CsaTcSizes s;
  s.Hp = H + (H & 1); s.Wp = W + (W & 1);
  s.Hl = s.Hp / 2; s.Wl = s.Wp / 2; s.L = s.Hl * s.Wl; s.ldS = (s.L + 3) / 4 * 4;
  s.HWp = s.Hp * s.Wp; s.ldP = (s.L + 7) / 8 * 8;
  s.Chp = (C / 2 + 3) / 4 * 4;                 // query-embedding channels, zero padded to the float4 gathers
  s.kq_slabs = (9 * s.Chp + KSLAB - 1) / KSLAB; s.ldQ = (9 * s.Chp + 7) / 8 * 8; s.kq_units = (s.L + UNIT_N - 1) / UNIT_N;
  s.vt_slabs = (s.L + KSLAB - 1) / KSLAB; s.vt_units = (36 * C + UNIT_N - 1) / UNIT_N;
  s.dn_slabs = (9 * C + KSLAB - 1) / KSLAB; s.dn_units = (C + UNIT_N - 1) / UNIT_N;
  // images per pass: tiles must not straddle images, and the score tensor stays <= 1 GiB
  long long g = (256LL << 20) / ((long long)s.HWp * s.ldS);
  if (g < 1) g = 1;
  if (g > B) g = B;
  if (s.HWp % ROWS != 0) g = 1;
  s.group = (int)g;
  return s;
}

// P.V accumulates 3072 keys per TMEM pass; the passes are summed in fp32 (gemm_tc.cuh).  Measured on a 192x192
// tile (L = 9216 keys, C = 64; max-abs vs the fp32 engine / time): one pass 5.7e-5 / 8.4 ms, 48 slabs 4.0e-5 /
// 8.9 ms, 16 slabs 3.1e-5 / 10.6 ms.  CIAOSR_CSA_KCHUNK overrides the slabs per pass (0 = one pass).
static int csa_kchunk() {
  static int kc = -1;
  if (kc < 0) {
    const char* e = getenv("CIAOSR_CSA_KCHUNK");
    kc = e ? atoi(e) / 4 * 4 : 48;
  }
  return kc;
}

bool cs_attn_tc_ok(const PlanLayout& L) { return L.non_local && L.C % 4 == 0; }

struct CsaTcBufs { float *E, *Mi, *R, *nrm, *S, *O, *cv; split_t *Ph, *Pl, *Qh, *Ql; uint8_t *kblob, *vblob, *dblob; };
static CsaTcBufs csa_tc_carve(Arena& a, const PlanLayout& L, const CsaTcSizes& s, int B) {
  CsaTcBufs b;
  const int C = L.C, Ch = C / 2, g = s.group;
  b.E = a.take<float>((size_t)B * s.HWp * C);
  b.Mi = a.take<float>((size_t)B * s.HWp * s.Chp);
  b.R = a.take<float>((size_t)B * s.L * Ch);
  b.nrm = a.take<float>((size_t)B * s.L);
  b.S = a.take<float>((size_t)g * s.HWp * s.ldS);
  b.Qh = a.take<split_t>((size_t)g * s.HWp * s.ldQ);
  b.Ql = a.take<split_t>((size_t)g * s.HWp * s.ldQ);
  b.Ph = a.take<split_t>((size_t)g * s.HWp * s.ldP);
  b.Pl = a.take<split_t>((size_t)g * s.HWp * s.ldP);
  b.O = a.take<float>((size_t)g * s.HWp * 36 * C);
  b.cv = a.take<float>((size_t)g * 4 * s.HWp * C);
  b.kblob = a.take<uint8_t>((size_t)g * tc_operand_blob_bytes(s.kq_slabs, s.kq_units));
  b.vblob = a.take<uint8_t>((size_t)g * tc_operand_blob_bytes(s.vt_slabs, s.vt_units));
  b.dblob = a.take<uint8_t>(tc_operand_blob_bytes(s.dn_slabs, s.dn_units));
  return b;
}

size_t cs_attn_tc_workspace(const PlanLayout& L, int B, int H, int W) {
  Arena a(nullptr, 0);
  csa_tc_carve(a, L, csa_tc_sizes(B, H, W, L.C), B);
  return a.used();
}

int run_cs_attn_tc(const PlanLayout& L, const float* plan, const float* featT, int B, int H, int W,
                   float* out_nhwc, int ldo, float* out_nchw, void* ws, size_t ws_bytes, cudaStream_t st) {
  CIAOSR_REQUIRE(H >= 2 && W >= 2, CIAOSR_E_INVALID,
                 "cross-scale attention needs H, W >= 2 (reflect padding), got %dx%d", H, W);
  const CsaTcSizes s = csa_tc_sizes(B, H, W, L.C);
  const int C = L.C, Ch = C / 2;
  Arena a(ws, ws_bytes);
  CsaTcBufs b = csa_tc_carve(a, L, s, B);
  CIAOSR_REQUIRE(a.ok, CIAOSR_E_WORKSPACE, "cross-scale attention workspace too small: need %zu, have %zu",
                 a.used(), ws_bytes);
  const float* scal = plan + L.scalars;
  int rc;
  // embeddings for every image of the call
  PadFeatBatchA pa{featT, H, W, s.Hp, s.Wp, C};
  if ((rc = gemm_simt(B * s.HWp, C, C, pa, RowMajorB{plan + L.as_wt, C},
                      EpiPrelu{b.E, C, plan + L.as_b, scal + 2}, st))) return rc;
  if (s.Chp != Ch) CIAOSR_CUDA_OK(cudaMemsetAsync(b.Mi, 0, (size_t)B * s.HWp * s.Chp * sizeof(float), st));
  if ((rc = gemm_simt(B * s.HWp, Ch, C, pa, RowMajorB{plan + L.m1_wt, Ch},
                      EpiPrelu{b.Mi, s.Chp, plan + L.m1_b, scal + 0}, st))) return rc;
  if ((rc = gemm_simt(B * s.L, Ch, C, PoolFeatBatchA{featT, H, W, s.Hl, s.Wl, C}, RowMajorB{plan + L.m2_wt, Ch},
                      EpiPrelu{b.R, Ch, plan + L.m2_b, scal + 1}, st))) return rc;
  CIAOSR_LAUNCH(csa_knorm_batch_kernel, B * s.L, 128, 0, st, b.R, b.nrm, s.Hl, s.Wl, Ch, scal);
  if ((rc = tc_pack_operand(b.dblob, 1, C, 9 * C, 0, DownSrc{plan + L.down_wt, C}, st))) return rc;

  const size_t kstride = tc_operand_blob_bytes(s.kq_slabs, s.kq_units);
  const size_t vstride = tc_operand_blob_bytes(s.vt_slabs, s.vt_units);
  for (int i0 = 0; i0 < B; i0 += s.group) {
    const int g = min(s.group, B - i0);
    const long long rows = (long long)g * s.HWp;
    if ((rc = tc_pack_operand(b.kblob, g, s.L, 9 * s.Chp, kstride, KhatSrc{b.R, b.nrm, s.Hl, s.Wl, Ch, s.Chp, i0}, st)))
      return rc;
    {
      const long long total = rows * (s.ldQ / 4);
      CIAOSR_LAUNCH(csa_qpatch_split_kernel, cdiv(total, 256), 256, 0, st, b.Mi, b.Qh, b.Ql, s.Hp, s.Wp, s.Chp, s.ldQ,
                    (long long)i0 * s.HWp, total);
      CUtensorMap qmap_hi, qmap_lo;
      if ((rc = tma_make_map_2d(&qmap_hi, b.Qh, rows, s.ldQ)) || (rc = tma_make_map_2d(&qmap_lo, b.Ql, rows, s.ldQ)))
        return rc;
      if ((rc = tc_gemm(GemmShape{rows, s.kq_slabs, s.kq_units, s.HWp, kstride}, b.kblob, TmaRowsGen{},
                        ScoreEpi{b.S, s.L, s.ldS, L.cs_softmax_scale}, st, &qmap_hi, &qmap_lo))) return rc;
    }
    if ((rc = softmax_rows_split(b.S, b.Ph, b.Pl, rows, s.L, s.ldS, s.ldP, st))) return rc;
    CUtensorMap map_hi, map_lo;
    if ((rc = tma_make_map_2d(&map_hi, b.Ph, rows, s.ldP)) || (rc = tma_make_map_2d(&map_lo, b.Pl, rows, s.ldP))) return rc;
    if ((rc = tc_pack_operand(b.vblob, g, 36 * C, s.L, vstride, VtSrc{b.E, s.Hp, s.Wp, s.Wl, C, i0}, st)))
      return rc;
    if ((rc = tc_gemm(GemmShape{rows, s.vt_slabs, s.vt_units, s.HWp, vstride, csa_kchunk()}, b.vblob,
                      TmaRowsGen{}, OutEpi{b.O, 36 * C, 1.0f / CSA_P_SCALE}, st, &map_hi, &map_lo))) return rc;
    const long long ctot = (long long)g * 4 * s.HWp * C;
    CIAOSR_LAUNCH(csa_fold_batch_kernel, cdiv(ctot, 256), 256, 0, st, b.O, b.cv, s.Hp, s.Wp, C, ctot);
    if ((rc = tc_gemm(GemmShape{(long long)g * H * W, s.dn_slabs, s.dn_units, (long long)g * H * W, 0}, b.dblob,
                      DownGen{b.cv, H, W, 2 * s.Hp, 2 * s.Wp, C, 9 * C, 0},
                      DownEpi{out_nhwc, ldo, out_nchw, H * W, C, plan + L.down_b, i0}, st))) return rc;
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr
