// fp32-grade Linear layer on the tensor cores:  out[rows, N] = act(x[rows, K] . W[N, K]^T + b)
//
// Encoder fast path for the SwinIR trunk (SURVEY.md section 8f "next" #2; reference: the nn.Linear layers of
// WindowAttention / Mlp, mmedited/models/backbones/sr_backbones/swinir_net.py:15-31, 66-146).  Parity with the
// fp32 reference rules out TF32 (features ~1e-3 off), and cuBLAS' fp32 CUDA-core sgemm runs these K = 180 / 360
// shapes at ~14 TFLOP/s; this is the functor GEMM of gemm_tc.cuh (fp16 hi/lo split, 3 UMMAs per product, fp32
// accumulation) reading the fp32 activations directly, with the bias and the exact (erf) GELU in the epilogue.
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ciaosr {

struct LinRowsGen {        // A = fp32 rows [M, K], K % 4 == 0; columns past K read as zero
  static constexpr bool kPrefetch = true;
  const float* x; int K;
  struct Row { const float* r; };
  struct Raw { float4 q[8]; };
  __device__ __forceinline__ Row row(long long m) const { return Row{x + m * K}; }
  __device__ __forceinline__ void issue(Row& r, long long, int k0, Raw& w) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      w.q[g] = __ldg(reinterpret_cast<const float4*>(r.r + (k < K ? k : 0)));
    }
  }
  __device__ __forceinline__ void finish(Row&, long long, int k0, const Raw& w, float (&v)[32]) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const bool ok = k0 + 4 * g < K;
      v[4 * g] = ok ? w.q[g].x : 0.f; v[4 * g + 1] = ok ? w.q[g].y : 0.f;
      v[4 * g + 2] = ok ? w.q[g].z : 0.f; v[4 * g + 3] = ok ? w.q[g].w : 0.f;
    }
  }
};
struct LinSrc {            // B[n, k] = W[n, k]
  const float* w; int K;
  __device__ __forceinline__ float operator()(int, int n, int k) const { return w[(long long)n * K + k]; }
};
// out[m, n] = act(acc + bias[n]) (+ res[m, n]), N % 4 == 0; bias 16-byte aligned.  o != nullptr: fp32 rows [M, N];
// else the fp16 hi / lo halves [M, ldo] (ldo % 8 == 0, pad columns zeroed) that the next TMA-fed Linear reads.
struct LinEpi {
  float* o; const float* bias; int N; int act; const float* res;
  split_t* o_hi = nullptr; split_t* o_lo = nullptr; int ldo = 0;
  template <class Row>
  __device__ __forceinline__ void store(const Row&, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(o + m * N + n0);
    const float4* rsrc = res ? reinterpret_cast<const float4*>(res + m * N + n0) : nullptr;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {                  // two halves of 16 columns: the 8 loads of a half are issued together
      float4 b4[4], r4[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * hh + jj;
        const bool in = n0 + 4 * j < N;
        b4[jj] = (bias && in) ? __ldg(reinterpret_cast<const float4*>(bias + n0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        r4[jj] = (rsrc && in) ? __ldg(rsrc + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * hh + jj;
        if (n0 + 4 * j >= N) break;
        const float bb[4] = {b4[jj].x, b4[jj].y, b4[jj].z, b4[jj].w};
        float t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = v[4 * j + i] + bb[i];
          if (act == 1) a = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));      // nn.GELU() (exact)
          else if (act == 2) a = fmaxf(a, 0.0f);                                         // nn.ReLU()
          t[i] = a;
        }
        t[0] += r4[jj].x; t[1] += r4[jj].y; t[2] += r4[jj].z; t[3] += r4[jj].w;
        if (o != nullptr) dst[j] = make_float4(t[0], t[1], t[2], t[3]);
        else {
          uint2 h, l;
          split2(t[0], t[1], h.x, l.x);
          split2(t[2], t[3], h.y, l.y);
          *reinterpret_cast<uint2*>(o_hi + m * ldo + n0 + 4 * j) = h;
          *reinterpret_cast<uint2*>(o_lo + m * ldo + n0 + 4 * j) = l;
        }
      }
    }
    if (o == nullptr && n0 <= N && N < n0 + 32 && N < ldo) {          // the chunk that holds column N: zero the pad
      for (int c = N; c < ldo; c += 4) {
        *reinterpret_cast<uint2*>(o_hi + m * ldo + c) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(o_lo + m * ldo + c) = make_uint2(0u, 0u);
      }
    }
  }
};

// The same epilogue as a TILE functor (gemm_tc.cuh, kTile): called with 4 consecutive columns of one row after the warp's
// accumulator chunk went through the smem transpose; used for SPLIT results, whose 8-byte-per-row stores are the worst
// case of the thread-per-row pattern (fc1 + GELU of the trunk: 173 -> 133 us).  fp32 results with a residual stay on
// the row functor above (proj: 71 us against 90 us in tile mode).
struct LinEpiTile {
  static constexpr bool kTile = true;
  const float* bias; int N; int act; const float* res;
  split_t* o_hi; split_t* o_lo; int ldo;
  __device__ __forceinline__ void store4(long long m, int n, float4 v) const {
    if (n >= N) {
      if (n < ldo) {                                            // the pad columns [N, ldo) of the split form are zero
        *reinterpret_cast<uint2*>(o_hi + m * ldo + n) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(o_lo + m * ldo + n) = make_uint2(0u, 0u);
      }
      return;
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (res) r = __ldg(reinterpret_cast<const float4*>(res + m * N + n));
    if (bias) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
      v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
    }
    if (act == 1) {                                             // nn.GELU() (exact)
      v.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752440f));
      v.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752440f));
      v.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752440f));
      v.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752440f));
    } else if (act == 2) {                                      // nn.ReLU()
      v.x = fmaxf(v.x, 0.0f); v.y = fmaxf(v.y, 0.0f); v.z = fmaxf(v.z, 0.0f); v.w = fmaxf(v.w, 0.0f);
    }
    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    uint2 h, l;
    split2(v.x, v.y, h.x, l.x);
    split2(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(o_hi + m * ldo + n) = h;
    *reinterpret_cast<uint2*>(o_lo + m * ldo + n) = l;
  }
};

// 3x3 convolution, stride 1, zero padding 1, on an NHWC map as an implicit GEMM: A[(b,y,x), k = t*Cin + ci] gathered
// from the map (t = ky*3 + kx; Cin % 4 == 0, so a float4 never straddles a tap), B[co, k] = weight[co, ci, ky, kx].
// The RSTB / trunk convolutions of SwinIR (swinir_net.py:446-483, 706-713) act on token tensors [B, HW, C], which ARE
// NHWC maps: no patch_unembed / patch_embed transposes, and the RSTB's residual rides in the epilogue.
struct Conv3Gen {
  static constexpr bool kPrefetch = true;
  const float* x; int H, W, C;
  struct Row { int y, x; };
  struct Raw { float4 q[8]; uint32_t ok; };
  __device__ __forceinline__ Row row(long long m) const {
    const int hw = (int)(m % ((long long)H * W));
    return Row{hw / W, hw % W};
  }
  __device__ __forceinline__ void issue(Row& r, long long m, int k0, Raw& w) const {
    const float* src[8];
    const float* self = x + m * C;
    w.ok = 0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      src[g] = self;
      if (k < 9 * C) {
        const int t = k / C, ch = k - t * C;
        const int dy = t / 3 - 1, dx = t - (dy + 1) * 3 - 1;
        if (r.y + dy >= 0 && r.y + dy < H && r.x + dx >= 0 && r.x + dx < W) {
          src[g] = self + (dy * W + dx) * C + ch;
          w.ok |= 1u << g;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) w.q[g] = __ldg(reinterpret_cast<const float4*>(src[g]));
  }
  __device__ __forceinline__ void finish(Row&, long long, int, const Raw& w, float (&v)[32]) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const bool ok = (w.ok >> g) & 1;
      v[4 * g] = ok ? w.q[g].x : 0.f; v[4 * g + 1] = ok ? w.q[g].y : 0.f;
      v[4 * g + 2] = ok ? w.q[g].z : 0.f; v[4 * g + 3] = ok ? w.q[g].w : 0.f;
    }
  }
};
struct Conv3Src {          // B[n = co, k = t*Cin + ci] = weight[co, ci, t]
  const float* w; int Cin;
  __device__ __forceinline__ float operator()(int, int n, int k) const {
    const int t = k / Cin, ci = k - t * Cin;
    return w[((long long)n * Cin + ci) * 9 + t];
  }
};

static int lin_check(const ciaosr_linear_desc* d) {
  CIAOSR_REQUIRE(d != nullptr, CIAOSR_E_INVALID, "desc is NULL");
  CIAOSR_REQUIRE(d->abi_version == CIAOSR_ABI_VERSION, CIAOSR_E_INVALID, "ABI version mismatch");
  CIAOSR_REQUIRE(d->in_features > 0 && d->out_features > 0 && d->in_features % 4 == 0 && d->out_features % 4 == 0,
                 CIAOSR_E_INVALID, "linear: in_features and out_features must be positive multiples of 4, got %d -> %d",
                 d->in_features, d->out_features);
  CIAOSR_REQUIRE(d->weight != nullptr, CIAOSR_E_INVALID, "linear: weight is NULL");
  CIAOSR_REQUIRE(((uintptr_t)d->bias % 16) == 0, CIAOSR_E_INVALID, "linear: bias must be 16-byte aligned");
  return CIAOSR_OK;
}

}  // namespace ciaosr

using namespace ciaosr;

extern "C" {

int ciaosr_linear_plan_bytes(const ciaosr_linear_desc* d, size_t* bytes) {
  CIAOSR_REQUIRE(bytes != nullptr, CIAOSR_E_INVALID, "bytes is NULL");
  int rc = lin_check(d);
  if (rc) return rc;
  *bytes = tc_operand_blob_bytes((d->in_features + KSLAB - 1) / KSLAB, (d->out_features + UNIT_N - 1) / UNIT_N);
  return CIAOSR_OK;
}

int ciaosr_linear_plan_init(const ciaosr_linear_desc* d, void* plan, size_t plan_bytes, void* stream) {
  int rc = lin_check(d);
  if (rc) return rc;
  size_t need = 0;
  ciaosr_linear_plan_bytes(d, &need);
  CIAOSR_REQUIRE(plan != nullptr && ((uintptr_t)plan % 256) == 0 && plan_bytes >= need, CIAOSR_E_WORKSPACE,
                 "linear plan buffer too small or misaligned: need %zu, have %zu", need, plan_bytes);
  return tc_pack_operand(reinterpret_cast<uint8_t*>(plan), 1, d->out_features, d->in_features, 0,
                         LinSrc{d->weight, d->in_features}, (cudaStream_t)stream);
}

int ciaosr_linear_forward(const ciaosr_linear_desc* d, const void* plan, const float* x, long long rows,
                          int activation, float* out, void* stream) {
  int rc = lin_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(plan && x && out, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(rows >= 0 && (activation == 0 || activation == 1), CIAOSR_E_INVALID,
                 "linear: bad rows=%lld or activation=%d", rows, activation);
  if (rows == 0) return CIAOSR_OK;
  return ciaosr_linear_forward_res(d, plan, x, rows, activation, nullptr, out, stream);
}

int ciaosr_linear_forward_res(const ciaosr_linear_desc* d, const void* plan, const float* x, long long rows,
                              int activation, const float* residual, float* out, void* stream) {
  int rc = lin_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(plan && x && out, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(rows >= 0 && activation >= 0 && activation <= 2, CIAOSR_E_INVALID,
                 "linear: bad rows=%lld or activation=%d", rows, activation);
  if (rows == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  const int kslabs = (d->in_features + KSLAB - 1) / KSLAB, nunits = (d->out_features + UNIT_N - 1) / UNIT_N;
  GemmShape g{rows, kslabs, nunits, rows, 0};
  g.a_resident = (kslabs <= 4 && nunits > 2) ? 1 : 0;        // short K, several N-chunks: generate A once per tile
  return tc_gemm(g, reinterpret_cast<const uint8_t*>(plan), LinRowsGen{x, d->in_features},
                 LinEpi{out, d->bias, d->out_features, activation, residual}, (cudaStream_t)stream);
}

int ciaosr_linear_forward_split(const ciaosr_linear_desc* d, const void* plan, const uint16_t* a_hi,
                                const uint16_t* a_lo, int lda, long long rows, int activation, const float* residual,
                                float* out, uint16_t* out_hi, uint16_t* out_lo, int ldo, void* stream) {
  int rc = lin_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(plan && a_hi && a_lo && (out || (out_hi && out_lo)), CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(rows >= 0 && activation >= 0 && activation <= 2, CIAOSR_E_INVALID,
                 "linear: bad rows=%lld or activation=%d", rows, activation);
  CIAOSR_REQUIRE(lda >= d->in_features && lda % 8 == 0, CIAOSR_E_INVALID,
                 "linear (split input): lda=%d must be a multiple of 8 >= in_features=%d", lda, d->in_features);
  CIAOSR_REQUIRE(out || (ldo >= d->out_features && ldo % 8 == 0), CIAOSR_E_INVALID,
                 "linear (split output): ldo=%d must be a multiple of 8 >= out_features=%d", ldo, d->out_features);
  if (rows == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  const int kslabs = (d->in_features + KSLAB - 1) / KSLAB, nunits = (d->out_features + UNIT_N - 1) / UNIT_N;
  // columns [in_features, lda) of the operand matrices are zero by contract, columns past lda read as zero (TMA)
  CUtensorMap map_hi, map_lo;
  if ((rc = tma_make_map_2d(&map_hi, const_cast<uint16_t*>(a_hi), rows, lda)) ||
      (rc = tma_make_map_2d(&map_lo, const_cast<uint16_t*>(a_lo), rows, lda))) return rc;
  if (out == nullptr)
    return tc_gemm(GemmShape{rows, kslabs, nunits, rows, 0}, reinterpret_cast<const uint8_t*>(plan), TmaRowsGen{},
                   LinEpiTile{d->bias, d->out_features, activation, residual, reinterpret_cast<split_t*>(out_hi),
                              reinterpret_cast<split_t*>(out_lo), ldo},
                   (cudaStream_t)stream, &map_hi, &map_lo);
  return tc_gemm(GemmShape{rows, kslabs, nunits, rows, 0}, reinterpret_cast<const uint8_t*>(plan), TmaRowsGen{},
                 LinEpi{out, d->bias, d->out_features, activation, residual}, (cudaStream_t)stream, &map_hi, &map_lo);
}

static int conv_check(const ciaosr_conv3x3_desc* d) {
  CIAOSR_REQUIRE(d != nullptr, CIAOSR_E_INVALID, "desc is NULL");
  CIAOSR_REQUIRE(d->abi_version == CIAOSR_ABI_VERSION, CIAOSR_E_INVALID, "ABI version mismatch");
  CIAOSR_REQUIRE(d->in_channels > 0 && d->out_channels > 0 && d->in_channels % 4 == 0 && d->out_channels % 4 == 0,
                 CIAOSR_E_INVALID, "conv3x3: channel counts must be positive multiples of 4, got %d -> %d",
                 d->in_channels, d->out_channels);
  CIAOSR_REQUIRE(d->weight != nullptr, CIAOSR_E_INVALID, "conv3x3: weight is NULL");
  CIAOSR_REQUIRE(((uintptr_t)d->bias % 16) == 0, CIAOSR_E_INVALID, "conv3x3: bias must be 16-byte aligned");
  return CIAOSR_OK;
}

int ciaosr_conv3x3_plan_bytes(const ciaosr_conv3x3_desc* d, size_t* bytes) {
  CIAOSR_REQUIRE(bytes != nullptr, CIAOSR_E_INVALID, "bytes is NULL");
  int rc = conv_check(d);
  if (rc) return rc;
  *bytes = tc_operand_blob_bytes((9 * d->in_channels + KSLAB - 1) / KSLAB, (d->out_channels + UNIT_N - 1) / UNIT_N);
  return CIAOSR_OK;
}

int ciaosr_conv3x3_plan_init(const ciaosr_conv3x3_desc* d, void* plan, size_t plan_bytes, void* stream) {
  int rc = conv_check(d);
  if (rc) return rc;
  size_t need = 0;
  ciaosr_conv3x3_plan_bytes(d, &need);
  CIAOSR_REQUIRE(plan != nullptr && ((uintptr_t)plan % 256) == 0 && plan_bytes >= need, CIAOSR_E_WORKSPACE,
                 "conv3x3 plan buffer too small or misaligned: need %zu, have %zu", need, plan_bytes);
  return tc_pack_operand(reinterpret_cast<uint8_t*>(plan), 1, d->out_channels, 9 * d->in_channels, 0,
                         Conv3Src{d->weight, d->in_channels}, (cudaStream_t)stream);
}

int ciaosr_conv3x3_nhwc_forward(const ciaosr_conv3x3_desc* d, const void* plan, const float* x, int B, int H, int W,
                                int activation, const float* residual, float* out, void* stream) {
  int rc = conv_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(plan && x && out, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(B >= 0 && H > 0 && W > 0 && (activation == 0 || activation == 2), CIAOSR_E_INVALID,
                 "conv3x3: bad shape B=%d H=%d W=%d or activation=%d (0 none, 2 ReLU)", B, H, W, activation);
  if (B == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  const long long rows = (long long)B * H * W;
  const int kslabs = (9 * d->in_channels + KSLAB - 1) / KSLAB, nunits = (d->out_channels + UNIT_N - 1) / UNIT_N;
  return tc_gemm(GemmShape{rows, kslabs, nunits, rows, 0}, reinterpret_cast<const uint8_t*>(plan),
                 Conv3Gen{x, H, W, d->in_channels}, LinEpi{out, d->bias, d->out_channels, activation, residual},
                 (cudaStream_t)stream);
}

}  // extern "C"

#ifdef CIAOSR_TC_TIMING
// diagnostic build only: cycles spent in mbarrier waits by the kernels of this translation unit (tools/wait_linear.py)
extern "C" int ciaosr_debug_wait_read_linear(unsigned long long* cycles, unsigned long long* counts, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(cycles, ciaosr::tc::g_wait_cycles, 64 * 8);
  cudaMemcpyFromSymbol(counts, ciaosr::tc::g_wait_count, 64 * 8);
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(ciaosr::tc::g_wait_cycles, z, 64 * 8);
    cudaMemcpyToSymbol(ciaosr::tc::g_wait_count, z, 64 * 8);
  }
  return 0;
}
#endif
