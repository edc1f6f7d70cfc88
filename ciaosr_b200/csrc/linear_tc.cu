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
  const float* x; int K;
  struct Row { const float* r; };
  __device__ __forceinline__ Row row(long long m) const { return Row{x + m * K}; }
  __device__ __forceinline__ void fill(Row& r, long long, int k0, float (&v)[32]) const {
    float4 q[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      q[g] = __ldg(reinterpret_cast<const float4*>(r.r + (k < K ? k : 0)));
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const bool ok = k0 + 4 * g < K;
      v[4 * g] = ok ? q[g].x : 0.f; v[4 * g + 1] = ok ? q[g].y : 0.f;
      v[4 * g + 2] = ok ? q[g].z : 0.f; v[4 * g + 3] = ok ? q[g].w : 0.f;
    }
  }
};
struct LinSrc {            // B[n, k] = W[n, k]
  const float* w; int K;
  __device__ __forceinline__ float operator()(int, int n, int k) const { return w[(long long)n * K + k]; }
};
struct LinEpi {            // out[m, n] = act(acc + bias[n]), N % 4 == 0
  float* o; const float* bias; int N; int act;
  __device__ __forceinline__ void store(const LinRowsGen::Row&, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(o + m * N + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + 4 * j;
      if (n >= N) break;
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float a = v[4 * j + i] + (bias ? __ldg(bias + n + i) : 0.0f);
        if (act == 1) a = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));      // nn.GELU() (exact)
        t[i] = a;
      }
      dst[j] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
};

static int lin_check(const ciaosr_linear_desc* d) {
  CIAOSR_REQUIRE(d != nullptr, CIAOSR_E_INVALID, "desc is NULL");
  CIAOSR_REQUIRE(d->abi_version == CIAOSR_ABI_VERSION, CIAOSR_E_INVALID, "ABI version mismatch");
  CIAOSR_REQUIRE(d->in_features > 0 && d->out_features > 0 && d->in_features % 4 == 0 && d->out_features % 4 == 0,
                 CIAOSR_E_INVALID, "linear: in_features and out_features must be positive multiples of 4, got %d -> %d",
                 d->in_features, d->out_features);
  CIAOSR_REQUIRE(d->weight != nullptr, CIAOSR_E_INVALID, "linear: weight is NULL");
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
  StageScope sc(6, (cudaStream_t)stream);
  const int kslabs = (d->in_features + KSLAB - 1) / KSLAB, nunits = (d->out_features + UNIT_N - 1) / UNIT_N;
  return tc_gemm(GemmShape{rows, kslabs, nunits, rows, 0}, reinterpret_cast<const uint8_t*>(plan),
                 LinRowsGen{x, d->in_features}, LinEpi{out, d->bias, d->out_features, activation},
                 (cudaStream_t)stream);
}

}  // extern "C"
