// Window attention of the SwinIR trunk (SURVEY.md section 8f "next" #2: encoder fast path).
//
// Reference: WindowAttention.forward (mmedited/models/backbones/sr_backbones/swinir_net.py:112-146) as called from
// SwinTransformerBlock.forward (:240-280): cyclic shift (torch.roll), window_partition, per-window multi-head
// attention  softmax(q * scale @ k^T + relative_position_bias [+ SW-MSA mask]) @ v,  window_reverse, reverse shift.
// In eager PyTorch that is two rolls, two partition copies, the [3, B_, heads, N, d] permute of qkv, an expanded
// [B_, heads, N, N] bias + mask tensor, the attention itself and the head-concatenating transpose: ~10 passes over
// the token tensor per block.  Here it is ONE kernel that reads the qkv Linear's output in NATURAL token order and
// writes the attention output in natural order; shift, partition, bias, mask and their inverses are index
// arithmetic.  fp32 CUDA cores throughout (the contraction is 64 x 64 x 30 per head: 3 % of the trunk's FLOPs), online
// softmax, no tensor-core rounding to account for.
//
// One CTA per window, one thread per (head, query token).  K and V of the window sit in shared memory with the head
// dimension padded to 32 floats, read as broadcast float4 (all lanes of a warp share the head).
#include "common.cuh"
#include "tc_common.cuh"

namespace ciaosr {
using tc::split_t;
using tc::split2;

constexpr int WA_DPAD = 32;           // head dim padded to this in smem (d <= 32)

struct WinAttnParams {
  const float* qkv;                   // [B, H*W, 3C]: q | k | v, each [heads, d]
  const float* bias_table;            // [(2ws-1)^2, heads]
  float* out;                         // [B, H*W, C] fp32, or nullptr when the split outputs below are used
  int H, W, C, heads, d, ws, shift;
  float scale;
  split_t* out_hi; split_t* out_lo;   // [B*H*W, ldo] fp16 hi / lo halves (what a TMA-fed Linear reads), pad columns zeroed
  int ldo;
};

// WS = window size when known at compile time (8: SwinIR; 4: the small test trunks), 0 = runtime `P.ws`.
template <int WS>
__global__ void __launch_bounds__(384) window_attention_kernel(const WinAttnParams P) {
  extern __shared__ __align__(16) float smem[];
  const int ws = WS ? WS : P.ws, n = ws * ws, heads = P.heads, d = P.d, C = P.C, H = P.H, W = P.W;
  float* ks = smem;                                   // [n][heads][WA_DPAD]
  float* vs = smem + (size_t)n * heads * WA_DPAD;
  int* tok = reinterpret_cast<int*>(vs + (size_t)n * heads * WA_DPAD);   // [n] token index of window position j
  int* lab = tok + n;                                                    // [n] SW-MSA region label of position j
  const int wins_x = W / ws, wins_y = H / ws;
  const int win = blockIdx.x % (wins_x * wins_y), b = blockIdx.x / (wins_x * wins_y);
  const int wy = win / wins_x, wx = win % wins_x;
  const float* qkv_b = P.qkv + (long long)b * H * W * 3 * C;
  // position j of this window = pixel (wy*ws + jy, wx*ws + jx) of the map rolled by -shift = original pixel
  // ((. + shift) mod H, (. + shift) mod W); SW-MSA regions of the rolled map (swinir_net.py:217-238): rows
  // [0, H-ws), [H-ws, H-shift), [H-shift, H), same for columns
  if ((int)threadIdx.x < n) {
    const int j = threadIdx.x, ry = wy * ws + j / ws, rx = wx * ws + j % ws;
    int oy = ry + P.shift, ox = rx + P.shift;
    if (oy >= H) oy -= H;
    if (ox >= W) ox -= W;
    tok[j] = oy * W + ox;
    const int ly = ry < H - ws ? 0 : (ry < H - P.shift ? 1 : 2), lx = rx < W - ws ? 0 : (rx < W - P.shift ? 1 : 2);
    lab[j] = P.shift > 0 ? ly * 3 + lx : 0;
  }
  // smem offset (within one key's [heads][WA_DPAD] block; V blocks follow all K blocks) of column c of a row's
  // [k | v] half (2C floats), so that the staging loop below needs no divisions
  int* dst = lab + n;                                                    // [2C]
  const int kv_stride = n * heads * WA_DPAD;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const int isv = c >= C, cc = c - isv * C, hh = cc / d;
    dst[c] = isv * kv_stride + hh * WA_DPAD + (cc - hh * d);
  }
  for (int i = threadIdx.x; i < n * heads * (WA_DPAD - d); i += blockDim.x) {      // zero the head-dim padding once
    const int e = d + i % (WA_DPAD - d), r = i / (WA_DPAD - d);
    ks[r * WA_DPAD + e] = 0.0f;
    vs[r * WA_DPAD + e] = 0.0f;
  }
  __syncthreads();
  // stage K and V: a key's [k | v] is 2C contiguous floats (C % 4 == 0: 16-byte aligned float4s).  `rows_pp` keys per
  // pass, one float4 per thread, every load of a thread independent of the others (all in flight together).
  {
    const int f4_per_row = C / 2;                                        // float4s in 2C floats
    const int rows_pp = blockDim.x / f4_per_row;                         // >= 1 (host checks C/2 <= blockDim)
    const int r_in = threadIdx.x / f4_per_row, f4 = threadIdx.x - r_in * f4_per_row;
    if (r_in < rows_pp) {
      constexpr int MAXP = 16;                                           // passes held in registers at once
      for (int j0 = 0; j0 < n; j0 += rows_pp * MAXP) {
        float4 buf[MAXP];
#pragma unroll
        for (int p = 0; p < MAXP; ++p) {
          const int j = j0 + p * rows_pp + r_in;
          if (j < n) buf[p] = __ldg(reinterpret_cast<const float4*>(qkv_b + (long long)tok[j] * 3 * C + C) + f4);
        }
#pragma unroll
        for (int p = 0; p < MAXP; ++p) {
          const int j = j0 + p * rows_pp + r_in;
          if (j < n) {
            float* base = ks + (size_t)j * heads * WA_DPAD;
            base[dst[4 * f4]] = buf[p].x; base[dst[4 * f4 + 1]] = buf[p].y;
            base[dst[4 * f4 + 2]] = buf[p].z; base[dst[4 * f4 + 3]] = buf[p].w;
          }
        }
      }
    }
  }
  __syncthreads();
  const int h = threadIdx.x / n, i = threadIdx.x - h * n;
  if (h >= heads) return;
  const int ti = tok[i];
  float q[WA_DPAD], acc[WA_DPAD];
  {
    const float* qrow = qkv_b + (long long)ti * 3 * C + h * d;
#pragma unroll
    for (int e = 0; e < WA_DPAD; ++e) {
      q[e] = e < d ? __ldg(qrow + e) * P.scale : 0.0f;
      acc[e] = 0.0f;
    }
  }
  const int iy = i / ws, ix = i - iy * ws, lab_i = lab[i];
  // relative position bias index of (i, j): (iy - jy + ws - 1) * (2 ws - 1) + (ix - jx + ws - 1)
  const float* bias_i = P.bias_table + ((iy + ws - 1) * (2 * ws - 1) + ix + ws - 1) * heads + h;
  const bool masked = P.shift > 0;
  float m = -INFINITY, l = 0.0f;
  // one window row (ws keys) per step: ws independent dot products, ONE running-max update (online softmax)
  constexpr int JB = WS ? WS : 8;
  for (int j0 = 0; j0 < n; j0 += JB) {
    float sc[JB];
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      const int j = j0 + jj;
      float s0 = 0.f, s1 = 0.f;
      if (WS || j < n) {
        const float4* kr = reinterpret_cast<const float4*>(ks + ((size_t)j * heads + h) * WA_DPAD);
#pragma unroll
        for (int e4 = 0; e4 < WA_DPAD / 4; ++e4) {
          const float4 k4 = kr[e4];
          s0 = fmaf(q[4 * e4], k4.x, s0); s1 = fmaf(q[4 * e4 + 1], k4.y, s1);
          s0 = fmaf(q[4 * e4 + 2], k4.z, s0); s1 = fmaf(q[4 * e4 + 3], k4.w, s1);
        }
        const int jy = j / ws, jx = j - jy * ws;
        s0 += s1 + __ldg(bias_i - (jy * (2 * ws - 1) + jx) * heads);
        if (masked && lab[j] != lab_i) s0 += -100.0f;
      } else {
        s0 = -INFINITY;
      }
      sc[jj] = s0;
    }
    float mn = m;
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) mn = fmaxf(mn, sc[jj]);
    const float f = __expf(m - mn);                    // exp(-inf) = 0 on the first step
    m = mn;
    l *= f;
#pragma unroll
    for (int e = 0; e < WA_DPAD; ++e) acc[e] *= f;
#pragma unroll
    for (int jj = 0; jj < JB; ++jj) {
      const int j = j0 + jj;
      if (!WS && j >= n) break;
      const float p = expf(sc[jj] - m);
      l += p;
      const float4* vr = reinterpret_cast<const float4*>(vs + ((size_t)j * heads + h) * WA_DPAD);
#pragma unroll
      for (int e4 = 0; e4 < WA_DPAD / 4; ++e4) {
        const float4 v4 = vr[e4];
        acc[4 * e4] = fmaf(p, v4.x, acc[4 * e4]); acc[4 * e4 + 1] = fmaf(p, v4.y, acc[4 * e4 + 1]);
        acc[4 * e4 + 2] = fmaf(p, v4.z, acc[4 * e4 + 2]); acc[4 * e4 + 3] = fmaf(p, v4.w, acc[4 * e4 + 3]);
      }
    }
  }
  const long long orow_i = (long long)b * H * W + ti;
  if (P.out != nullptr) {
    float* orow = P.out + orow_i * C + h * d;
#pragma unroll
    for (int e = 0; e < WA_DPAD; ++e)
      if (e < d) orow[e] = __fdiv_rn(acc[e], l);
  } else {                                             // d is even: pairs of columns as one 32-bit store per half
    uint32_t* hrow = reinterpret_cast<uint32_t*>(P.out_hi + orow_i * P.ldo + h * d);
    uint32_t* lrow = reinterpret_cast<uint32_t*>(P.out_lo + orow_i * P.ldo + h * d);
#pragma unroll
    for (int e = 0; e < WA_DPAD; e += 2) {
      if (e < d) {
        uint32_t hi, lo;
        split2(__fdiv_rn(acc[e], l), __fdiv_rn(acc[e + 1], l), hi, lo);
        hrow[e >> 1] = hi;
        lrow[e >> 1] = lo;
      }
    }
    if (h == heads - 1)
      for (int c = C; c < P.ldo; ++c) {
        P.out_hi[orow_i * P.ldo + c] = split_t(0.0f);
        P.out_lo[orow_i * P.ldo + c] = split_t(0.0f);
      }
  }
}

// LayerNorm over the last dimension (nn.LayerNorm(C), swinir_net.py:195,207,702): one warp per row, the row held
// in registers (C <= 32 * LN_MAXV), two-pass mean / variance like ATen, y = (x - mean) / sqrt(var + eps) * g + b.
constexpr int LN_MAXV = 16;
// y != nullptr: fp32 output; else the fp16 hi / lo halves [rows, ld] a TMA-fed Linear reads (pad columns zeroed).
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                             const float* __restrict__ b, float eps, long long rows,
                                                             int C, float* __restrict__ y, split_t* __restrict__ y_hi,
                                                             split_t* __restrict__ y_lo, int ld) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * C;
  float v[LN_MAXV];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < C ? __ldg(xr + c) : 0.0f;
    sum += v[i];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const float d = (lane + 32 * i) < C ? v[i] - mean : 0.0f;
    sq = fmaf(d, d, sq);
  }
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.0f / sqrtf(sq / (float)C + eps);
  if (y != nullptr) {
    float* yr = y + row * C;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < C) yr[c] = (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c);
    }
  } else {
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int c = lane + 32 * i;
      if (c < ld) {
        const float o = c < C ? (v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c) : 0.0f;
        split_t hi, lo;
        tc::split_scalar(o, hi, lo);
        y_hi[row * ld + c] = hi;
        y_lo[row * ld + c] = lo;
      }
    }
  }
}

}  // namespace ciaosr

using namespace ciaosr;

extern "C" int ciaosr_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps,
                                        long long rows, int C, float* out, void* stream) {
  CIAOSR_REQUIRE(x && gamma && beta && out, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(rows >= 0 && C > 0 && C <= 32 * LN_MAXV, CIAOSR_E_INVALID,
                 "layernorm: rows=%lld, C=%d outside the supported range (C <= %d)", rows, C, 32 * LN_MAXV);
  if (rows == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  CIAOSR_LAUNCH(layernorm_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, x, gamma, beta, eps,
                rows, C, out, (split_t*)nullptr, (split_t*)nullptr, 0);
  return CIAOSR_OK;
}

extern "C" int ciaosr_layernorm_split_forward(const float* x, const float* gamma, const float* beta, float eps,
                                              long long rows, int C, uint16_t* out_hi, uint16_t* out_lo, int ld,
                                              void* stream) {
  CIAOSR_REQUIRE(x && gamma && beta && out_hi && out_lo, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(rows >= 0 && C > 0 && ld >= C && ld % 8 == 0 && ld <= 32 * LN_MAXV, CIAOSR_E_INVALID,
                 "layernorm (split): rows=%lld, C=%d, ld=%d outside the supported range (ld %% 8 == 0, C <= ld <= %d)",
                 rows, C, ld, 32 * LN_MAXV);
  if (rows == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  CIAOSR_LAUNCH(layernorm_rows_kernel, (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, x, gamma, beta, eps,
                rows, C, (float*)nullptr, reinterpret_cast<split_t*>(out_hi), reinterpret_cast<split_t*>(out_lo), ld);
  return CIAOSR_OK;
}

static int window_attention_launch(const float* qkv, const float* bias_table, int B, int H, int W, int C, int heads,
                                   int ws, int shift, float scale, float* out, split_t* out_hi, split_t* out_lo,
                                   int ldo, void* stream) {
  CIAOSR_REQUIRE(qkv && bias_table && (out || (out_hi && out_lo)), CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(out || (ldo >= C && ldo % 8 == 0 && (C / (heads > 0 ? heads : 1)) % 2 == 0), CIAOSR_E_INVALID,
                 "window attention (split output): ld=%d must be a multiple of 8 >= C=%d and the head dim even", ldo, C);
  CIAOSR_REQUIRE(B >= 0 && H > 0 && W > 0 && ws > 0 && H % ws == 0 && W % ws == 0, CIAOSR_E_INVALID,
                 "window attention: the %dx%d map must be a multiple of the %d-pixel window", H, W, ws);
  CIAOSR_REQUIRE(heads > 0 && C % heads == 0 && C % 4 == 0 && C / heads <= WA_DPAD && ws * ws * heads <= 384 &&
                     C / 2 <= (ws * ws * heads + 31) / 32 * 32,
                 CIAOSR_E_INVALID,
                 "window attention: unsupported geometry C=%d heads=%d window=%d (need C %% 4 == 0, head dim <= %d, "
                 "window^2 * heads <= 384 and C / 2 <= that thread count)", C, heads, ws, WA_DPAD);
  CIAOSR_REQUIRE(shift >= 0 && shift < ws, CIAOSR_E_INVALID, "window attention: shift %d outside [0, %d)", shift, ws);
  if (B == 0) return CIAOSR_OK;
  StageScope sc(6, (cudaStream_t)stream);
  WinAttnParams P{qkv, bias_table, out, H, W, C, heads, C / heads, ws, shift, scale, out_hi, out_lo, ldo};
  const int n = ws * ws;
  const int smem = 2 * n * heads * WA_DPAD * (int)sizeof(float) + (2 * n + 2 * C) * (int)sizeof(int);
  static DynSmemOptIn optin[3];
  const int threads = (n * heads + 31) / 32 * 32;
  const long long grid = (long long)B * (H / ws) * (W / ws);
  if (ws == 8) {
    if (smem > 48 * 1024)
      if (int rc = optin[0].ensure(window_attention_kernel<8>, smem)) return rc;
    CIAOSR_LAUNCH(window_attention_kernel<8>, (unsigned)grid, threads, smem, (cudaStream_t)stream, P);
  } else if (ws == 4) {
    if (smem > 48 * 1024)
      if (int rc = optin[1].ensure(window_attention_kernel<4>, smem)) return rc;
    CIAOSR_LAUNCH(window_attention_kernel<4>, (unsigned)grid, threads, smem, (cudaStream_t)stream, P);
  } else {
    if (smem > 48 * 1024)
      if (int rc = optin[2].ensure(window_attention_kernel<0>, smem)) return rc;
    CIAOSR_LAUNCH(window_attention_kernel<0>, (unsigned)grid, threads, smem, (cudaStream_t)stream, P);
  }
  return CIAOSR_OK;
}

extern "C" int ciaosr_window_attention_forward(const float* qkv, const float* bias_table, int B, int H, int W, int C,
                                               int heads, int ws, int shift, float scale, float* out, void* stream) {
  CIAOSR_REQUIRE(out != nullptr, CIAOSR_E_INVALID, "NULL pointer argument");
  return window_attention_launch(qkv, bias_table, B, H, W, C, heads, ws, shift, scale, out, nullptr, nullptr, 0, stream);
}

extern "C" int ciaosr_window_attention_split_forward(const float* qkv, const float* bias_table, int B, int H, int W,
                                                     int C, int heads, int ws, int shift, float scale,
                                                     uint16_t* out_hi, uint16_t* out_lo, int ld, void* stream) {
  CIAOSR_REQUIRE(out_hi && out_lo, CIAOSR_E_INVALID, "NULL pointer argument");
  return window_attention_launch(qkv, bias_table, B, H, W, C, heads, ws, shift, scale, nullptr,
                                 reinterpret_cast<split_t*>(out_hi), reinterpret_cast<split_t*>(out_lo), ld, stream);
}
