// RDN encoder on the tensor cores (SURVEY.md section 8f "next" #2: the encoder fast path).
//
// The reference keeps the encoder in PyTorch (mmedit's RDN, hoisted at ciaosr_net.py:314-318; forward
// ciaosr_net.py:321-342).  cuDNN offers two ways to run its 146 convolutions: TF32 (10 ms for the bench
// batch, but ~1e-3 relative error on the features, which the head turns into ~1e-3 output error: 10x the
// parity tolerance) or fp32 CUDA cores (47 ms).  This file runs them as implicit GEMMs on tcgen05 with the
// same bf16 hi/lo x3 scheme as the head: fp32-grade results at tensor-core speed.
//
// Layout trick ("linearised padded convolution"): activations live in HBM as NHWC with one zero pixel of
// padding on every side, pitch P = W + 2, stored as two bf16 tensors (hi, lo) [B*(H+2)*P pixels, C].  In that
// layout a 3x3 tap is a constant offset dy*P + dx in the linear pixel index, so the A operand of tap t for a
// tile of 128 consecutive linear pixels is ONE 2-D TMA box [128 pixels x 64 channels] at pixel
// tile*128 + dy*P + dx -- TMA writes it 128B-swizzled exactly as UMMA wants it (no thread touches A), zero-fills
// out-of-range pixels, and the padding pixels supply the convolution's zero padding.  Outputs for padding
// pixels are computed and discarded (4 % waste at W = 48); they are stored as zeros so buffers stay padded.
//
// Per layer: one persistent launch, K-slab = (64-channel block, tap), N = 64 output channels:
//   warp 0  producer: 2 TMA tensor loads (A_hi, A_lo) + 1 bulk copy (W_hi|W_lo, 16 KB) per K-slab
//   warp 1  UMMA issuer: 12 tcgen05.mma (M128 N64 K16) per K-slab, accumulators alternate per tile
//   warps 4-7 epilogue: bias (+ fp32 residual) (+ ReLU), bf16 hi/lo split, stores to the next layer's buffers
// Dense blocks never concatenate: every RDB owns one 576-channel buffer and each layer writes its 64
// channels into its slice; the LFF output goes straight into the next block's buffer and the global
// fusion buffer.  The residual trunk is kept in fp32.
#include <cuda.h>
#include <cstdlib>
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace ciaosr {

constexpr int CONV_THREADS = 256;
constexpr int CV_A_READY = BAR_A_READY, CV_A_FREE = BAR_A_FREE;

struct ConvDst { __nv_bfloat16* hi; __nv_bfloat16* lo; int ld; int choff; };

struct ConvParams {
  int n_tiles, P, HP2P, H, W, Np;      // Np = B*(H+2)*P valid linear pixels
  int ntaps, cblocks;                  // 9 or 1; Cin / 64
  const uint8_t* blob;                 // per K-slab: [W_hi 64x64 (8 KB)][W_lo (8 KB)], SW128
  const float* bias;                   // [64]
  const float* res32;                  // fp32 [Np_alloc, 64] residual or nullptr
  int relu;
  ConvDst d1, d2;                      // d2.hi == nullptr if unused
  float* out32;                        // fp32 [Np_alloc, 64] copy (residual trunk) or nullptr
  float* out_nchw;                     // final feature [B,64,H,W] or nullptr
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// ---- epilogue shared by both convolution kernels: one thread per output pixel ------------------------------
struct ConvEpiBars { const float* bias_s; uint32_t d_ready[2]; uint32_t d_free[2]; bool stacked; };
__device__ __forceinline__ void conv_epilogue(const ConvEpiBars& s, const ConvParams& P, uint32_t tmem_base, int warp,
                                              int lane) {
    const int row = threadIdx.x - EPI_T0;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      const long long g = (long long)tile * ROWS + row;
      const int rr = (int)(g % P.HP2P), b = (int)(g / P.HP2P);
      const int yy = rr / P.P, xx = rr - yy * P.P;
      const bool valid = g < P.Np && yy >= 1 && yy <= P.H && xx >= 1 && xx <= P.W;
      mbar_wait(s.d_ready[d], n & 1, 450);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32(lane_taddr + d * 256 + c0, v);
        if (s.stacked) {                 // the A.W_lo partial sums live 64 columns further
          float u[32];
          tmem_ld32(lane_taddr + d * 256 + 64 + c0, u);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += u[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += s.bias_s[c0 + i];
        if (P.res32 != nullptr) {
          const float4* rp = reinterpret_cast<const float4*>(P.res32 + g * 64 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 q = __ldg(rp + j);
            v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = valid ? (P.relu ? fmaxf(v[i], 0.0f) : v[i]) : 0.0f;
        uint32_t h[16], l[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
        if (P.d1.hi != nullptr) {
          uint4* ph = reinterpret_cast<uint4*>(P.d1.hi + g * P.d1.ld + P.d1.choff + c0);
          uint4* pl = reinterpret_cast<uint4*>(P.d1.lo + g * P.d1.ld + P.d1.choff + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ph[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
            pl[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
          }
        }
        if (P.d2.hi != nullptr) {
          uint4* ph = reinterpret_cast<uint4*>(P.d2.hi + g * P.d2.ld + P.d2.choff + c0);
          uint4* pl = reinterpret_cast<uint4*>(P.d2.lo + g * P.d2.ld + P.d2.choff + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ph[j] = make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
            pl[j] = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
          }
        }
        if (P.out32 != nullptr) {
          float4* po = reinterpret_cast<float4*>(P.out32 + g * 64 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) po[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (P.out_nchw != nullptr && valid) {
          float* po = P.out_nchw + (((long long)b * 64 + c0) * P.H + (yy - 1)) * P.W + (xx - 1);
#pragma unroll
          for (int i = 0; i < 32; ++i) po[(long long)i * P.H * P.W] = v[i];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s.d_free[d]);
    }
  }

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const ConvParams P, const __grid_constant__ CUtensorMap map_hi,
               const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcShared s = tc_carve(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) s.consts[threadIdx.x] = P.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_at(s, BAR_W_FULL + i), 1); mbar_init(bar_at(s, BAR_W_EMPTY + i), 1);
      mbar_init(bar_at(s, CV_A_READY + i), 1); mbar_init(bar_at(s, CV_A_FREE + i), 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(bar_at(s, BAR_D_READY + i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_at(s, BAR_D_FREE + i), 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(smem) + SM_SLOT, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + SM_SLOT);
  const int nslabs = P.ntaps * P.cblocks;

  if (warp == 0) {
    // ---- producer: A through TMA tensor loads, W through bulk copies; both rings are 4 K-slabs deep ----
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int sl = 0; sl < nslabs; ++sl, ++cnt) {
        const int slot = cnt & 3;
        const uint32_t par = ((cnt >> 2) & 1) ^ 1;
        const int cb = sl / P.ntaps, tap = sl - cb * P.ntaps;
        const int dy = P.ntaps == 9 ? tap / 3 - 1 : 0, dx = P.ntaps == 9 ? tap % 3 - 1 : 0;
        mbar_wait(bar_at(s, CV_A_FREE + slot), par, 400 + slot);
        if (lane == 0) {
          const uint32_t full = bar_at(s, CV_A_READY + slot);
          mbar_arrive_expect_tx(full, 2 * SLAB_BYTES);
          const int pix = tile * ROWS + dy * P.P + dx;
          tma_load_2d(s.a_hi + slot * SLAB_BYTES, &map_hi, cb * 64, pix, full);
          tma_load_2d(s.a_lo + slot * SLAB_BYTES, &map_lo, cb * 64, pix, full);
        }
        __syncwarp();
        mbar_wait(bar_at(s, BAR_W_EMPTY + slot), par, 410 + slot);
        if (lane == 0) {
          const uint32_t full = bar_at(s, BAR_W_FULL + slot);
          mbar_arrive_expect_tx(full, SLAB_BYTES);
          bulk_g2s(s.w + slot * SLAB_BYTES, P.blob + (size_t)sl * SLAB_BYTES, SLAB_BYTES, full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer ----
    const uint32_t idesc = make_idesc_bf16(ROWS, 64);
    uint32_t cnt = 0, job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      mbar_wait(bar_at(s, BAR_D_FREE + d), (n + 1) & 1, 420);
      tc_fence_after();
      const uint32_t dcol = tmem_base + d * 256;
      for (int sl = 0; sl < nslabs; ++sl, ++cnt) {
        const int slot = cnt & 3;
        const uint32_t par = (cnt >> 2) & 1;
        mbar_wait(bar_at(s, CV_A_READY + slot), par, 430 + slot);
        mbar_wait(bar_at(s, BAR_W_FULL + slot), par, 440 + slot);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_hi = desc_lo(s.a_hi + slot * SLAB_BYTES), a_lo = desc_lo(s.a_lo + slot * SLAB_BYTES);
          const uint32_t b_hi = desc_lo(s.w + slot * SLAB_BYTES), b_lo = desc_lo(s.w + slot * SLAB_BYTES + SLAB_BYTES / 2);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_lo(dcol, a_lo + 2 * ks, b_hi + 2 * ks, idesc, (sl | ks) != 0 ? 1u : 0u);
            umma_lo(dcol, a_hi + 2 * ks, b_hi + 2 * ks, idesc, 1u);
            umma_lo(dcol, a_hi + 2 * ks, b_lo + 2 * ks, idesc, 1u);
          }
          umma_commit(bar_at(s, CV_A_FREE + slot));
          umma_commit(bar_at(s, BAR_W_EMPTY + slot));
        }
        __syncwarp();
      }
      if (lane == 0) umma_commit(bar_at(s, BAR_D_READY + 2 * d));
      __syncwarp();
    }
  } else if (warp >= 4) {
    const ConvEpiBars eb{s.consts, {bar_at(s, BAR_D_READY), bar_at(s, BAR_D_READY + 2)},
                         {bar_at(s, BAR_D_FREE), bar_at(s, BAR_D_FREE + 1)}, false};
    conv_epilogue(eb, P, tmem_base, warp, lane);
  }
  tc_teardown<1>(tmem_base);
}

// ---- 3x3 convolution with halo staging ("conv3") ------------------------------------------------------------
// The nine taps of a 3x3 convolution read the same pixels shifted by dy*P + dx rows of the linearised
// layout, so ONE TMA box of HR = 128 + 2P + 2 rows per (tile, 64-channel block) serves all of them: tap
// (dy, dx) is the same smem tile read through a UMMA descriptor whose start address is advanced by
// r0 = (dy+1)*P + (dx+1) rows of 128 bytes.  The 128B swizzle is a function of the absolute smem address
// (bits 4-6 ^= bits 7-9), for the TMA write and the UMMA read alike, so a start address that is not
// 1024-byte aligned needs nothing else: the descriptor's base-offset field stays 0 (measured on B200:
// base offset = r0 & 7, the other reading of the PTX text, gives wrong sums).  L2 -> SM traffic per (tile, channel block) drops from
// 9 x (32 KB A + 16 KB W) = 432 KB to 2 x HR x 128 B (59 KB at P = 50) + 144 KB of weights.
// Needs HR <= 256 (TMA box limit), i.e. W <= 61; wider images use the per-tap kernel above.
constexpr int C3_A_HALF = 256 * 128;                     // bytes reserved per hi (or lo) halo tile
constexpr int C3_A_STAGES = 2, C3_W_STAGES = 5;
constexpr int C3_SM_W = C3_A_STAGES * 2 * C3_A_HALF;     // 128 KB of A stages first
constexpr int C3_SM_CONST = C3_SM_W + C3_W_STAGES * SLAB_BYTES;
constexpr int C3_SM_BAR = C3_SM_CONST + 64 * 4;
constexpr int C3_W_FULL = 0, C3_W_EMPTY = 5, C3_A_READY = 10, C3_A_FREE = 12, C3_D_READY = 14, C3_D_FREE = 16,
              C3_NBARS = 18;
constexpr int C3_SM_SLOT = C3_SM_BAR + C3_NBARS * 8;
constexpr int C3_SM_TOTAL = C3_SM_SLOT + 16;

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3_tc_kernel(const ConvParams P, const int HR, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + C3_SM_BAR;
  float* bias_s = reinterpret_cast<float*>(smem + C3_SM_CONST);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) bias_s[threadIdx.x] = P.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < C3_W_STAGES; ++i) { mbar_init(bars + 8 * (C3_W_FULL + i), 1); mbar_init(bars + 8 * (C3_W_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + 8 * (C3_A_READY + i), 1); mbar_init(bars + 8 * (C3_A_FREE + i), 1);
      mbar_init(bars + 8 * (C3_D_READY + i), 1); mbar_init(bars + 8 * (C3_D_FREE + i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(sbase + C3_SM_SLOT, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + C3_SM_SLOT);

  if (warp == 0) {
    // ---- producer: one halo box (hi, lo) per channel block, one 16 KB weight slab per tap ----
    uint32_t acnt = 0, wcnt = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt & 1;
        mbar_wait(bars + 8 * (C3_A_FREE + ast), ((acnt >> 1) & 1) ^ 1, 400 + ast);
        if (lane == 0) {
          const uint32_t full = bars + 8 * (C3_A_READY + ast);
          mbar_arrive_expect_tx(full, 2u * (uint32_t)HR * 128u);
          const int pix = tile * ROWS - P.P - 1;
          tma_load_2d(sbase + (2 * ast) * C3_A_HALF, &map_hi, cb * 64, pix, full);
          tma_load_2d(sbase + (2 * ast + 1) * C3_A_HALF, &map_lo, cb * 64, pix, full);
        }
        __syncwarp();
        for (int tap = 0; tap < 9; ++tap, ++wcnt) {
          const int wst = wcnt % C3_W_STAGES;
          mbar_wait(bars + 8 * (C3_W_EMPTY + wst), ((wcnt / C3_W_STAGES) & 1) ^ 1, 410 + wst);
          if (lane == 0) {
            const uint32_t full = bars + 8 * (C3_W_FULL + wst);
            mbar_arrive_expect_tx(full, SLAB_BYTES);
            bulk_g2s(sbase + C3_SM_W + wst * SLAB_BYTES, P.blob + (size_t)(cb * 9 + tap) * SLAB_BYTES, SLAB_BYTES, full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer ----
    const uint32_t idesc = make_idesc_bf16(ROWS, 128);   // D[:, 0:64] = A.W_hi, D[:, 64:128] = A.W_lo (summed in the epilogue)
    uint32_t acnt = 0, wcnt = 0, job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      mbar_wait(bars + 8 * (C3_D_FREE + d), (n + 1) & 1, 420);
      tc_fence_after();
      const uint32_t dcol = tmem_base + d * 256;
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt & 1;
        mbar_wait(bars + 8 * (C3_A_READY + ast), (acnt >> 1) & 1, 430 + ast);
        for (int tap = 0; tap < 9; ++tap, ++wcnt) {
          const int wst = wcnt % C3_W_STAGES;
          mbar_wait(bars + 8 * (C3_W_FULL + wst), (wcnt / C3_W_STAGES) & 1, 440 + wst);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t r0 = (uint32_t)((tap / 3) * P.P + tap % 3);
            const uint32_t a_hi = desc_lo(sbase + (2 * ast) * C3_A_HALF + r0 * 128u);
            const uint32_t a_lo = desc_lo(sbase + (2 * ast + 1) * C3_A_HALF + r0 * 128u);
            const uint32_t b = desc_lo(sbase + C3_SM_W + wst * SLAB_BYTES);     // 128 rows: [W_hi (64); W_lo (64)]
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_lo(dcol, a_lo + 2 * ks, b + 2 * ks, idesc, (cb | tap | ks) != 0 ? 1u : 0u);
              umma_lo(dcol, a_hi + 2 * ks, b + 2 * ks, idesc, 1u);
            }
            umma_commit(bars + 8 * (C3_W_EMPTY + wst));
            if (tap == 8) umma_commit(bars + 8 * (C3_A_FREE + ast));
          }
          __syncwarp();
        }
      }
      if (lane == 0) umma_commit(bars + 8 * (C3_D_READY + d));
      __syncwarp();
    }
  } else if (warp >= 4) {
    const ConvEpiBars eb{bias_s, {bars + 8 * C3_D_READY, bars + 8 * (C3_D_READY + 1)},
                         {bars + 8 * C3_D_FREE, bars + 8 * (C3_D_FREE + 1)}, true};
    conv_epilogue(eb, P, tmem_base, warp, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}


// ---- convolution with the weights in tensor memory ("convw") -------------------------------------------------
// Measured on B200 (profiles/r01e): a tcgen05.mma whose A operand comes from shared memory is bound by
// operand fetch, ~85 B/cycle/SM for A and B together, not by the tensor pipe, once N is small: with the 64
// output channels as N (the kernels above) every M128 x N64 x K16 instruction fetches 6 KB for 32 cycles of
// math and takes 78.  So the roles are swapped here:
//   A (M = 128) = the weights of one (channel block, tap) K-slab, rows [W_hi (64 out channels); W_lo (64)],
//                 staged in TENSOR MEMORY by four "stager" warps (ld.global -> tcgen05.st; the weights never
//                 touch shared memory), 8 slabs deep;
//   B (N = 128) = 128 consecutive pixels of the linearised padded activation (hi or lo) in shared memory: the
//                 halo box of conv3 above, one TMA load per (tile, channel block), row-shifted per tap;
//   D[lane, col] = two accumulators of 128 columns: lanes 0-63 hold W_hi.X, lanes 64-127 hold W_lo.X with
//                 X = X_hi + X_lo (all four bf16 products: 2 instructions per K16, 64 cycles each, 64 B/cycle).
// The epilogue transposes D through shared memory (out[pixel][c] = D[c][pixel] + D[64 + c][pixel] + bias),
// applies residual / ReLU, splits to bf16 hi/lo and writes the next layer's NHWC buffers with 128-byte rows.
// Any image width: narrow images (128 + 2P + 2 <= 256 rows) use one halo box, wider ones three row-band
// boxes of 136 rows (one per dy); 1x1 convolutions use one 128-row box.
constexpr int CW_THREADS = 512;           // warps: 0 TMA, 1 UMMA issuer, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 stagers
constexpr int CW_W_STAGES = 8;            // weight slabs resident in TMEM: 8 x 32 columns after the accumulators
constexpr int CW_T_LD = 132;              // floats per pixel row of the transpose buffer (padded)
constexpr int CW_A_READY = 0, CW_A_FREE = 2, CW_W_FULL = 4, CW_W_EMPTY = 12, CW_D_READY = 20, CW_D_FREE = 22,
              CW_NBARS = 24;
constexpr int CW_FIXED_BYTES = 32 * CW_T_LD * 4 + 64 * 4 + CW_NBARS * 8 + 16;

struct ConvWParams {
  ConvParams c;
  const uint8_t* wrows;       // per K-slab (cblock, tap): 128 rows [W_hi; W_lo] x 64 bf16 as [chunk of 8][row][8]
  int half_bytes;             // smem bytes of one (hi or lo) stage half = nbox * box_rows * 128
  int nbox, box_rows;         // TMA boxes per stage half and rows per box
  int box_pix0[3];            // first pixel of box b relative to the tile start
  int tap_off[9];             // byte offset of tap t inside a stage half
};

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(CW_THREADS, 1)
convw_tc_kernel(const ConvWParams Q, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams& P = Q.c;
  const uint32_t sbase = smem_u32(smem);
  const int fixed0 = 4 * Q.half_bytes;                       // [stage][hi|lo] halves first
  float* tbuf = reinterpret_cast<float*>(smem + fixed0);     // [32 pixels][CW_T_LD]
  float* bias_s = tbuf + 32 * CW_T_LD;
  const uint32_t bars = sbase + fixed0 + 32 * CW_T_LD * 4 + 64 * 4;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + fixed0 + 32 * CW_T_LD * 4 + 64 * 4 + CW_NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) bias_s[threadIdx.x] = P.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + 8 * (CW_A_READY + i), 1); mbar_init(bars + 8 * (CW_A_FREE + i), 1);
      mbar_init(bars + 8 * (CW_D_READY + i), 1); mbar_init(bars + 8 * (CW_D_FREE + i), 4);
    }
    for (int i = 0; i < CW_W_STAGES; ++i) { mbar_init(bars + 8 * (CW_W_FULL + i), 8); mbar_init(bars + 8 * (CW_W_EMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot);
  const int units = P.cblocks * P.ntaps;                     // K-slabs per tile

  if (warp == 0) {
    // ---- TMA producer: the activation boxes of one channel block (hi and lo) per stage ----
    uint32_t acnt = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt & 1;
        mbar_wait(bars + 8 * (CW_A_FREE + ast), ((acnt >> 1) & 1) ^ 1, 400 + ast);
        if (lane == 0) {
          const uint32_t full = bars + 8 * (CW_A_READY + ast);
          mbar_arrive_expect_tx(full, 2u * (uint32_t)Q.half_bytes);
          const uint32_t dst = sbase + (uint32_t)(2 * ast) * Q.half_bytes;
          for (int b = 0; b < Q.nbox; ++b) {
            const int pix = tile * ROWS + Q.box_pix0[b];
            tma_load_2d(dst + b * Q.box_rows * 128, &map_hi, cb * 64, pix, full);
            tma_load_2d(dst + Q.half_bytes + b * Q.box_rows * 128, &map_lo, cb * 64, pix, full);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer ----
    const uint32_t idesc = make_idesc_bf16(ROWS, 128);
    uint32_t acnt = 0, wcnt = 0, job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      mbar_wait(bars + 8 * (CW_D_FREE + d), (n + 1) & 1, 420);
      tc_fence_after();
      const uint32_t dcol = tmem_base + d * 128;
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt & 1;
        mbar_wait(bars + 8 * (CW_A_READY + ast), (acnt >> 1) & 1, 430 + ast);
        const uint32_t x_hi = sbase + (uint32_t)(2 * ast) * Q.half_bytes, x_lo = x_hi + Q.half_bytes;
        for (int tap = 0; tap < P.ntaps; ++tap, ++wcnt) {
          const int wst = wcnt % CW_W_STAGES;
          mbar_wait(bars + 8 * (CW_W_FULL + wst), (wcnt / CW_W_STAGES) & 1, 440 + wst);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_hi = desc_lo(x_hi + Q.tap_off[tap]), b_lo = desc_lo(x_lo + Q.tap_off[tap]);
            const uint32_t a = tmem_base + 256 + wst * 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_ts(dcol, a + 8 * ks, b_lo + 2 * ks, idesc, (cb | tap | ks) != 0 ? 1u : 0u);
              umma_ts(dcol, a + 8 * ks, b_hi + 2 * ks, idesc, 1u);
            }
            umma_commit(bars + 8 * (CW_W_EMPTY + wst));
            if (tap == P.ntaps - 1) umma_commit(bars + 8 * (CW_A_FREE + ast));
          }
          __syncwarp();
        }
      }
      if (lane == 0) umma_commit(bars + 8 * (CW_D_READY + d));
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ---- weight stagers: row r of every K-slab, global -> registers -> tensor memory ----
    // warps 8-11 stage K columns 0-31 of every slab (TMEM columns 0-15 of its stage), warps 12-15 the rest;
    // four register buffers per thread: a load is issued three K-slabs (>= 1500 tensor-core cycles) ahead
    const int r = (warp & 3) * 32 + lane, khalf = (warp - 8) >> 2;
    const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + khalf * 16;
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) ++my_tiles;
    const long long total = (long long)my_tiles * units;
    const uint4* src0 = reinterpret_cast<const uint4*>(Q.wrows) + khalf * 512 + r;    // + unit * 1024 + j * 128
    auto load = [&](long long i, uint32_t (&dst)[16]) {
      if (i >= total) return;
      const uint4* sp = src0 + (size_t)(i % units) * 1024;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q = __ldg(sp + j * 128);
        dst[4 * j] = q.x; dst[4 * j + 1] = q.y; dst[4 * j + 2] = q.z; dst[4 * j + 3] = q.w;
      }
    };
    auto stage = [&](long long i, const uint32_t (&src)[16]) {
      if (i >= total) return;
      const int wst = (int)(i % CW_W_STAGES);
      mbar_wait(bars + 8 * (CW_W_EMPTY + wst), (uint32_t)(((i / CW_W_STAGES) & 1) ^ 1), 460 + wst);
      tc_fence_after();
      tmem_st16(taddr0 + wst * 32, src);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * (CW_W_FULL + wst));
    };
    uint32_t b0[16], b1[16], b2[16], b3[16];
    load(0, b0);
    load(1, b1);
    load(2, b2);
    for (long long i = 0; i < total; i += 4) {
      load(i + 3, b3); stage(i, b0);
      load(i + 4, b0); stage(i + 1, b1);
      load(i + 5, b1); stage(i + 2, b2);
      load(i + 6, b2); stage(i + 3, b3);
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue: transpose through shared memory, one 32-pixel chunk at a time ----
    const int t = threadIdx.x - EPI_T0;                      // 0..127
    const int lane_row = t;                                  // TMEM lane = stacked weight row
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int pp = t >> 2, c0 = (t & 3) * 16;                // pixel within the chunk, first of 16 channels
    uint32_t job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      mbar_wait(bars + 8 * (CW_D_READY + d), n & 1, 450);
      tc_fence_after();
      for (int ch = 0; ch < 4; ++ch) {
        float v[32];
        tmem_ld32(lane_taddr + d * 128 + ch * 32, v);
        if (ch == 3) {                                       // accumulator fully read: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + 8 * (CW_D_FREE + d));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");       // previous chunk's readers are done
#pragma unroll
        for (int j = 0; j < 32; ++j) tbuf[j * CW_T_LD + lane_row] = v[j];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const long long g = (long long)tile * ROWS + ch * 32 + pp;
        const int rr = (int)(g % P.HP2P), b = (int)(g / P.HP2P);
        const int yy = rr / P.P, xx = rr - yy * P.P;
        const bool valid = g < P.Np && yy >= 1 && yy <= P.H && xx >= 1 && xx <= P.W;
        float o[16];
        const float4* t0 = reinterpret_cast<const float4*>(tbuf + pp * CW_T_LD + c0);
        const float4* t1 = reinterpret_cast<const float4*>(tbuf + pp * CW_T_LD + 64 + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 a = t0[j], l = t1[j];
          const float4 bb = reinterpret_cast<const float4*>(bias_s + c0)[j];
          o[4 * j] = a.x + l.x + bb.x; o[4 * j + 1] = a.y + l.y + bb.y;
          o[4 * j + 2] = a.z + l.z + bb.z; o[4 * j + 3] = a.w + l.w + bb.w;
        }
        if (P.res32 != nullptr) {
          const float4* rp = reinterpret_cast<const float4*>(P.res32 + g * 64 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 q = __ldg(rp + j);
            o[4 * j] += q.x; o[4 * j + 1] += q.y; o[4 * j + 2] += q.z; o[4 * j + 3] += q.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = valid ? (P.relu ? fmaxf(o[i], 0.0f) : o[i]) : 0.0f;
        uint32_t h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split2(o[2 * i], o[2 * i + 1], h[i], l[i]);
        if (P.d1.hi != nullptr) {
          uint4* ph = reinterpret_cast<uint4*>(P.d1.hi + g * P.d1.ld + P.d1.choff + c0);
          uint4* pl = reinterpret_cast<uint4*>(P.d1.lo + g * P.d1.ld + P.d1.choff + c0);
          ph[0] = make_uint4(h[0], h[1], h[2], h[3]); ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
          pl[0] = make_uint4(l[0], l[1], l[2], l[3]); pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
        }
        if (P.d2.hi != nullptr) {
          uint4* ph = reinterpret_cast<uint4*>(P.d2.hi + g * P.d2.ld + P.d2.choff + c0);
          uint4* pl = reinterpret_cast<uint4*>(P.d2.lo + g * P.d2.ld + P.d2.choff + c0);
          ph[0] = make_uint4(h[0], h[1], h[2], h[3]); ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
          pl[0] = make_uint4(l[0], l[1], l[2], l[3]); pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
        }
        if (P.out32 != nullptr) {
          float4* po = reinterpret_cast<float4*>(P.out32 + g * 64 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) po[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }
        if (P.out_nchw != nullptr && valid) {
          float* po = P.out_nchw + (((long long)b * 64 + c0) * P.H + (yy - 1)) * P.W + (xx - 1);
#pragma unroll
          for (int i = 0; i < 16; ++i) po[(long long)i * P.H * P.W] = o[i];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// weight packing for convw: conv weight [64, Cin, kh, kw] -> per K-slab (cblock, tap) 128 rows x 64 bf16, chunk-major
__global__ void rdn_pack_rows_kernel(__nv_bfloat16* __restrict__ dst, const float* __restrict__ w, int Cin, int ntaps) {
  const int cblocks = Cin / 64;
  const long long total = (long long)cblocks * ntaps * 64 * 64;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % 64), n = (int)((i / 64) % 64);
  const int sl = (int)(i / 4096);
  const int cb = sl / ntaps, tap = sl % ntaps;
  const float v = w[((long long)n * Cin + cb * 64 + k) * ntaps + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  // chunk-major: [K-slab][16-byte chunk j = k / 8][row][k % 8], so that the 32 rows a stager warp loads with
  // one instruction are 512 contiguous bytes
  __nv_bfloat16* ub = dst + (size_t)sl * 128 * 64;
  ub[((k >> 3) * 128 + n) * 8 + (k & 7)] = hi;
  ub[((k >> 3) * 128 + 64 + n) * 8 + (k & 7)] = lo;
}

// ---- weight packing: conv weight [64, Cin, kh, kw] -> per K-slab (cblock, tap): hi 64x64 | lo 64x64 ----------
__global__ void rdn_pack_conv_kernel(uint8_t* __restrict__ dst, const float* __restrict__ w, int Cin, int ntaps) {
  const int cblocks = Cin / 64;
  const long long total = (long long)cblocks * ntaps * 64 * 64;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % 64), n = (int)((i / 64) % 64);
  const int sl = (int)(i / 4096);
  const int cb = sl / ntaps, tap = sl % ntaps;
  const float v = w[((long long)n * Cin + cb * 64 + k) * ntaps + tap];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  uint8_t* ub = dst + (size_t)sl * SLAB_BYTES;
  const uint32_t off = sw128_offset(n, k);
  *reinterpret_cast<__nv_bfloat16*>(ub + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(ub + SLAB_BYTES / 2 + off) = lo;
}

// ---- first convolution (3 input channels): functor GEMM straight from the NCHW image ------------------------
struct Sfe1Gen {          // A[(b,y,x), k = ci*9 + tap], K = 27 (weights are [64, 3, 3, 3] = [64, 27] row-major)
  const float* x; int H, W;
  struct Row { int b, y, xx; };
  __device__ __forceinline__ Row row(long long m) const {
    const int hw = (int)(m % ((long long)H * W));
    return Row{(int)(m / ((long long)H * W)), hw / W, hw % W};
  }
  __device__ __forceinline__ void fill(Row& r, long long, int k0, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = k0 + i;
      float q = 0.0f;
      if (k < 27) {
        const int ci = k / 9, t = k % 9;
        const int yy = r.y + t / 3 - 1, xx = r.xx + t % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) q = __ldg(x + (((long long)r.b * 3 + ci) * H + yy) * W + xx);
      }
      v[i] = q;
    }
  }
};
struct Sfe1Src {          // B[n, k] = w[n*27 + k]
  const float* w;
  __device__ __forceinline__ float operator()(int, int n, int k) const { return w[n * 27 + k]; }
};
struct Sfe1Epi {          // + bias -> padded bf16 hi/lo + padded fp32
  __nv_bfloat16* hi; __nv_bfloat16* lo; float* out32; const float* bias; int H, W, P, HP2P;
  __device__ __forceinline__ void store(const Sfe1Gen::Row& r, long long, int n0, const float (&v)[32]) const {
    if (n0 >= 64) return;
    const long long g = (long long)r.b * HP2P + (r.y + 1) * P + (r.xx + 1);
    float t[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) t[i] = v[i] + bias[n0 + i];
    uint4* ph = reinterpret_cast<uint4*>(hi + g * 64 + n0);
    uint4* pl = reinterpret_cast<uint4*>(lo + g * 64 + n0);
    float4* po = reinterpret_cast<float4*>(out32 + g * 64 + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split2(t[8 * j + 2 * i], t[8 * j + 2 * i + 1], h[i], l[i]);
      ph[j] = make_uint4(h[0], h[1], h[2], h[3]);
      pl[j] = make_uint4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) po[j] = make_float4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
  }
};

// ---- host side ----------------------------------------------------------------------------------------------
struct RdnGeom { int nb, nl, C; };       // blocks, layers per block, channels (= growth = 64)

struct RdnPlanLayout {
  // byte offsets of the per-layer weight blobs and fp32 biases inside the plan buffer
  size_t sfe1_blob, sfe2, gff0, gff1, dense0, lff0;    // dense/lff: consecutive per block
  size_t bias0;                                         // floats: [sfe1, sfe2, dense..., lff..., gff0, gff1] x 64
  size_t dense_stride_block, total;
  size_t rows_delta;                                    // blob offset + rows_delta = its row-major twin (convw)
  size_t dense_off[64];                                 // offset of layer l inside a block's dense blob
};

static int rdn_check(const ciaosr_rdn_desc* d) {
  CIAOSR_REQUIRE(d != nullptr && d->abi_version == CIAOSR_ABI_VERSION, CIAOSR_E_INVALID, "bad rdn desc / abi_version");
  CIAOSR_REQUIRE(d->mid_channels == 64 && d->channel_growth == 64, CIAOSR_E_INVALID,
                 "native RDN needs mid_channels == channel_growth == 64 (got %d, %d)", d->mid_channels,
                 d->channel_growth);
  CIAOSR_REQUIRE(d->num_blocks >= 1 && d->num_blocks <= 32 && d->num_layers >= 1 && d->num_layers <= 16,
                 CIAOSR_E_INVALID, "unsupported RDN depth %d x %d", d->num_blocks, d->num_layers);
  CIAOSR_REQUIRE(d->sfe1_w && d->sfe1_b && d->sfe2_w && d->sfe2_b && d->dense_w && d->dense_b && d->lff_w &&
                     d->lff_b && d->gff0_w && d->gff0_b && d->gff1_w && d->gff1_b,
                 CIAOSR_E_INVALID, "NULL RDN parameter pointer");
  return CIAOSR_OK;
}

static RdnPlanLayout rdn_layout(const ciaosr_rdn_desc* d) {
  RdnPlanLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += (bytes + 255) / 256 * 256; return r; };
  L.sfe1_blob = take(tc_operand_blob_bytes(1, 1));
  L.sfe2 = take((size_t)9 * SLAB_BYTES);
  size_t blk = 0;
  for (int l = 0; l < d->num_layers; ++l) { L.dense_off[l] = blk; blk += (size_t)9 * (1 + l) * SLAB_BYTES; }
  L.dense_stride_block = blk;
  L.dense0 = take(blk * d->num_blocks);
  L.lff0 = take((size_t)(1 + d->num_layers) * SLAB_BYTES * d->num_blocks);
  L.gff0 = take((size_t)d->num_blocks * SLAB_BYTES);
  L.gff1 = take((size_t)9 * SLAB_BYTES);
  const int nconv = 2 + d->num_blocks * (d->num_layers + 1) + 2;
  L.bias0 = take((size_t)nconv * 64 * sizeof(float));
  L.rows_delta = off;
  L.total = 2 * off;
  return L;
}

__global__ void rdn_copy64_kernel(float* dst, const float* src) { dst[threadIdx.x] = src[threadIdx.x]; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D map over a bf16 [pixels, channels] tensor: box = 64 channels x `rows` pixels, 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* m, void* base, long long pixels, int channels, int rows = 128) {
  EncodeTiledFn enc = get_encode();
  CIAOSR_REQUIRE(enc != nullptr, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)channels, (cuuint64_t)pixels};
  const cuuint64_t gstride[1] = {(cuuint64_t)channels * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CIAOSR_REQUIRE(r == CUDA_SUCCESS, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CIAOSR_OK;
}

struct RdnWs {
  long long Np, Npa;   // valid / allocated (multiple of 128) linear pixels
  __nv_bfloat16 *f1h, *f1l, *rbh[2], *rbl[2], *gfh, *gfl, *g1h, *g1l;
  float *f1_32, *xr[2];
};
static RdnWs rdn_carve(Arena& a, const ciaosr_rdn_desc* d, int B, int H, int W) {
  RdnWs w;
  w.Np = (long long)B * (H + 2) * (W + 2);
  w.Npa = (w.Np + ROWS - 1) / ROWS * ROWS;
  const size_t n = (size_t)w.Npa;
  const int cb = 64 * (1 + d->num_layers);
  w.f1h = a.take<__nv_bfloat16>(n * 64); w.f1l = a.take<__nv_bfloat16>(n * 64);
  for (int i = 0; i < 2; ++i) { w.rbh[i] = a.take<__nv_bfloat16>(n * cb); w.rbl[i] = a.take<__nv_bfloat16>(n * cb); }
  w.gfh = a.take<__nv_bfloat16>(n * 64 * d->num_blocks); w.gfl = a.take<__nv_bfloat16>(n * 64 * d->num_blocks);
  w.g1h = a.take<__nv_bfloat16>(n * 64); w.g1l = a.take<__nv_bfloat16>(n * 64);
  w.f1_32 = a.take<float>(n * 64);
  for (int i = 0; i < 2; ++i) w.xr[i] = a.take<float>(n * 64);
  return w;
}

static int conv_attrs() {
  static bool attr_set = false;
  if (!attr_set) {
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr_set = true;
  }
  return CIAOSR_OK;
}
static int launch_conv(const ConvParams& P, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
  int rc = conv_attrs();
  if (rc) return rc;
  CIAOSR_LAUNCH(conv_tc_kernel, tc_grid_size(P.n_tiles), CONV_THREADS, SM_TOTAL, st, P, mh, ml);
  return CIAOSR_OK;
}
static int launch_convw(const ConvParams& P, const uint8_t* wrows, int pitch, CUtensorMap (*maps)[2], cudaStream_t st) {
  // maps[0] = halo box (or 128-row box for 1x1), maps[1] = 136-row band box; chosen here
  ConvWParams q{};
  q.c = P; q.wrows = wrows;
  int which = 0;
  if (P.ntaps == 1) {
    q.nbox = 1; q.box_rows = ROWS; q.box_pix0[0] = 0; q.tap_off[0] = 0; which = 2;
  } else {
    const int hr = (ROWS + 2 * pitch + 2 + 7) / 8 * 8;
    if (hr <= 256) {
      q.nbox = 1; q.box_rows = hr; q.box_pix0[0] = -pitch - 1;
      for (int t = 0; t < 9; ++t) q.tap_off[t] = ((t / 3) * pitch + t % 3) * 128;
      which = 0;
    } else {
      q.nbox = 3; q.box_rows = 136;
      for (int b = 0; b < 3; ++b) q.box_pix0[b] = (b - 1) * pitch - 1;
      for (int t = 0; t < 9; ++t) q.tap_off[t] = (t / 3) * 136 * 128 + (t % 3) * 128;
      which = 1;
    }
  }
  q.half_bytes = (q.nbox * q.box_rows * 128 + 1023) / 1024 * 1024;
  const int smem_bytes = 4 * q.half_bytes + CW_FIXED_BYTES;
  static int max_set = 0;
  if (smem_bytes > max_set) {
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(convw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    max_set = smem_bytes;
  }
  CIAOSR_LAUNCH(convw_tc_kernel, tc_grid_size(P.n_tiles), CW_THREADS, smem_bytes, st, q, maps[which][0], maps[which][1]);
  return CIAOSR_OK;
}
// rows of the halo box of conv3_tc_kernel for pitch P (0: image too wide for one TMA box)
static int conv3_halo_rows(int P) {
  const int hr = (ROWS + 2 * P + 2 + 7) / 8 * 8;
  return hr <= 256 ? hr : 0;
}
static int launch_conv3(const ConvParams& P, int HR, const CUtensorMap& mh, const CUtensorMap& ml, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(conv3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C3_SM_TOTAL));
    attr_set = true;
  }
  CIAOSR_LAUNCH(conv3_tc_kernel, tc_grid_size(P.n_tiles), CONV_THREADS, C3_SM_TOTAL, st, P, HR, mh, ml);
  return CIAOSR_OK;
}

}  // namespace ciaosr

using namespace ciaosr;

extern "C" {

int ciaosr_rdn_plan_bytes(const ciaosr_rdn_desc* desc, size_t* bytes) {
  CIAOSR_REQUIRE(bytes != nullptr, CIAOSR_E_INVALID, "bytes is NULL");
  int rc = rdn_check(desc);
  if (rc) return rc;
  *bytes = rdn_layout(desc).total;
  return CIAOSR_OK;
}

int ciaosr_rdn_plan_init(const ciaosr_rdn_desc* d, void* plan, size_t plan_bytes, void* stream) {
  int rc = rdn_check(d);
  if (rc) return rc;
  const RdnPlanLayout L = rdn_layout(d);
  CIAOSR_REQUIRE(plan != nullptr && ((uintptr_t)plan % 256) == 0 && plan_bytes >= L.total, CIAOSR_E_WORKSPACE,
                 "RDN plan buffer too small or misaligned: need %zu, have %zu", L.total, plan_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* p = (uint8_t*)plan;
  float* bias = reinterpret_cast<float*>(p + L.bias0);
  int bi = 0;
  auto pack = [&](size_t off, const float* w, const float* b, int Cin, int ntaps) -> int {
    const long long total = (long long)(Cin / 64) * ntaps * 4096;
    CIAOSR_LAUNCH(rdn_pack_conv_kernel, cdiv(total, 256), 256, 0, st, p + off, w, Cin, ntaps);
    CIAOSR_LAUNCH(rdn_pack_rows_kernel, cdiv(total, 256), 256, 0, st,
                  reinterpret_cast<__nv_bfloat16*>(p + off + L.rows_delta), w, Cin, ntaps);
    CIAOSR_LAUNCH(rdn_copy64_kernel, 1, 64, 0, st, bias + 64 * bi, b);
    ++bi;
    return CIAOSR_OK;
  };
  if ((rc = tc_pack_operand(p + L.sfe1_blob, 1, 64, 27, 0, Sfe1Src{d->sfe1_w}, st))) return rc;
  CIAOSR_LAUNCH(rdn_copy64_kernel, 1, 64, 0, st, bias + 64 * bi, d->sfe1_b);
  ++bi;
  if ((rc = pack(L.sfe2, d->sfe2_w, d->sfe2_b, 64, 9))) return rc;
  for (int r = 0; r < d->num_blocks; ++r)
    for (int l = 0; l < d->num_layers; ++l)
      if ((rc = pack(L.dense0 + r * L.dense_stride_block + L.dense_off[l], d->dense_w[r * d->num_layers + l],
                     d->dense_b[r * d->num_layers + l], 64 * (1 + l), 9))) return rc;
  for (int r = 0; r < d->num_blocks; ++r)
    if ((rc = pack(L.lff0 + (size_t)r * (1 + d->num_layers) * SLAB_BYTES, d->lff_w[r], d->lff_b[r],
                   64 * (1 + d->num_layers), 1))) return rc;
  if ((rc = pack(L.gff0, d->gff0_w, d->gff0_b, 64 * d->num_blocks, 1))) return rc;
  if ((rc = pack(L.gff1, d->gff1_w, d->gff1_b, 64, 9))) return rc;
  return CIAOSR_OK;
}

int ciaosr_rdn_workspace_bytes(const ciaosr_rdn_desc* d, int B, int H, int W, size_t* bytes) {
  CIAOSR_REQUIRE(bytes != nullptr, CIAOSR_E_INVALID, "bytes is NULL");
  int rc = rdn_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(B > 0 && H > 0 && W > 0, CIAOSR_E_INVALID, "bad shape");
  Arena a(nullptr, 0);
  rdn_carve(a, d, B, H, W);
  *bytes = a.used();
  return CIAOSR_OK;
}

int ciaosr_rdn_forward(const ciaosr_rdn_desc* d, const void* plan, const float* x, int B, int H, int W,
                       float* feature, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = rdn_check(d);
  if (rc) return rc;
  CIAOSR_REQUIRE(plan && x && feature && workspace, CIAOSR_E_INVALID, "NULL pointer argument");
  CIAOSR_REQUIRE(B > 0 && H > 0 && W > 0, CIAOSR_E_INVALID, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const RdnPlanLayout L = rdn_layout(d);
  Arena a(workspace, workspace_bytes);
  RdnWs w = rdn_carve(a, d, B, H, W);
  CIAOSR_REQUIRE(a.ok && ((uintptr_t)workspace % 256) == 0, CIAOSR_E_WORKSPACE,
                 "RDN workspace too small or misaligned: need %zu, have %zu", a.used(), workspace_bytes);
  const uint8_t* p = (const uint8_t*)plan;
  const float* bias = reinterpret_cast<const float*>(p + L.bias0);
  const int P = W + 2, HP2P = (H + 2) * P, nl = d->num_layers, nb = d->num_blocks, cbuf = 64 * (1 + nl);
  StageScope sc(5, st);

  // tensor maps of every source buffer (hi, lo)
  CUtensorMap m_f1[2], m_rb[2][2], m_gf[2], m_g1[2];
  if ((rc = make_map(&m_f1[0], w.f1h, w.Npa, 64)) || (rc = make_map(&m_f1[1], w.f1l, w.Npa, 64))) return rc;
  for (int i = 0; i < 2; ++i)
    if ((rc = make_map(&m_rb[i][0], w.rbh[i], w.Npa, cbuf)) || (rc = make_map(&m_rb[i][1], w.rbl[i], w.Npa, cbuf)))
      return rc;
  if ((rc = make_map(&m_gf[0], w.gfh, w.Npa, 64 * nb)) || (rc = make_map(&m_gf[1], w.gfl, w.Npa, 64 * nb))) return rc;
  if ((rc = make_map(&m_g1[0], w.g1h, w.Npa, 64)) || (rc = make_map(&m_g1[1], w.g1l, w.Npa, 64))) return rc;
  // halo boxes for the 3x3 layers (conv3_tc_kernel), when the image is narrow enough for one box
  const int HR = conv3_halo_rows(W + 2);
  CUtensorMap h_f1[2], h_rb[2][2], h_g1[2];
  if (HR) {
    if ((rc = make_map(&h_f1[0], w.f1h, w.Npa, 64, HR)) || (rc = make_map(&h_f1[1], w.f1l, w.Npa, 64, HR))) return rc;
    for (int i = 0; i < 2; ++i)
      if ((rc = make_map(&h_rb[i][0], w.rbh[i], w.Npa, cbuf, HR)) || (rc = make_map(&h_rb[i][1], w.rbl[i], w.Npa, cbuf, HR)))
        return rc;
    if ((rc = make_map(&h_g1[0], w.g1h, w.Npa, 64, HR)) || (rc = make_map(&h_g1[1], w.g1l, w.Npa, 64, HR))) return rc;
  }

  // sfe1: 3 -> 64 from the NCHW image (padding pixels of its outputs are zeroed first)
  CIAOSR_CUDA_OK(cudaMemsetAsync(w.f1h, 0, (size_t)w.Npa * 64 * 2, st));
  CIAOSR_CUDA_OK(cudaMemsetAsync(w.f1l, 0, (size_t)w.Npa * 64 * 2, st));
  CIAOSR_CUDA_OK(cudaMemsetAsync(w.f1_32, 0, (size_t)w.Npa * 64 * 4, st));
  if ((rc = tc_gemm(GemmShape{(long long)B * H * W, 1, 1, (long long)B * H * W, 0}, p + L.sfe1_blob,
                    Sfe1Gen{x, H, W}, Sfe1Epi{w.f1h, w.f1l, w.f1_32, bias, H, W, P, HP2P}, st))) return rc;

  // convw boxes: [0] halo (narrow images), [1] 136-row band, [2] 128 rows
  static const char* impl_env = getenv("CIAOSR_CONV_IMPL");
  const bool use_w = !(impl_env && impl_env[0] == '3');
  CUtensorMap w_f1[3][2], w_rb[2][3][2], w_gf[3][2], w_g1[3][2];
  if (use_w) {
    auto mk = [&](CUtensorMap (*M)[2], __nv_bfloat16* bh, __nv_bfloat16* bl, int ch) -> int {
      int r2;
      if (HR && ((r2 = make_map(&M[0][0], bh, w.Npa, ch, HR)) || (r2 = make_map(&M[0][1], bl, w.Npa, ch, HR)))) return r2;
      if ((r2 = make_map(&M[1][0], bh, w.Npa, ch, 136)) || (r2 = make_map(&M[1][1], bl, w.Npa, ch, 136))) return r2;
      if ((r2 = make_map(&M[2][0], bh, w.Npa, ch, 128)) || (r2 = make_map(&M[2][1], bl, w.Npa, ch, 128))) return r2;
      return 0;
    };
    if ((rc = mk(w_f1, w.f1h, w.f1l, 64)) || (rc = mk(w_rb[0], w.rbh[0], w.rbl[0], cbuf)) ||
        (rc = mk(w_rb[1], w.rbh[1], w.rbl[1], cbuf)) || (rc = mk(w_gf, w.gfh, w.gfl, 64 * nb)) ||
        (rc = mk(w_g1, w.g1h, w.g1l, 64))) return rc;
  }
  CUtensorMap (*wsel)[2] = nullptr;

  ConvParams c{};
  c.n_tiles = (int)(w.Npa / ROWS); c.P = P; c.HP2P = HP2P; c.H = H; c.W = W; c.Np = (int)w.Np;
  // bias rows in the plan: [sfe1, sfe2, dense (block-major), lff (per block), gff0, gff1]
  auto conv = [&](const CUtensorMap* src, const CUtensorMap* halo, int Cin, int ntaps, size_t blob_off, int bias_row,
                  int relu, const float* res32, ConvDst d1, ConvDst d2, float* out32, float* out_nchw) -> int {
    c.ntaps = ntaps; c.cblocks = Cin / 64; c.blob = p + blob_off; c.bias = bias + 64 * bias_row; c.res32 = res32;
    c.relu = relu; c.d1 = d1; c.d2 = d2; c.out32 = out32; c.out_nchw = out_nchw;
    if (use_w) return launch_convw(c, p + blob_off + L.rows_delta, P, wsel, st);
    if (ntaps == 9 && HR && halo != nullptr) return launch_conv3(c, HR, halo[0], halo[1], st);
    return launch_conv(c, src[0], src[1], st);
  };
  const int row_dense = 2, row_lff = 2 + nb * nl, row_gff = 2 + nb * nl + nb;
  const ConvDst none{nullptr, nullptr, 0, 0};
  // sfe2: F1 -> first 64 channels of RDB buffer 0 (+ fp32 trunk copy)
  wsel = w_f1;
  if ((rc = conv(m_f1, h_f1, 64, 9, L.sfe2, 1, 0, nullptr, ConvDst{w.rbh[0], w.rbl[0], cbuf, 0}, none, w.xr[0], nullptr)))
    return rc;
  for (int r = 0; r < nb; ++r) {
    const int cur = r & 1, nxt = cur ^ 1;
    wsel = w_rb[cur];
    for (int l = 0; l < nl; ++l)       // dense layer: conv3x3 + ReLU over channels [0, 64(1+l)) -> slice l+1
      if ((rc = conv(m_rb[cur], h_rb[cur], 64 * (1 + l), 9, L.dense0 + r * L.dense_stride_block + L.dense_off[l], row_dense + r * nl + l, 1, nullptr,
                     ConvDst{w.rbh[cur], w.rbl[cur], cbuf, 64 * (1 + l)}, none, nullptr, nullptr))) return rc;
    // local feature fusion 1x1 + residual (fp32 trunk) -> next block's input slice and the global fusion buffer
    if ((rc = conv(m_rb[cur], nullptr, cbuf, 1, L.lff0 + (size_t)r * (1 + nl) * SLAB_BYTES, row_lff + r, 0, w.xr[cur],
                   ConvDst{w.rbh[nxt], w.rbl[nxt], cbuf, 0}, ConvDst{w.gfh, w.gfl, 64 * nb, 64 * r}, w.xr[nxt],
                   nullptr))) return rc;
  }
  // global feature fusion: 1x1 over all block outputs, then 3x3, + sfe1 output -> feature (NCHW fp32)
  wsel = w_gf;
  if ((rc = conv(m_gf, nullptr, 64 * nb, 1, L.gff0, row_gff, 0, nullptr, ConvDst{w.g1h, w.g1l, 64, 0}, none, nullptr, nullptr)))
    return rc;
  wsel = w_g1;
  if ((rc = conv(m_g1, h_g1, 64, 9, L.gff1, row_gff + 1, 0, w.f1_32, none, none, nullptr, feature))) return rc;
  return CIAOSR_OK;
}

}  // extern "C"

#ifdef CIAOSR_TC_TIMING
extern "C" int ciaosr_debug_wait_read_rdn(unsigned long long* cycles, unsigned long long* counts, int reset) {
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
