// RDN encoder on the tensor cores (SURVEY.md section 8f "next" #2: the encoder fast path).
//
// The reference keeps the encoder in PyTorch (mmedit's RDN, hoisted at ciaosr_net.py:314-318; forward
// ciaosr_net.py:321-342).  cuDNN offers two ways to run its 146 convolutions: TF32 (10 ms for the bench
// batch, but ~1e-3 relative error on the features, which the head turns into ~1e-3 output error: 10x the
// parity tolerance) or fp32 CUDA cores (47 ms).  This file runs them as implicit GEMMs on tcgen05 with
// fp16 hi/lo operand splits (fp32-grade results at tensor-core speed).
//
// Layout: activations live in HBM as plain NHWC, two 16-bit tensors (hi, lo halves) [B, H, W, C]; nothing is padded.
// One work tile = 16 rows x 8 columns of one image = 128 output pixels x 64 output channels.
//
// "Weights in tensor memory" formulation (convw).  Measured on B200 (profiles/r01e): a tcgen05.mma
// M128 x N x K16 costs ~41 + 0.58 N cycles whatever its operand sources, so the 64 output channels make a poor
// N.  The roles are therefore swapped and the hi/lo halves of the weights stacked along M:
//   A (M = 128) = the weights of one (64-channel block, tap) K-slab, rows [W_hi (64 out channels); W_lo (64)],
//                 staged in TENSOR MEMORY by eight "stager" warps (ld.global -> tcgen05.st; the weights never
//                 touch shared memory), 8 slabs deep;
//   B (N = 128) = the 128 pixels of the tile, from shared memory.  ONE 4-D TMA box [18 rows x 10 columns x 64
//                 channels] per (tile, channel block, hi|lo) serves all nine taps: out-of-image coordinates are
//                 zero-filled by TMA (that is the convolution's padding), and tap (dy, dx) is the same smem
//                 tile read through a UMMA descriptor whose start address is advanced by (dy+1)*10 + (dx+1)
//                 box rows and whose 8-row-group stride is one box row pitch (10 x 128 B).  The 128B swizzle
//                 is a function of the absolute smem address for the TMA write and the UMMA read alike, so a
//                 start address that is not 1024-byte aligned needs nothing else (descriptor base offset 0;
//                 measured: the other reading of the PTX text, base offset = row & 7, gives wrong sums);
//   D[lane, col] = two accumulators of 128 columns: lanes 0-63 hold W_hi.X, lanes 64-127 hold W_lo.X with
//                 X = X_hi + X_lo (all four hi/lo products, 2 instructions per K16).
// The epilogue transposes D through shared memory (out[pixel][c] = D[c][pixel] + D[64 + c][pixel] + bias),
// applies residual / ReLU, splits to fp16 hi/lo and writes the next layer's buffers with 128-byte rows.
// Dense blocks never concatenate: every RDB owns one 576-channel buffer and each layer writes its 64
// channels into its slice; the LFF output goes straight into the next block's buffer and the global
// fusion buffer.  The residual trunk is kept in fp32.
#include "gemm_tc.cuh"
#include "tma.cuh"
#include "kernels.cuh"

namespace ciaosr {

constexpr int TILE_X = 8, TILE_Y = 16;    // pixels per tile: one 8-row group of the B operand per image row
constexpr int CW_THREADS = 512;           // warps: 0 TMA, 1 UMMA issuer, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 stagers
constexpr int CW_W_STAGES = 8;            // weight slabs resident in TMEM: 8 x 32 columns after the accumulators
constexpr int CW_A_STAGES = 3;            // activation boxes in flight (per stage: hi half + lo half)
constexpr int CW_T_LD = 132;              // floats per pixel row of the transpose buffer (padded)
constexpr int CW_A_READY = 0, CW_A_FREE = 4, CW_W_FULL = 8, CW_W_EMPTY = 16, CW_D_READY = 24, CW_D_FREE = 26,
              CW_NBARS = 28;
constexpr int CW_FIXED_BYTES = 32 * CW_T_LD * 4 + 64 * 4 + CW_NBARS * 8 + 16;

struct ConvDst { split_t* hi; split_t* lo; int ld; int choff; };
#ifdef CIAOSR_CONV_STATS
// diagnostic build (tools/conv_issuer_stats.py): the issuer's clock64 deltas around its three mbarrier waits, accumulated in
// registers and added up once per CTA -- [0] wait W_FULL, [1] wait A_READY, [2] wait D_FREE, [3] total, [4] taps, [5] tiles, [6] CTAs
__device__ unsigned long long g_conv_stats[8];
#define CONV_STAT_BEGIN() const long long stat_c0 = clock64()
#define CONV_STAT_END(acc) acc += clock64() - stat_c0
#else
#define CONV_STAT_BEGIN() do {} while (0)
#define CONV_STAT_END(acc) do {} while (0)
#endif

struct ConvParams {
  int B, H, W, tiles_x, tiles_y, n_tiles;
  int ntaps, cblocks;                  // 9 or 1; Cin / 64
  int half_bytes;                      // smem bytes of one (hi or lo) stage half, multiple of 1024
  int box_x;                           // columns of the TMA box (10 for 3x3, 8 for 1x1) = 8-row-group stride / 128
  const uint8_t* wrows;                // per K-slab (cblock, tap): 128 rows [W_hi; W_lo] x 64 halves as [chunk of 8][row][8]
  const float* bias;                   // [64]
  const float* res32;                  // fp32 [B*H*W, 64] residual or nullptr
  int relu;
  ConvDst d1, d2;                      // d2.hi == nullptr if unused
  float* out32;                        // fp32 [B*H*W, 64] copy (residual trunk) or nullptr
  float* out_nchw;                     // final feature [B,64,H,W] or nullptr
};

// D[tmem] (+)= A[tmem] * B[smem]^T, B descriptor high word given explicitly (8-row-group stride varies)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t b_hi32,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(b_hi32)
      : "memory");
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
convw_tc_kernel(const ConvParams P, const __grid_constant__ CUtensorMap map_hi,
                const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int fixed0 = 2 * CW_A_STAGES * P.half_bytes;         // [stage][hi|lo] halves first
  float* tbuf = reinterpret_cast<float*>(smem + fixed0);     // [32 pixels][CW_T_LD]
  float* bias_s = tbuf + 32 * CW_T_LD;
  const uint32_t bars = sbase + fixed0 + 32 * CW_T_LD * 4 + 64 * 4;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + fixed0 + 32 * CW_T_LD * 4 + 64 * 4 + CW_NBARS * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) bias_s[threadIdx.x] = P.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < CW_A_STAGES; ++i) { mbar_init(bars + 8 * (CW_A_READY + i), 1); mbar_init(bars + 8 * (CW_A_FREE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bars + 8 * (CW_D_READY + i), 1); mbar_init(bars + 8 * (CW_D_FREE + i), 4); }
    for (int i = 0; i < CW_W_STAGES; ++i) { mbar_init(bars + 8 * (CW_W_FULL + i), 8); mbar_init(bars + 8 * (CW_W_EMPTY + i), 1); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(slot);
  // Programmatic dependent launch: the next layer's CTAs may take over this SM as soon as this CTA exits (their
  // prologue and weight prefetch overlap this layer's tail); everything that touches the previous layer's output
  // -- the TMA producer and the epilogue (residual reads, all writes) -- first waits for that grid to complete.
  // The weight stagers read constants only and start right away.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp < 8) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int units = P.cblocks * P.ntaps;                     // K-slabs per tile
  const int tiles_per_image = P.tiles_x * P.tiles_y;
  const int halo = P.ntaps == 9 ? 1 : 0;

  if (warp == 0) {
    // ---- TMA producer: the activation box of one channel block (hi and lo) per stage ----
    uint32_t acnt = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_image, tt = tile - b * tiles_per_image;
      const int y0 = (tt / P.tiles_x) * TILE_Y - halo, x0 = (tt % P.tiles_x) * TILE_X - halo;
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt % CW_A_STAGES;
        mbar_wait(bars + 8 * (CW_A_FREE + ast), ((acnt / CW_A_STAGES) & 1) ^ 1, 400 + ast);
        if (lane == 0) {
          const uint32_t full = bars + 8 * (CW_A_READY + ast);
          const uint32_t box_bytes = (uint32_t)((TILE_Y + 2 * halo) * P.box_x * 128);
          mbar_arrive_expect_tx(full, 2u * box_bytes);
          const uint32_t dst = sbase + (uint32_t)(2 * ast) * P.half_bytes;
          tma_load_4d(dst, &map_hi, cb * 64, x0, y0, b, full);
          tma_load_4d(dst + P.half_bytes, &map_lo, cb * 64, x0, y0, b, full);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer ----
    const uint32_t idesc = make_idesc_split(ROWS, 128);
    const uint32_t bhw = (uint32_t)((P.box_x * 128) >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
    uint32_t acnt = 0, wcnt = 0, job = 0;
#ifdef CIAOSR_CONV_STATS
    long long tw = 0, ta = 0, td = 0;
    const long long tstart = clock64();
#endif
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      { CONV_STAT_BEGIN();
      mbar_wait(bars + 8 * (CW_D_FREE + d), (n + 1) & 1, 420);
      CONV_STAT_END(td); }
      tc_fence_after();
      const uint32_t dcol = tmem_base + d * 128;
      for (int cb = 0; cb < P.cblocks; ++cb, ++acnt) {
        const int ast = acnt % CW_A_STAGES;
        { CONV_STAT_BEGIN();
        mbar_wait(bars + 8 * (CW_A_READY + ast), (acnt / CW_A_STAGES) & 1, 430 + ast);
        CONV_STAT_END(ta); }
        const uint32_t x_hi = sbase + (uint32_t)(2 * ast) * P.half_bytes, x_lo = x_hi + P.half_bytes;
        for (int tap = 0; tap < P.ntaps; ++tap, ++wcnt) {
          const int wst = wcnt % CW_W_STAGES;
          { CONV_STAT_BEGIN();
          mbar_wait(bars + 8 * (CW_W_FULL + wst), (wcnt / CW_W_STAGES) & 1, 440 + wst);
          CONV_STAT_END(tw); }
          tc_fence_after();
          if (lane == 0) {
            const uint32_t toff = (uint32_t)((tap / 3) * P.box_x + tap % 3) * 128u;
            const uint32_t b_hi = desc_lo(x_hi + toff), b_lo = desc_lo(x_lo + toff);
            const uint32_t a = tmem_base + 256 + wst * 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_ts(dcol, a + 8 * ks, b_lo + 2 * ks, bhw, idesc, (cb | tap | ks) != 0 ? 1u : 0u);
              umma_ts(dcol, a + 8 * ks, b_hi + 2 * ks, bhw, idesc, 1u);
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
#ifdef CIAOSR_CONV_STATS
    if (lane == 0) {
      atomicAdd(&g_conv_stats[0], (unsigned long long)tw); atomicAdd(&g_conv_stats[1], (unsigned long long)ta);
      atomicAdd(&g_conv_stats[2], (unsigned long long)td); atomicAdd(&g_conv_stats[3], (unsigned long long)(clock64() - tstart));
      atomicAdd(&g_conv_stats[4], (unsigned long long)wcnt); atomicAdd(&g_conv_stats[5], (unsigned long long)job);
      atomicAdd(&g_conv_stats[6], 1ull);
    }
#endif
  } else if (warp >= 8) {
    // ---- weight stagers: row r of every K-slab, global -> registers -> tensor memory ----
    // warps 8-11 stage K columns 0-31 of every slab (TMEM columns 0-15 of its stage), warps 12-15 the rest;
    // four register buffers per thread: a load is issued three K-slabs ahead of its use
    const int r = (warp & 3) * 32 + lane, khalf = (warp - 8) >> 2;
    const uint32_t taddr0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256 + khalf * 16;
    int my_tiles = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) ++my_tiles;
    const long long total = (long long)my_tiles * units;
    const uint4* src0 = reinterpret_cast<const uint4*>(P.wrows) + khalf * 512 + r;    // + unit * 1024 + j * 128
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
  } else if (warp >= 4) {
    // ---- epilogue: transpose through shared memory, one 32-pixel chunk (4 image rows of the tile) at a time ----
    const int t = threadIdx.x - EPI_T0;                      // 0..127 = TMEM lane = stacked weight row
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int pp = t >> 2, c0 = (t & 3) * 16;                // pixel within the chunk, first of 16 channels
    uint32_t job = 0;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x, ++job) {
      const uint32_t d = job & 1, n = job >> 1;
      const int b = tile / tiles_per_image, tt = tile - b * tiles_per_image;
      const int y0 = (tt / P.tiles_x) * TILE_Y, x0 = (tt % P.tiles_x) * TILE_X;
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
        for (int j = 0; j < 32; ++j) tbuf[j * CW_T_LD + t] = v[j];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int yy = y0 + ch * 4 + (pp >> 3), xx = x0 + (pp & 7);
        const bool valid = yy < P.H && xx < P.W;
        if (!valid) continue;                                // (barriers above are reached by every thread)
        const long long g = ((long long)b * P.H + yy) * P.W + xx;
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
        if (P.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = fmaxf(o[i], 0.0f);
        }
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
        if (P.out_nchw != nullptr) {
          float* po = P.out_nchw + (((long long)b * 64 + c0) * P.H + yy) * P.W + xx;
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

// ---- weight packing: conv weight [64, Cin, kh, kw] -> per K-slab (cblock, tap) 128 rows x 64 halves -------------
// chunk-major: [K-slab][16-byte chunk j = k / 8][row][k % 8], rows = [W_hi (64); W_lo (64)], so that the 32 rows
// a stager warp loads with one instruction are 512 contiguous bytes
constexpr int WSLAB_BYTES = 128 * 64 * 2;
__global__ void rdn_pack_rows_kernel(split_t* __restrict__ dst, const float* __restrict__ w, int Cin, int ntaps) {
  const int cblocks = Cin / 64;
  const long long total = (long long)cblocks * ntaps * 64 * 64;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % 64), n = (int)((i / 64) % 64);
  const int sl = (int)(i / 4096);
  const int cb = sl / ntaps, tap = sl % ntaps;
  const float v = w[((long long)n * Cin + cb * 64 + k) * ntaps + tap];
  split_t hi, lo;
  split_scalar(v, hi, lo);
  split_t* ub = dst + (size_t)sl * 128 * 64;
  ub[((k >> 3) * 128 + n) * 8 + (k & 7)] = hi;
  ub[((k >> 3) * 128 + 64 + n) * 8 + (k & 7)] = lo;
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
struct Sfe1Epi {          // + bias -> NHWC fp16 hi/lo + fp32
  split_t* hi; split_t* lo; float* out32; const float* bias;
  __device__ __forceinline__ void store(const Sfe1Gen::Row&, long long g, int n0, const float (&v)[32]) const {
    if (n0 >= 64) return;
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
struct RdnPlanLayout {
  // byte offsets of the per-layer weight blobs and fp32 biases inside the plan buffer
  size_t sfe1_blob, sfe2, gff0, gff1, dense0, lff0;    // dense/lff: consecutive per block
  size_t bias0;                                         // floats: [sfe1, sfe2, dense..., lff..., gff0, gff1] x 64
  size_t dense_stride_block, total;
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
  L.sfe2 = take((size_t)9 * WSLAB_BYTES);
  size_t blk = 0;
  for (int l = 0; l < d->num_layers; ++l) { L.dense_off[l] = blk; blk += (size_t)9 * (1 + l) * WSLAB_BYTES; }
  L.dense_stride_block = blk;
  L.dense0 = take(blk * d->num_blocks);
  L.lff0 = take((size_t)(1 + d->num_layers) * WSLAB_BYTES * d->num_blocks);
  L.gff0 = take((size_t)d->num_blocks * WSLAB_BYTES);
  L.gff1 = take((size_t)9 * WSLAB_BYTES);
  const int nconv = 2 + d->num_blocks * (d->num_layers + 1) + 2;
  L.bias0 = take((size_t)nconv * 64 * sizeof(float));
  L.total = off;
  return L;
}

__global__ void rdn_copy64_kernel(float* dst, const float* src) { dst[threadIdx.x] = src[threadIdx.x]; }

// 4-D map over a 16-bit NHWC tensor [B, H, W, channels]: box = 64 channels x box_x x box_y x 1 image,
// 128B swizzle, out-of-bounds elements read as zero (the convolution's padding)
static int make_map(CUtensorMap* m, void* base, int B, int H, int W, int channels, int box_x, int box_y) {
  EncodeTiledFn enc = tma_get_encode();
  CIAOSR_REQUIRE(enc != nullptr, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)channels * 2, (cuuint64_t)W * channels * 2,
                                 (cuuint64_t)H * W * channels * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)box_x, (cuuint32_t)box_y, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, TMA_SPLIT_DTYPE, 4, base, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CIAOSR_REQUIRE(r == CUDA_SUCCESS, CIAOSR_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CIAOSR_OK;
}

struct RdnWs {
  long long M;         // B*H*W pixels
  split_t *f1h, *f1l, *rbh[2], *rbl[2], *gfh, *gfl, *g1h, *g1l;
  float *f1_32, *xr[2];
};
static RdnWs rdn_carve(Arena& a, const ciaosr_rdn_desc* d, int B, int H, int W) {
  RdnWs w;
  w.M = (long long)B * H * W;
  const size_t n = (size_t)w.M;
  const int cb = 64 * (1 + d->num_layers);
  w.f1h = a.take<split_t>(n * 64); w.f1l = a.take<split_t>(n * 64);
  for (int i = 0; i < 2; ++i) { w.rbh[i] = a.take<split_t>(n * cb); w.rbl[i] = a.take<split_t>(n * cb); }
  w.gfh = a.take<split_t>(n * 64 * d->num_blocks); w.gfl = a.take<split_t>(n * 64 * d->num_blocks);
  w.g1h = a.take<split_t>(n * 64); w.g1l = a.take<split_t>(n * 64);
  w.f1_32 = a.take<float>(n * 64);
  for (int i = 0; i < 2; ++i) w.xr[i] = a.take<float>(n * 64);
  return w;
}

struct MapPair { CUtensorMap m[2][2]; };     // [0] 3x3 box (10 x 18), [1] 1x1 box (8 x 16); [hi, lo]
static int make_maps(MapPair* mp, split_t* hi, split_t* lo, int B, int H, int W, int channels) {
  int rc;
  if ((rc = make_map(&mp->m[0][0], hi, B, H, W, channels, TILE_X + 2, TILE_Y + 2)) ||
      (rc = make_map(&mp->m[0][1], lo, B, H, W, channels, TILE_X + 2, TILE_Y + 2)) ||
      (rc = make_map(&mp->m[1][0], hi, B, H, W, channels, TILE_X, TILE_Y)) ||
      (rc = make_map(&mp->m[1][1], lo, B, H, W, channels, TILE_X, TILE_Y))) return rc;
  return CIAOSR_OK;
}

static int launch_convw(ConvParams P, const MapPair& mp, cudaStream_t st) {
  const int k = P.ntaps == 9 ? 0 : 1;
  P.box_x = P.ntaps == 9 ? TILE_X + 2 : TILE_X;
  const int box_y = P.ntaps == 9 ? TILE_Y + 2 : TILE_Y;
  P.half_bytes = (P.box_x * box_y * 128 + 1023) / 1024 * 1024;
  const int smem_bytes = 2 * CW_A_STAGES * P.half_bytes + CW_FIXED_BYTES;
  static DynSmemOptIn optin;          // per device (common.cuh)
  if (int rc = optin.ensure(convw_tc_kernel, smem_bytes)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tc_grid_size(P.n_tiles)); cfg.blockDim = dim3(CW_THREADS);
  cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, convw_tc_kernel, P, mp.m[k][0], mp.m[k][1]);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CIAOSR_REQUIRE(e == cudaSuccess, CIAOSR_E_CUDA, "launch of convw_tc_kernel failed: %s", cudaGetErrorString(e));
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
    CIAOSR_LAUNCH(rdn_pack_rows_kernel, cdiv(total, 256), 256, 0, st, reinterpret_cast<split_t*>(p + off), w,
                  Cin, ntaps);
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
    if ((rc = pack(L.lff0 + (size_t)r * (1 + d->num_layers) * WSLAB_BYTES, d->lff_w[r], d->lff_b[r],
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
  const int nl = d->num_layers, nb = d->num_blocks, cbuf = 64 * (1 + nl);
  StageScope sc(5, st);

  // tensor maps of every source buffer
  MapPair m_f1, m_rb[2], m_gf, m_g1;
  if ((rc = make_maps(&m_f1, w.f1h, w.f1l, B, H, W, 64)) || (rc = make_maps(&m_rb[0], w.rbh[0], w.rbl[0], B, H, W, cbuf)) ||
      (rc = make_maps(&m_rb[1], w.rbh[1], w.rbl[1], B, H, W, cbuf)) ||
      (rc = make_maps(&m_gf, w.gfh, w.gfl, B, H, W, 64 * nb)) || (rc = make_maps(&m_g1, w.g1h, w.g1l, B, H, W, 64)))
    return rc;

  // sfe1: 3 -> 64 from the NCHW image
  if ((rc = tc_gemm(GemmShape{w.M, 1, 1, w.M, 0}, p + L.sfe1_blob, Sfe1Gen{x, H, W},
                    Sfe1Epi{w.f1h, w.f1l, w.f1_32, bias}, st))) return rc;

  ConvParams c{};
  c.B = B; c.H = H; c.W = W;
  c.tiles_x = (W + TILE_X - 1) / TILE_X; c.tiles_y = (H + TILE_Y - 1) / TILE_Y;
  c.n_tiles = B * c.tiles_x * c.tiles_y;
  // bias rows in the plan: [sfe1, sfe2, dense (block-major), lff (per block), gff0, gff1]
  auto conv = [&](const MapPair& src, int Cin, int ntaps, size_t blob_off, int bias_row, int relu,
                  const float* res32, ConvDst d1, ConvDst d2, float* out32, float* out_nchw) -> int {
    c.ntaps = ntaps; c.cblocks = Cin / 64; c.wrows = p + blob_off; c.bias = bias + 64 * bias_row; c.res32 = res32;
    c.relu = relu; c.d1 = d1; c.d2 = d2; c.out32 = out32; c.out_nchw = out_nchw;
    return launch_convw(c, src, st);
  };
  const int row_dense = 2, row_lff = 2 + nb * nl, row_gff = 2 + nb * nl + nb;
  const ConvDst none{nullptr, nullptr, 0, 0};
  // sfe2: F1 -> first 64 channels of RDB buffer 0 (+ fp32 trunk copy)
  if ((rc = conv(m_f1, 64, 9, L.sfe2, 1, 0, nullptr, ConvDst{w.rbh[0], w.rbl[0], cbuf, 0}, none, w.xr[0], nullptr)))
    return rc;
  for (int r = 0; r < nb; ++r) {
    const int cur = r & 1, nxt = cur ^ 1;
    for (int l = 0; l < nl; ++l)       // dense layer: conv3x3 + ReLU over channels [0, 64(1+l)) -> slice l+1
      if ((rc = conv(m_rb[cur], 64 * (1 + l), 9, L.dense0 + r * L.dense_stride_block + L.dense_off[l],
                     row_dense + r * nl + l, 1, nullptr, ConvDst{w.rbh[cur], w.rbl[cur], cbuf, 64 * (1 + l)}, none,
                     nullptr, nullptr))) return rc;
    // local feature fusion 1x1 + residual (fp32 trunk) -> next block's input slice and the global fusion buffer
    if ((rc = conv(m_rb[cur], cbuf, 1, L.lff0 + (size_t)r * (1 + nl) * WSLAB_BYTES, row_lff + r, 0, w.xr[cur],
                   ConvDst{w.rbh[nxt], w.rbl[nxt], cbuf, 0}, ConvDst{w.gfh, w.gfl, 64 * nb, 64 * r}, w.xr[nxt],
                   nullptr))) return rc;
  }
  // global feature fusion: 1x1 over all block outputs, then 3x3, + sfe1 output -> feature (NCHW fp32)
  if ((rc = conv(m_gf, 64 * nb, 1, L.gff0, row_gff, 0, nullptr, ConvDst{w.g1h, w.g1l, 64, 0}, none, nullptr, nullptr)))
    return rc;
  if ((rc = conv(m_g1, 64, 9, L.gff1, row_gff + 1, 0, w.f1_32, none, none, nullptr, feature))) return rc;
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

#ifdef CIAOSR_CONV_STATS
extern "C" int ciaosr_debug_conv_stats(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, ciaosr::g_conv_stats, 8 * 8);
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(ciaosr::g_conv_stats, z, 8 * 8); }
  return 0;
}
#endif
