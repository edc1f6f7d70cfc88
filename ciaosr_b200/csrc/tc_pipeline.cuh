// Shared pipeline of the tcgen05 kernels (head_tc.cu, gemm_tc.cuh): smem carve-up, mbarrier
// protocol, weight-unit producer, UMMA job issuer, accumulator drain helpers.
//
// One CTA per SM.  smem: 4 A-operand slots (bf16 hi + lo, [128 rows x 64 K] SW128 slabs), a 2-stage
// ring of 32 KB weight units ([128 N x 64 K] hi + lo), 16 KB of per-kernel constants, 16 mbarriers.
// TMEM: all 512 columns = two 128 x 256 fp32 accumulators D[0], D[1], used alternately by
// consecutive jobs.  A "job" = D[j&1][:, 0:128*units) = A[128, 64*nslabs] . W[128*units, 64*nslabs]^T
// evaluated as three bf16 UMMAs per product term (A_lo.W_hi + A_hi.W_lo + A_hi.W_hi).
//
// barrier      arrivals            producer -> consumer
//   W_FULL[2]   1 + tx bytes        weight producer (bulk copy) -> UMMA issuer
//   W_EMPTY[2]  1 (tcgen05.commit)  UMMA issuer -> weight producer
//   A_READY[4]  128 row threads     operand writers -> UMMA issuer      (per slab slot)
//   A_FREE[4]   1 (tcgen05.commit)  UMMA issuer -> operand writers      (only waited on when K > 256)
//   D_READY[2]  1 (tcgen05.commit)  UMMA issuer -> row threads          (accumulator complete)
//   D_FREE[2]   128 row threads     row threads -> UMMA issuer          (accumulator drained)
#pragma once
#include "tc_common.cuh"

namespace ciaosr {
using namespace tc;

// ---- static smem layout (bytes) --------------------------------------------------------------
constexpr int SM_A_HI = 0;                              // 4 slabs
constexpr int SM_A_LO = 4 * SLAB_BYTES;                 // 4 slabs
constexpr int SM_W = 8 * SLAB_BYTES;                    // 2 ring stages x (hi, lo)
constexpr int W_STAGES = 2;
constexpr int SM_CONST = SM_W + W_STAGES * UNIT_BYTES;  // floats: per-kernel constants
constexpr int CONST_FLOATS = 16 * HID;
constexpr int SM_BAR = SM_CONST + CONST_FLOATS * 4;
// barriers (8 B each): W_full[2] W_empty[2] A_ready[4] A_free[4] D_ready[2] D_free[2]; then tmem slot
constexpr int BAR_W_FULL = 0, BAR_W_EMPTY = 2, BAR_A_READY = 4, BAR_A_FREE = 8, BAR_D_READY = 12,
              BAR_D_FREE = 14, N_BARS = 16;
constexpr int SM_TOTAL = SM_BAR + N_BARS * 8 + 16;
constexpr int TC_THREADS = 256;
constexpr int EPI_T0 = 128;                             // first epilogue thread

struct TcShared {
  uint32_t a_hi, a_lo, w, bar, slot;
  float* consts;
};

__device__ __forceinline__ TcShared tc_carve(uint8_t* smem) {
  TcShared s;
  const uint32_t base = smem_u32(smem);
  s.a_hi = base + SM_A_HI; s.a_lo = base + SM_A_LO; s.w = base + SM_W;
  s.bar = base + SM_BAR; s.slot = base + SM_BAR + N_BARS * 8;
  s.consts = reinterpret_cast<float*>(smem + SM_CONST);
  return s;
}
__device__ __forceinline__ uint32_t bar_at(const TcShared& s, int i) { return s.bar + 8u * i; }

// common prologue: barrier init + TMEM allocation; returns the TMEM base address
__device__ __forceinline__ uint32_t tc_prologue(const TcShared& s, uint8_t* smem) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(bar_at(s, BAR_W_FULL + i), 1); mbar_init(bar_at(s, BAR_W_EMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_at(s, BAR_A_READY + i), 128); mbar_init(bar_at(s, BAR_A_FREE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_at(s, BAR_D_READY + i), 1); mbar_init(bar_at(s, BAR_D_FREE + i), 128); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(s.slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + N_BARS * 8);
}
__device__ __forceinline__ void tc_epilogue_dealloc(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- weight producer (warp 0; the whole warp walks the loop, lane 0 issues) -------------------------
struct ProdState { int stage; uint32_t phase; };

// stream `nunits` consecutive 32 KB weight units starting at `blob` through the ring
__device__ __forceinline__ void produce_units(const TcShared& s, ProdState& ps, const uint8_t* blob, int nunits) {
  const bool leader = (threadIdx.x & 31) == 0;
  for (int u = 0; u < nunits; ++u) {
    mbar_wait(bar_at(s, BAR_W_EMPTY + ps.stage), ps.phase ^ 1, 100);
    if (leader) {
      mbar_arrive_expect_tx(bar_at(s, BAR_W_FULL + ps.stage), UNIT_BYTES);
      const uint8_t* src = blob + (size_t)u * UNIT_BYTES;
      const uint32_t dst = s.w + ps.stage * UNIT_BYTES;
      bulk_g2s(dst, src, SLAB_BYTES, bar_at(s, BAR_W_FULL + ps.stage));
      bulk_g2s(dst + SLAB_BYTES, src + SLAB_BYTES, SLAB_BYTES, bar_at(s, BAR_W_FULL + ps.stage));
    }
    __syncwarp();
    if (++ps.stage == W_STAGES) { ps.stage = 0; ps.phase ^= 1; }
  }
}
__device__ __forceinline__ void producer_loop(const TcShared& s, const uint8_t* blob, int units_per_tile,
                                              int n_tiles) {
  ProdState ps{0, 0};
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) produce_units(s, ps, blob, units_per_tile);
}

// ---- UMMA issuer (warp 1; the whole warp walks the loop, lane 0 issues and commits) -------------------
struct MmaState { int stage; uint32_t wphase; uint32_t jobctr; uint32_t aready_bits; };

// one job: D[jobctr & 1][:, 0 : 128*units) = A (nslabs x 64 K) * W^T ; A slabs cycle through the 4 smem slots
__device__ __forceinline__ void mma_job(const TcShared& s, uint32_t tmem_base, MmaState& m, int nslabs,
                                        int units, bool a_new) {
  constexpr uint32_t IDESC = make_idesc_bf16(ROWS, UNIT_N);
  const bool leader = (threadIdx.x & 31) == 0;
  const uint32_t d = m.jobctr & 1, n = m.jobctr >> 1;
  mbar_wait(bar_at(s, BAR_D_FREE + d), (n + 1) & 1, 200);
  tc_fence_after();
  for (int sl = 0; sl < nslabs; ++sl) {
    const int slot = sl & 3;
    if (a_new) {
      mbar_wait(bar_at(s, BAR_A_READY + slot), (m.aready_bits >> slot) & 1, 210 + slot);
      m.aready_bits ^= 1u << slot;
      tc_fence_after();
    }
    const uint32_t a_hi = s.a_hi + slot * SLAB_BYTES, a_lo = s.a_lo + slot * SLAB_BYTES;
    for (int u = 0; u < units; ++u) {
      mbar_wait(bar_at(s, BAR_W_FULL + m.stage), m.wphase, 220);
      tc_fence_after();
      const uint32_t w_hi = s.w + m.stage * UNIT_BYTES, w_lo = w_hi + SLAB_BYTES;
      const uint32_t dcol = tmem_base + d * 256 + u * UNIT_N;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = make_desc_sw128(a_hi + ks * 32), al = make_desc_sw128(a_lo + ks * 32);
          const uint64_t bh = make_desc_sw128(w_hi + ks * 32), bl = make_desc_sw128(w_lo + ks * 32);
          umma_bf16(dcol, al, bh, IDESC, (sl | ks) != 0 ? 1u : 0u);   // small terms first
          umma_bf16(dcol, ah, bl, IDESC, 1u);
          umma_bf16(dcol, ah, bh, IDESC, 1u);
        }
        umma_commit(bar_at(s, BAR_W_EMPTY + m.stage));
      }
      __syncwarp();
      if (++m.stage == W_STAGES) { m.stage = 0; m.wphase ^= 1; }
    }
    if (leader) umma_commit(bar_at(s, BAR_A_FREE + slot));
    __syncwarp();
  }
  if (leader) umma_commit(bar_at(s, BAR_D_READY + d));
  __syncwarp();
  ++m.jobctr;
}

// ---- epilogue-side helpers (threads 128..255, one row each) -----------------------------------------
struct EpiState { uint32_t jobctr; uint32_t afree_bits; };   // afree_bits: parity to wait on next, per slot

__device__ __forceinline__ void slab_begin(const TcShared& s, EpiState& e, int slot, bool wait_free) {
  if (wait_free) {
    mbar_wait(bar_at(s, BAR_A_FREE + slot), (e.afree_bits >> slot) & 1, 300 + slot);
  }
  e.afree_bits ^= 1u << slot;
}
__device__ __forceinline__ void slab_done(const TcShared& s, int slot) {
  fence_proxy_async();
  mbar_arrive(bar_at(s, BAR_A_READY + slot));
}
__device__ __forceinline__ uint32_t epi_wait_d(const TcShared& s, EpiState& e) {
  const uint32_t d = e.jobctr & 1, n = e.jobctr >> 1;
  mbar_wait(bar_at(s, BAR_D_READY + d), n & 1, 310);
  tc_fence_after();
  return d;
}
__device__ __forceinline__ void epi_release_d(const TcShared& s, EpiState& e) {
  tc_fence_before();
  mbar_arrive(bar_at(s, BAR_D_FREE + (e.jobctr & 1)));
  ++e.jobctr;
}

// hidden layer epilogue: next A = relu(D + bias), written slab by slab
template <bool WAIT_FREE>
__device__ __forceinline__ void epi_hidden(const TcShared& s, EpiState& e, uint32_t lane_taddr, int row,
                                           const float* __restrict__ bias_s) {
  const uint32_t d = epi_wait_d(s, e);
#pragma unroll 1
  for (int sl = 0; sl < 4; ++sl) {
    slab_begin(s, e, sl, WAIT_FREE);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      tmem_ld32(lane_taddr + d * 256 + sl * 64 + half * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bias_s[sl * 64 + half * 32 + i], 0.0f);
      a_store32(s.a_hi + sl * SLAB_BYTES, s.a_lo + sl * SLAB_BYTES, row, half * 32, v);
    }
    slab_done(s, sl);
  }
  epi_release_d(s, e);
}


}  // namespace ciaosr
