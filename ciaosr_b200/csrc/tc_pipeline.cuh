// Shared pipeline of the tcgen05 kernels (head_tc.cu, gemm_tc.cuh): smem carve-up, mbarrier
// protocol, weight producer, UMMA job issuer, accumulator drain helpers.
//
// One CTA per SM.  smem: 4 A-operand slots (bf16 hi + lo, [128 rows x 64 K] SW128 slabs), a 4-stage
// ring of 16 KB weight slabs ([128 N x 64 K]; a weight "unit" = its hi slab then its lo slab),
// 24 KB of per-kernel constants, 20 mbarriers.  TMEM: all 512 columns = two 128 x 256 fp32
// accumulators D[0], D[1], used alternately by consecutive jobs.
// A "job" = D[j&1][:, 0:128*units) = A[128, 64*nslabs] . W[128*units, 64*nslabs]^T evaluated as three
// bf16 UMMAs per product term (A_lo.W_hi + A_hi.W_hi on the hi slab, A_hi.W_lo on the lo slab).
//
// Template parameters:  CL   = CTAs per cluster sharing one weight stream (1 or 2).  With CL = 2 both
//                              CTAs walk the same job sequence on different row tiles; CTA 0 loads every
//                              hi slab, CTA 1 every lo slab, each copy is `.multicast::cluster` to both,
//                              and a stage is recycled only after BOTH issuers committed it
//                              (tcgen05.commit multicast) -> L2 weight reads per SM are halved.
//                       NEPI = row threads (128: one per row; 256: two per row, 32 columns of each
//                              64-column slab each).
//
// barrier      arrivals            producer -> consumer
//   W_FULL[4]   1 + 16 KB tx        weight producer (bulk copy) -> UMMA issuer
//   W_EMPTY[4]  CL (tcgen05.commit) UMMA issuer(s) -> weight producer
//   A_READY[4]  NEPI row threads    operand writers -> UMMA issuer      (per slab slot)
//   A_FREE[4]   1 (tcgen05.commit)  UMMA issuer -> operand writers      (only waited on when K > 256)
//   D_READY[2]  1 (tcgen05.commit)  UMMA issuer -> row threads          (accumulator complete)
//   D_FREE[2]   NEPI row threads    row threads -> UMMA issuer          (accumulator drained)
#pragma once
#include "tc_common.cuh"

namespace ciaosr {
using namespace tc;

// ---- static smem layout (bytes) --------------------------------------------------------------
constexpr int SM_A_HI = 0;                              // 4 slabs
constexpr int SM_A_LO = 4 * SLAB_BYTES;                 // 4 slabs
constexpr int SM_W = 8 * SLAB_BYTES;                    // ring stages
constexpr int W_STAGES = 4;
constexpr int SM_CONST = SM_W + W_STAGES * SLAB_BYTES;  // floats: per-kernel constants
constexpr int CONST_FLOATS = 16 * HID + 2048;
constexpr int SM_BAR = SM_CONST + CONST_FLOATS * 4;
constexpr int BAR_W_FULL = 0, BAR_W_EMPTY = 4, BAR_A_READY = 8, BAR_A_FREE = 12, BAR_D_READY = 16,
              BAR_D_FREE = 18, N_BARS = 20;
constexpr int SM_SLOT = SM_BAR + N_BARS * 8;            // TMEM base address (4 B, padded to 16)
constexpr int SM_XCHG = SM_SLOT + 16;                   // 1024 floats of row-thread exchange space
constexpr int SM_TOTAL = SM_XCHG + 1024 * 4;
constexpr int EPI_T0 = 128;                             // first row thread (warps 0..3 are control warps)

struct TcShared {
  uint32_t a_hi, a_lo, w, bar;
  float* consts;
  float* xchg;
};

__device__ __forceinline__ TcShared tc_carve(uint8_t* smem) {
  TcShared s;
  const uint32_t base = smem_u32(smem);
  s.a_hi = base + SM_A_HI; s.a_lo = base + SM_A_LO; s.w = base + SM_W;
  s.bar = base + SM_BAR;
  s.consts = reinterpret_cast<float*>(smem + SM_CONST);
  s.xchg = reinterpret_cast<float*>(smem + SM_XCHG);
  return s;
}
__device__ __forceinline__ uint32_t bar_at(const TcShared& s, int i) { return s.bar + 8u * i; }

// common prologue: barrier init + TMEM allocation; returns the TMEM base address
template <int CL, int NEPI>
__device__ __forceinline__ uint32_t tc_prologue(const TcShared& s, uint8_t* smem) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(bar_at(s, BAR_W_FULL + i), 1); mbar_init(bar_at(s, BAR_W_EMPTY + i), CL); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_at(s, BAR_A_READY + i), NEPI); mbar_init(bar_at(s, BAR_A_FREE + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_at(s, BAR_D_READY + i), 1); mbar_init(bar_at(s, BAR_D_FREE + i), NEPI); }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(smem) + SM_SLOT, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peer barriers are initialised before any multicast reaches them
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem + SM_SLOT);
}
template <int CL>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA exits while a peer may still signal its barriers
  if ((threadIdx.x >> 5) == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- weight producer (warp 0; the whole warp walks the loop, lane 0 issues) -------------------------
struct ProdState { int stage; uint32_t phase; uint32_t count; };

// stream `nunits` consecutive weight units (hi slab, lo slab: 2 x 16 KB each) starting at `blob`
template <int CL>
__device__ __forceinline__ void produce_units(const TcShared& s, ProdState& ps, const uint8_t* blob, int nunits,
                                              uint32_t cta_rank) {
  const bool leader = (threadIdx.x & 31) == 0;
  for (int i = 0; i < 2 * nunits; ++i) {
    mbar_wait(bar_at(s, BAR_W_EMPTY + ps.stage), ps.phase ^ 1, 100);
    if (leader) {
      const uint32_t full = bar_at(s, BAR_W_FULL + ps.stage);
      mbar_arrive_expect_tx(full, SLAB_BYTES);
      const uint8_t* src = blob + (size_t)i * SLAB_BYTES;
      const uint32_t dst = s.w + ps.stage * SLAB_BYTES;
      if (CL == 1) bulk_g2s(dst, src, SLAB_BYTES, full);
      else if ((ps.count & 1u) == cta_rank) bulk_g2s_mc(dst, src, SLAB_BYTES, full, (uint16_t)((1u << CL) - 1));
    }
    __syncwarp();
    ++ps.count;
    if (++ps.stage == W_STAGES) { ps.stage = 0; ps.phase ^= 1; }
  }
}

// ---- UMMA issuer (warp 1; the whole warp walks the loop, lane 0 issues and commits) -------------------
struct MmaState { int stage; uint32_t wphase; uint32_t jobctr; uint32_t aready_bits; };

template <int CL>
__device__ __forceinline__ void release_stage(const TcShared& s, MmaState& m, bool leader) {
  if (leader) {
    if (CL == 1) umma_commit(bar_at(s, BAR_W_EMPTY + m.stage));
    else umma_commit_mc(bar_at(s, BAR_W_EMPTY + m.stage), (uint16_t)((1u << CL) - 1));
  }
  __syncwarp();
  if (++m.stage == W_STAGES) { m.stage = 0; m.wphase ^= 1; }
}

// one job: D[jobctr & 1][:, 0 : 128*units) = A (nslabs x 64 K) * W^T ; A slabs cycle through the 4 smem slots
template <int CL>
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
      const uint32_t dcol = tmem_base + d * 256 + u * UNIT_N;
      // hi slab of the unit: A_lo.W_hi (small term first) and A_hi.W_hi
      mbar_wait(bar_at(s, BAR_W_FULL + m.stage), m.wphase, 220);
      tc_fence_after();
      if (leader) {
        const uint32_t w_hi = s.w + m.stage * SLAB_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bh = make_desc_sw128(w_hi + ks * 32);
          umma_bf16(dcol, make_desc_sw128(a_lo + ks * 32), bh, IDESC, (sl | ks) != 0 ? 1u : 0u);
          umma_bf16(dcol, make_desc_sw128(a_hi + ks * 32), bh, IDESC, 1u);
        }
      }
      release_stage<CL>(s, m, leader);
      // lo slab: A_hi.W_lo
      mbar_wait(bar_at(s, BAR_W_FULL + m.stage), m.wphase, 221);
      tc_fence_after();
      if (leader) {
        const uint32_t w_lo = s.w + m.stage * SLAB_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_bf16(dcol, make_desc_sw128(a_hi + ks * 32), make_desc_sw128(w_lo + ks * 32), IDESC, 1u);
      }
      release_stage<CL>(s, m, leader);
    }
    if (leader) umma_commit(bar_at(s, BAR_A_FREE + slot));
    __syncwarp();
  }
  if (leader) umma_commit(bar_at(s, BAR_D_READY + d));
  __syncwarp();
  ++m.jobctr;
}

// ---- row-thread helpers --------------------------------------------------------------------------------
struct EpiState { uint32_t jobctr; uint32_t afree_bits; };   // afree_bits: parity to wait on next, per slot

__device__ __forceinline__ void slab_begin(const TcShared& s, EpiState& e, int slot, bool wait_free) {
  if (wait_free) mbar_wait(bar_at(s, BAR_A_FREE + slot), (e.afree_bits >> slot) & 1, 300 + slot);
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
// barrier among the NEPI row threads only (named barrier 1)
template <int NEPI>
__device__ __forceinline__ void epi_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
}
__device__ __forceinline__ void epi_release_d(const TcShared& s, EpiState& e) {
  tc_fence_before();
  mbar_arrive(bar_at(s, BAR_D_FREE + (e.jobctr & 1)));
  ++e.jobctr;
}

// Hidden-layer epilogue: next A = relu(D + bias).  HALVES = 1: this thread converts all 64 columns of
// every slab; HALVES = 2: only columns [32*half, 32*half + 32).  The TMEM load of the next slab is in
// flight while the current one is converted and stored.
template <bool WAIT_FREE, int HALVES>
__device__ __forceinline__ void epi_hidden(const TcShared& s, EpiState& e, uint32_t lane_taddr, int row, int half,
                                           const float* __restrict__ bias_s) {
  const uint32_t d = epi_wait_d(s, e);
  constexpr int NCH = 4 * (2 / HALVES);          // 32-column chunks this thread handles
  uint32_t buf[2][32];
  auto col_of = [&](int c) { return HALVES == 2 ? (c * 64 + half * 32) : (c * 32); };
  tmem_ld32_issue(lane_taddr + d * 256 + col_of(0), buf[0]);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    tmem_ld32_wait(buf[c & 1]);
    if (c + 1 < NCH) tmem_ld32_issue(lane_taddr + d * 256 + col_of(c + 1), buf[(c + 1) & 1]);
    const int col = col_of(c), sl = col >> 6;
    if (HALVES == 2 || (c & 1) == 0) slab_begin(s, e, sl, WAIT_FREE);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(__uint_as_float(buf[c & 1][i]) + bias_s[col + i], 0.0f);
    a_store32(s.a_hi + sl * SLAB_BYTES, s.a_lo + sl * SLAB_BYTES, row, col & 63, v);
    if (HALVES == 2 || (c & 1) == 1) slab_done(s, sl);
  }
  epi_release_d(s, e);
}

}  // namespace ciaosr
