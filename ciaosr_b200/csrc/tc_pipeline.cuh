// Shared pipeline of the tcgen05 kernels (head_tc.cu, gemm_tc.cuh): smem carve-up, mbarrier
// protocol, weight producer, UMMA job issuer, accumulator drain helpers.
//
// One CTA per SM.  smem: 4 A-operand slots (fp16 hi + lo, [128 rows x 64 K] SW128 slabs), a 4-stage
// ring of 16 KB weight slabs ([128 N x 64 K]; a weight "unit" = its hi slab then its lo slab),
// 24 KB of per-kernel constants, 20 mbarriers.  TMEM: all 512 columns = two 128 x 256 fp32
// accumulators D[0], D[1], used alternately by consecutive jobs.
// A "job" = D[j&1][:, 0:128*units) = A[128, 64*nslabs] . W[128*units, 64*nslabs]^T evaluated as three
// fp16 UMMAs per product term (A_lo.W_hi + A_hi.W_hi on the hi slab, A_hi.W_lo on the lo slab).
//
// Both column halves of an accumulator are produced by the same N = 256 UMMAs; the D_READY[d][half]
// pair is committed together at the end of the job.
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
//   A_READY[4]  NEPI/32 row warps   operand writers -> UMMA issuer      (per slab slot)
//   A_FREE[4]   1 (tcgen05.commit)  UMMA issuer -> operand writers      (only waited on when K > 256)
//   D_READY[2][2] 1 (tcgen05.commit) UMMA issuer -> row threads         (one 128-column half of an accumulator complete)
//   D_FREE[2]   NEPI/32 row warps   row threads -> UMMA issuer          (accumulator drained)
#pragma once
#include "tc_common.cuh"
#include "tma.cuh"

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
              BAR_D_FREE = 20, BAR_W_FULL2 = 22, BAR_W_EMPTY2 = 24, N_BARS = 26;     // D_READY[d][n-half]: 16 + 2 d + nh
// Ring stages 4, 5 (W2 = true only): a SECOND pair of W_hi stages.  The hi stages of a K-slab are read by 8 of its
// 12 UMMAs, so with one pair they are free for only ~1/3 of a slab's issue time -- less than the L2 -> smem latency
// of the next slab's 32 KB, which stalled the issuer on W_FULL (round 1: 4.1 k cycles per slab against 2.3 k of
// UMMA issue in the long-K GEMMs).  Alternating two hi pairs gives every hi load a whole slab of slack; the lo pair
// (4 UMMAs per slab) already had 2/3 of a slab.  Used by tc_gemm (which has no per-kernel constants and can afford
// the extra 32 KB); the head kernels keep the 4-stage ring.
__host__ __device__ constexpr int w_full_bar(int stage) { return stage < 4 ? BAR_W_FULL + stage : BAR_W_FULL2 + stage - 4; }
__host__ __device__ constexpr int w_empty_bar(int stage) { return stage < 4 ? BAR_W_EMPTY + stage : BAR_W_EMPTY2 + stage - 4; }
constexpr int SM_SLOT = SM_BAR + N_BARS * 8;            // TMEM base address (4 B, padded to 16)
static_assert(SM_BAR + N_BARS * 8 + 16 + 1024 * 4 <= 227 * 1024, "head kernels' shared memory exceeds 227 KB");
constexpr int SM_XCHG = SM_SLOT + 16;                   // 1024 floats of row-thread exchange space
constexpr int SM_TOTAL = SM_XCHG + 1024 * 4;
constexpr int EPI_T0 = 128;                             // first row thread (warps 0..3 are control warps)

struct TcShared {
  uint32_t a_hi, a_lo, w, bar;
  float* consts;
  float* xchg;
  int pair_rank = -1;      // >= 0: CTA-pair mode (cta_group::2 UMMAs issued by rank 0): this CTA's rank in the pair
  // product terms the issuer accumulates: 1 = A_lo.W_hi, 2 = A_hi.W_hi, 4 = A_hi.W_lo.  7 = the fp32-grade default; 2 / 3 are
  // the opt-in reduced-precision modes (CIAOSR_TC_TERMS, DESIGN.md 4 "cost / accuracy frontier"), outside the parity tolerance
  uint32_t terms = 7;
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
// smem layout of tc_gemm (gemm_tc.cuh): no per-kernel constants; the 4-stage weight ring, then a 32 KB staging area in
// which every row warp transposes its 32 x 32 accumulator chunks so that global stores / loads of the epilogue are
// coalesced (see tc_gemm_kernel), then barriers etc.  (A second W_hi stage pair -- W2 above -- was measured in this
// slot first: it did not shorten the long-K GEMMs, r02e, so the space went to the epilogue, which did bound them.)
constexpr int GM_W_STAGES = 4;
constexpr int GM_STAGE = SM_W + GM_W_STAGES * SLAB_BYTES;       // 8 row warps x 4 KB
constexpr int GM_BAR = GM_STAGE + 8 * 4096;
constexpr int GM_SLOT = GM_BAR + (N_BARS * 8 + 15) / 16 * 16;
constexpr int GM_XCHG = GM_SLOT + 16;                   // 256 floats: the two partial sums per row of combining AGens
constexpr int GM_TOTAL = GM_XCHG + 256 * 4;
static_assert(GM_TOTAL <= 227 * 1024, "tc_gemm shared memory exceeds the 227 KB a CTA may opt in to");
__device__ __forceinline__ TcShared tc_carve_gemm(uint8_t* smem) {
  TcShared s;
  const uint32_t base = smem_u32(smem);
  s.a_hi = base + SM_A_HI; s.a_lo = base + SM_A_LO; s.w = base + SM_W;
  s.bar = base + GM_BAR;
  s.consts = nullptr;
  s.xchg = reinterpret_cast<float*>(smem + GM_XCHG);
  return s;
}

template <int CL, int NEPI>
__device__ __forceinline__ uint32_t tc_prologue(const TcShared& s, uint8_t* smem, int slot_off = SM_SLOT) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES + 2; ++i) { mbar_init(bar_at(s, w_full_bar(i)), 1); mbar_init(bar_at(s, w_empty_bar(i)), CL); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_at(s, BAR_A_READY + i), NEPI / 32); mbar_init(bar_at(s, BAR_A_FREE + i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(bar_at(s, BAR_D_READY + i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_at(s, BAR_D_FREE + i), NEPI / 32);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(smem) + slot_off, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // peer barriers are initialised before any multicast reaches them
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem + slot_off);
}
template <int CL>
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA exits while a peer may still signal its barriers
  if ((threadIdx.x >> 5) == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ---- weight producer (warp 0; the whole warp walks the loop, lane 0 issues) -------------------------
// Ring protocol: one K-slab of a job occupies exactly one revolution of the 4-stage ring:
//   stage 0 = W_hi rows   0..127, stage 1 = W_hi rows 128..255   (adjacent: one N = 256 B operand)
//   stage 2 = W_lo rows   0..127, stage 3 = W_lo rows 128..255
// Jobs with a single 128-column unit leave stages 1 and 3 empty (plain arrive, no bytes).
// bits: per ring stage, the parity of its NEXT use (0 on a fresh barrier); slabs: K-slabs streamed so far (the hi
// pair alternates with it when W2).  The first member keeps its old name / meaning for the 4-stage users.
struct ProdState { uint32_t bits; uint32_t slabs; };

// stream the weights of one job: units are stored (K-slab major) as [hi slab][lo slab] pairs
template <int CL, bool W2 = false>
__device__ __forceinline__ void produce_job(const TcShared& s, ProdState& ps, const uint8_t* blob, int nslabs,
                                            int units, uint32_t cta_rank) {
  const bool leader = (threadIdx.x & 31) == 0;
  for (int sl = 0; sl < nslabs; ++sl) {
    const int hi0 = (W2 && (ps.slabs & 1)) ? 4 : 0;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      const int u = part & 1, lo = part >> 1;
      const int stage = lo ? part : hi0 + part;
      mbar_wait(bar_at(s, w_empty_bar(stage)), ((ps.bits >> stage) & 1) ^ 1, 100 + stage);
      ps.bits ^= 1u << stage;
      if (leader) {
        const uint32_t full = bar_at(s, w_full_bar(stage));
        if (u < units) {
          mbar_arrive_expect_tx(full, SLAB_BYTES);
          const uint8_t* src = blob + ((size_t)(sl * units + u) * 2 + lo) * SLAB_BYTES;
          const uint32_t dst = s.w + stage * SLAB_BYTES;
          if (CL == 1) bulk_g2s(dst, src, SLAB_BYTES, full);
          else if ((uint32_t)u == cta_rank) bulk_g2s_mc(dst, src, SLAB_BYTES, full, (uint16_t)((1u << CL) - 1));
        } else {
          mbar_arrive(full);
        }
      }
      __syncwarp();
    }
    ++ps.slabs;
  }
}

// One job whose A operand already sits in HBM as two row-major 16-bit matrices (hi, lo halves): the producer
// lands each [128 rows x 64 K] slab pair in the operand slots with two 2-D TMA tile loads (128B swizzle = the
// UMMA layout; rows / columns outside the matrix read as zero) and interleaves them with the job's weight
// slabs.  A slabs run up to 3 ahead of the weights (4 operand slots): the warp blocks only for the slab it
// needs now.  A_READY counts NEPI/32 arrivals (row-warp-written slabs); for a TMA slab the producer supplies
// all of them itself, one carrying the transaction byte count.  `afree_bits`: parity to wait on next per slot.
template <int CL, int NEPI, bool W2 = false>
__device__ __forceinline__ void produce_job_tma_a(const TcShared& s, ProdState& ps, uint32_t& afree_bits,
                                                  const uint8_t* blob, int nslabs, int units, uint32_t cta_rank,
                                                  const CUtensorMap* map_hi, const CUtensorMap* map_lo, int kslab0,
                                                  int row0) {
  const int lane = threadIdx.x & 31;
  int a_next = 0;
  for (int sl = 0; sl < nslabs; ++sl) {
    while (a_next < nslabs && a_next <= sl + 3) {
      const int slot = a_next & 3;
      const uint32_t fr = bar_at(s, BAR_A_FREE + slot), par = (afree_bits >> slot) & 1;
      if (a_next == sl) mbar_wait(fr, par, 120 + slot);
      else if (!__shfl_sync(0xffffffffu, (int)mbar_try_wait(fr, par), 0)) break;
      afree_bits ^= 1u << slot;
      if (lane == 0) {
        const uint32_t rdy = bar_at(s, BAR_A_READY + slot);
        mbar_arrive_expect_tx(rdy, 2 * SLAB_BYTES);
        tma_load_2d(s.a_hi + slot * SLAB_BYTES, map_hi, (kslab0 + a_next) * KSLAB, row0, rdy);
        tma_load_2d(s.a_lo + slot * SLAB_BYTES, map_lo, (kslab0 + a_next) * KSLAB, row0, rdy);
        for (int k = 1; k < NEPI / 32; ++k) mbar_arrive(rdy);
      }
      __syncwarp();
      ++a_next;
    }
    produce_job<CL, W2>(s, ps, blob + (size_t)sl * units * UNIT_BYTES, 1, units, cta_rank);
  }
}

// ---- UMMA issuer (warp 1; the whole warp walks the loop, lane 0 issues and commits) -------------------
// wbits: per ring stage, the parity of the W_FULL completion to wait for next; slabs: K-slabs issued so far
struct MmaState { uint32_t wbits; uint32_t jobctr; uint32_t aready_bits; uint32_t slabs = 0; };

// descriptor for a slab at smem address `addr` (+ k-step offset): only the low word varies
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_lo(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}

// The twelve UMMAs of one K-slab as TWO asm statements (8 on the W_hi stage pair, 4 on W_lo).  Every separate `asm volatile`
// with register operands costs the issuing thread ~10 SASS instructions (R2UR moves into uniform registers around each
// UTCHMMA) on a sub-partition it shares with busy row warps; with the descriptor arithmetic inside the block that is paid
// once per group (r03h: the issuer's instruction stream, not the tensor pipe, paced the convolution kernel).
#define CIAOSR_UMMA_SLAB_FNS(NAME, CG)                                                                                        \
  __device__ __forceinline__ void NAME##_hi(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_hi, uint32_t idesc,         \
                                            uint32_t acc_first) {                                                            \
    asm volatile(                                                                                                            \
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 x, y, z;\n\t"                                                 \
        "setp.ne.b32 p, %6, 0;\n\t"                                                                                          \
        "mov.b64 db, {%3, %5};\n\tmov.b64 da, {%1, %5};\n\t"                                                                 \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, p;\n\t"                                                    \
        "mov.b64 da, {%2, %5};\n\t"                                                                                          \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "add.u32 x, %1, 2;\n\tadd.u32 y, %2, 2;\n\tadd.u32 z, %3, 2;\n\t"                                                    \
        "mov.b64 db, {z, %5};\n\tmov.b64 da, {x, %5};\n\t"                                                                   \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "mov.b64 da, {y, %5};\n\t"                                                                                           \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "add.u32 x, %1, 4;\n\tadd.u32 y, %2, 4;\n\tadd.u32 z, %3, 4;\n\t"                                                    \
        "mov.b64 db, {z, %5};\n\tmov.b64 da, {x, %5};\n\t"                                                                   \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "mov.b64 da, {y, %5};\n\t"                                                                                           \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "add.u32 x, %1, 6;\n\tadd.u32 y, %2, 6;\n\tadd.u32 z, %3, 6;\n\t"                                                    \
        "mov.b64 db, {z, %5};\n\tmov.b64 da, {x, %5};\n\t"                                                                   \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t"                                                    \
        "mov.b64 da, {y, %5};\n\t"                                                                                           \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %4, 1;\n\t}" ::"r"(d),                                         \
        "r"(a_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(DESC_HI), "r"(acc_first)                                            \
        : "memory");                                                                                                         \
  }                                                                                                                          \
  __device__ __forceinline__ void NAME##_lo(uint32_t d, uint32_t a_hi, uint32_t b_lo, uint32_t idesc) {                      \
    asm volatile(                                                                                                            \
        "{\n\t.reg .b64 da, db;\n\t.reg .b32 x, z;\n\t"                                                                     \
        "mov.b64 db, {%2, %4};\n\tmov.b64 da, {%1, %4};\n\t"                                                                 \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %3, 1;\n\t"                                                    \
        "add.u32 x, %1, 2;\n\tadd.u32 z, %2, 2;\n\tmov.b64 db, {z, %4};\n\tmov.b64 da, {x, %4};\n\t"                         \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %3, 1;\n\t"                                                    \
        "add.u32 x, %1, 4;\n\tadd.u32 z, %2, 4;\n\tmov.b64 db, {z, %4};\n\tmov.b64 da, {x, %4};\n\t"                         \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %3, 1;\n\t"                                                    \
        "add.u32 x, %1, 6;\n\tadd.u32 z, %2, 6;\n\tmov.b64 db, {z, %4};\n\tmov.b64 da, {x, %4};\n\t"                         \
        "tcgen05.mma.cta_group::" CG ".kind::f16 [%0], da, db, %3, 1;\n\t}" ::"r"(d),                                         \
        "r"(a_hi), "r"(b_lo), "r"(idesc), "r"(DESC_HI)                                                                       \
        : "memory");                                                                                                         \
  }
CIAOSR_UMMA_SLAB_FNS(umma_slab1, "1")
CIAOSR_UMMA_SLAB_FNS(umma_slab2, "2")
#undef CIAOSR_UMMA_SLAB_FNS

template <int CL>
__device__ __forceinline__ void release_stage(const TcShared& s, int stage) {
  if (CL == 1) umma_commit(bar_at(s, w_empty_bar(stage)));
  else umma_commit_mc(bar_at(s, w_empty_bar(stage)), (uint16_t)((1u << CL) - 1));
}

// one job: D[jobctr & 1][:, 0 : 128*units) = A (nslabs x 64 K) * W^T ; A slabs cycle through the 4 smem slots.
// One UMMA covers all 128*units columns (N = 256 for full jobs): 12 instructions per K-slab, each worth
// 128 tensor-core cycles, so the single issuing thread is never the bottleneck.
// a_release = false: the A slabs stay in their slots for the next job (same rows, next N-chunk); A_FREE is then
// committed only by the last job that reads them, so waiters see exactly one completion per slot use.
template <int CL, bool W2 = false>
__device__ __forceinline__ void mma_job(const TcShared& s, uint32_t tmem_base, MmaState& m, int nslabs,
                                        int units, bool a_new, bool a_release = true) {
  const bool leader = (threadIdx.x & 31) == 0;
  const uint32_t idesc = make_idesc_split(ROWS, UNIT_N * units);
  const uint32_t d = m.jobctr & 1, n = m.jobctr >> 1;
  mbar_wait(bar_at(s, BAR_D_FREE + d), (n + 1) & 1, 200);
  tc_fence_after();
  const uint32_t dcol = tmem_base + d * 256;
  const uint32_t b_lo = desc_lo(s.w + 2 * SLAB_BYTES);
  for (int sl = 0; sl < nslabs; ++sl) {
    const int slot = sl & 3;
    const int hi0 = (W2 && (m.slabs & 1)) ? 4 : 0;
    const uint32_t b_hi = desc_lo(s.w + hi0 * SLAB_BYTES);
    if (a_new) {
      mbar_wait(bar_at(s, BAR_A_READY + slot), (m.aready_bits >> slot) & 1, 210 + slot);
      m.aready_bits ^= 1u << slot;
    }
    TC_TRACE(1000 + sl);          // issuer: operand slab sl available
    const uint32_t a_hi = desc_lo(s.a_hi + slot * SLAB_BYTES), a_lo = desc_lo(s.a_lo + slot * SLAB_BYTES);
    // W_hi (stages 0,1): A_lo.W_hi (small term first) and A_hi.W_hi
    mbar_wait(bar_at(s, w_full_bar(hi0)), (m.wbits >> hi0) & 1, 220);
    mbar_wait(bar_at(s, w_full_bar(hi0 + 1)), (m.wbits >> (hi0 + 1)) & 1, 221);
    m.wbits ^= 3u << hi0;
    tc_fence_after();
    if (leader) {
      if (s.terms == 7) {
        umma_slab1_hi(dcol, a_lo, a_hi, b_hi, idesc, sl != 0 ? 1u : 0u);
      } else {                                         // reduced-precision modes: A_hi.W_hi (+ A_lo.W_hi)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_lo(dcol, a_hi + 2 * ks, b_hi + 2 * ks, idesc, (sl | ks) != 0 ? 1u : 0u);
          if (s.terms & 1) umma_lo(dcol, a_lo + 2 * ks, b_hi + 2 * ks, idesc, 1u);
        }
      }
      release_stage<CL>(s, hi0);
      release_stage<CL>(s, hi0 + 1);
    }
    __syncwarp();
    // W_lo (stages 2,3): A_hi.W_lo
    mbar_wait(bar_at(s, BAR_W_FULL + 2), (m.wbits >> 2) & 1, 222);
    mbar_wait(bar_at(s, BAR_W_FULL + 3), (m.wbits >> 3) & 1, 223);
    m.wbits ^= 3u << 2;
    tc_fence_after();
    if (leader) {
      if (s.terms & 4) umma_slab1_lo(dcol, a_hi, b_lo, idesc);
      release_stage<CL>(s, 2);
      release_stage<CL>(s, 3);
      if (a_release) umma_commit(bar_at(s, BAR_A_FREE + slot));
    }
    __syncwarp();
    ++m.slabs;
  }
  if (leader) {
    umma_commit(bar_at(s, BAR_D_READY + 2 * d));
    if (units == 2) umma_commit(bar_at(s, BAR_D_READY + 2 * d + 1));   // waited on only by 2-unit jobs
  }
  __syncwarp();
  TC_TRACE(1010);                 // issuer: job fully issued
  ++m.jobctr;
}


// =====================================================================================================================
// CTA-pair mode (cta_group::2): see tc_common.cuh.  Per CTA the weight ring holds HALF of every N = 256 operand:
// 4 stages of 16 KB = 2 K-slabs deep ([hi | lo] x 2), twice the lookahead of the single-CTA ring in the same space.
//   W_FULL[stage]   lives in the LEADER: 1 arrival (its producer's expect_tx) + the bytes of BOTH CTAs' tile loads
//   W_EMPTY[stage]  in each CTA, 1 arrival: the leader's tcgen05.commit, multicast to both
//   A_READY[slot]   in the leader, 2 x NEPI/32 arrivals (the peer's row warps arrive remotely)
//   A_FREE / D_READY  in each CTA: multicast commits;   D_FREE[d] in the leader, 2 x NEPI/32 arrivals
// =====================================================================================================================
template <int NEPI>
__device__ __forceinline__ uint32_t tc_prologue_pair(const TcShared& s, uint8_t* smem, int slot_off = SM_SLOT) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES + 2; ++i) { mbar_init(bar_at(s, w_full_bar(i)), 1); mbar_init(bar_at(s, w_empty_bar(i)), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(bar_at(s, BAR_A_READY + i), 2 * NEPI / 32); mbar_init(bar_at(s, BAR_A_FREE + i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(bar_at(s, BAR_D_READY + i), 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_at(s, BAR_D_FREE + i), 2 * NEPI / 32);
    fence_mbar_init();
  }
  cluster_sync_all();                      // both CTAs are resident before the pair allocation
  if (warp == 2) tmem_alloc_cg2(smem_u32(smem) + slot_off, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // peer barriers are initialised before anything signals them
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(smem + slot_off);
}
__device__ __forceinline__ void tc_teardown_pair(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // no CTA frees or exits while the pair may still use its TMEM / barriers
  if ((threadIdx.x >> 5) == 2) { tc_fence_after(); tmem_dealloc_cg2(tmem_base, 512); }
}

// Weights of one job for this CTA's half of the pair.  `wmap`: un-swizzled row map over the blob (tma_make_map_linear_rows);
// `row0`: first 128-byte row of the job in the blob (units are stored K-slab major as [hi slab][lo slab] pairs of 128 rows).
// N = 256 jobs: the CTA takes unit `rank` (128 weight rows); N = 128 jobs: rows 64 rank .. of the single unit.
__device__ __forceinline__ void produce_job_pair(const TcShared& s, ProdState& ps, const CUtensorMap* wmap, long long row0,
                                                 int nslabs, int units) {
  const bool leader_lane = (threadIdx.x & 31) == 0;
  const int rank = s.pair_rank;
  const uint32_t bytes = units == 2 ? SLAB_BYTES : SLAB_BYTES / 2;      // per CTA per stage
  for (int sl = 0; sl < nslabs; ++sl) {
    const int sp = (int)(ps.slabs & 1) * 2;
#pragma unroll
    for (int lo = 0; lo < 2; ++lo) {
      const int stage = sp + lo;
      mbar_wait(bar_at(s, w_empty_bar(stage)), ((ps.bits >> stage) & 1) ^ 1, 100 + stage);
      ps.bits ^= 1u << stage;
      if (leader_lane) {
        const uint32_t full = bar_at(s, w_full_bar(stage));
        if (TC_DBG(2)) {
          if (rank == 0) mbar_arrive(full);
        } else {
        if (rank == 0) mbar_arrive_expect_tx(full, 2 * bytes);
        const long long r = row0 + ((long long)(sl * units + (units == 2 ? rank : 0)) * 2 + lo) * ROWS +
                            (units == 2 ? 0 : 64 * rank);
        const uint32_t dst = s.w + stage * SLAB_BYTES;
        tma_load_2d_cg2(dst, wmap, 0, (int)r, full);
        if (units == 2) tma_load_2d_cg2(dst + SLAB_BYTES / 2, wmap, 0, (int)r + 64, full);
        }
      }
      __syncwarp();
    }
    ++ps.slabs;
  }
}

// CTA-pair form of produce_job_tma_a: each CTA lands ITS 128 operand rows of every K-slab in its own slots (2-D tensor
// loads in the .cta_group::2 form, so the bytes of both CTAs count on the LEADER's A_READY, which carries
// 2 x NEPI/32 arrivals: the leader's producer supplies all of them, one with the transaction bytes of both CTAs) and
// interleaves them with its half of the slab's weights.  A_FREE is local to each CTA (multicast commit of the leader).
template <int NEPI>
__device__ __forceinline__ void produce_job_tma_a_pair(const TcShared& s, ProdState& ps, uint32_t& afree_bits,
                                                       const CUtensorMap* wmap, long long wrow0, int nslabs, int units,
                                                       const CUtensorMap* map_hi, const CUtensorMap* map_lo, int row0) {
  const int lane = threadIdx.x & 31;
  int a_next = 0;
  for (int sl = 0; sl < nslabs; ++sl) {
    while (a_next < nslabs && a_next <= sl + 3) {
      const int slot = a_next & 3;
      const uint32_t fr = bar_at(s, BAR_A_FREE + slot), par = (afree_bits >> slot) & 1;
      if (a_next == sl) mbar_wait(fr, par, 120 + slot);
      else if (!__shfl_sync(0xffffffffu, (int)mbar_try_wait(fr, par), 0)) break;
      afree_bits ^= 1u << slot;
      if (lane == 0) {
        const uint32_t rdy = bar_at(s, BAR_A_READY + slot);
        if (s.pair_rank == 0) {
          mbar_arrive_expect_tx(rdy, 4 * SLAB_BYTES);                 // hi + lo slabs of both CTAs
          for (int k = 1; k < 2 * NEPI / 32; ++k) mbar_arrive(rdy);
        }
        tma_load_2d_cg2(s.a_hi + slot * SLAB_BYTES, map_hi, a_next * KSLAB, row0, rdy);
        tma_load_2d_cg2(s.a_lo + slot * SLAB_BYTES, map_lo, a_next * KSLAB, row0, rdy);
      }
      __syncwarp();
      ++a_next;
    }
    produce_job_pair(s, ps, wmap, wrow0 + (long long)sl * units * 2 * ROWS, 1, units);
  }
}

__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
#if defined(CIAOSR_TC_TIMING) && !defined(CIAOSR_TC_TRACE_ONLY)
  const long long tt = clock64();
  struct Rec { int tag; long long t; __device__ ~Rec() { tc_time_add(tag / 10, t); } } rec{tag, tt};
#endif
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {
      printf("ciaosr tcgen05: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag, blockIdx.x,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// one job of the pair, issued by the leader's warp 1:  D[jobctr & 1] of BOTH CTAs = A (256 rows) * W^T
__device__ __forceinline__ void mma_job_pair(const TcShared& s, uint32_t tmem_base, MmaState& m, int nslabs, int units,
                                             bool a_new, bool a_release = true) {
  const bool leader = (threadIdx.x & 31) == 0;
  const uint32_t idesc = make_idesc_split(2 * ROWS, UNIT_N * units);
  const uint32_t d = m.jobctr & 1, n = m.jobctr >> 1;
  mbar_wait(bar_at(s, BAR_D_FREE + d), (n + 1) & 1, 200);        // CTA-scope acquire: see mbar_arrive_cluster
  tc_fence_after();
  const uint32_t dcol = tmem_base + d * 256;
  for (int sl = 0; sl < nslabs; ++sl) {
    const int slot = sl & 3;
    const int sp = (int)(m.slabs & 1) * 2;
    if (a_new) {
      mbar_wait(bar_at(s, BAR_A_READY + slot), (m.aready_bits >> slot) & 1, 210 + slot);
      m.aready_bits ^= 1u << slot;
    }
    TC_TRACE(1000 + sl);          // issuer: operand slab sl available
    const uint32_t a_hi = desc_lo(s.a_hi + slot * SLAB_BYTES), a_lo = desc_lo(s.a_lo + slot * SLAB_BYTES);
    const uint32_t b_hi = desc_lo(s.w + sp * SLAB_BYTES), b_lo = desc_lo(s.w + (sp + 1) * SLAB_BYTES);
    mbar_wait(bar_at(s, w_full_bar(sp)), (m.wbits >> sp) & 1, 220);
    m.wbits ^= 1u << sp;
    tc_fence_after();
    if (leader) {
      if (s.terms == 7) {
        umma_slab2_hi(dcol, a_lo, a_hi, b_hi, idesc, sl != 0 ? 1u : 0u);
      } else {                                         // reduced-precision modes: A_hi.W_hi (+ A_lo.W_hi)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_cg2(dcol, a_hi + 2 * ks, b_hi + 2 * ks, DESC_HI, idesc, (sl | ks) != 0 ? 1u : 0u);
          if (s.terms & 1) umma_cg2(dcol, a_lo + 2 * ks, b_hi + 2 * ks, DESC_HI, idesc, 1u);
        }
      }
      umma_commit_cg2(bar_at(s, w_empty_bar(sp)), 3);
    }
    __syncwarp();
    mbar_wait(bar_at(s, w_full_bar(sp + 1)), (m.wbits >> (sp + 1)) & 1, 222);
    m.wbits ^= 1u << (sp + 1);
    tc_fence_after();
    if (leader) {
      if (s.terms & 4) umma_slab2_lo(dcol, a_hi, b_lo, idesc);
      umma_commit_cg2(bar_at(s, w_empty_bar(sp + 1)), 3);
      if (a_release) umma_commit_cg2(bar_at(s, BAR_A_FREE + slot), 3);
    }
    __syncwarp();
    ++m.slabs;
  }
  if (leader) {
    umma_commit_cg2(bar_at(s, BAR_D_READY + 2 * d), 3);
    if (units == 2) umma_commit_cg2(bar_at(s, BAR_D_READY + 2 * d + 1), 3);
  }
  __syncwarp();
  TC_TRACE(1010);                 // issuer: job fully issued
  ++m.jobctr;
}

// ---- CTA-pair mode, N-split schedule ------------------------------------------------------------------------------
// Measured (tools/ubench/mma_bench.cu, r02x): a dependent chain of N = 256 UMMAs runs at the full 128 cycles each (M = 128
// per CTA), N <= 128 at a 74-cycle issue floor, and switching accumulators between consecutive UMMAs costs a ~180-cycle
// drain.  In the kernels the N = 256 UMMAs took 169 cycles (shared-memory bandwidth: operand reads + weight writes +
// the row threads' operand stores), and -- worse -- a layer's epilogue could not start before ALL of its UMMAs were done
// (one N = 256 instruction covers both accumulator halves), so a third of every layer's period had an idle tensor pipe
// (r02w trace: 3.0 k cycles of row work exposed + 0.9 k of latencies per 12.0 k).  Here every layer is issued as two
// N = 128 column halves H0, H1 over two K-slab groups, in the order
//     (H0: s0 s1) (H1: s0 s1) (H0: s2 s3) -> commit D_READY[H0]     (H1: s2 s3) -> commit D_READY[H1]
// so that the row threads convert H0 (= the next layer's K-slabs 0, 1) while the UMMAs of (H1: s2 s3) run, and the next
// layer's first four units only need those two slabs: the tensor pipe has work throughout.  Accumulators are switched 4
// times per layer.  One ring stage = one (K-slab, column half) unit of this CTA: [W_hi 64 rows | W_lo 64 rows] = 16 KB.
__device__ __forceinline__ void produce_job_pair_split(const TcShared& s, ProdState& ps, const CUtensorMap* wmap,
                                                       long long row0, int units) {
  const bool leader_lane = (threadIdx.x & 31) == 0;
  const int rank = s.pair_rank;
  for (int g = 0; g < 2; ++g)
    for (int h = 0; h < units; ++h)
      for (int sl = 2 * g; sl < 2 * g + 2; ++sl) {
        const int stage = (int)(ps.slabs & 3);
        mbar_wait(bar_at(s, w_empty_bar(stage)), ((ps.bits >> stage) & 1) ^ 1, 100 + stage);
        ps.bits ^= 1u << stage;
        if (leader_lane) {
          const uint32_t full = bar_at(s, w_full_bar(stage));
          if (TC_DBG(2)) {
            if (rank == 0) mbar_arrive(full);
          } else {
          if (rank == 0) mbar_arrive_expect_tx(full, 2 * SLAB_BYTES);          // both CTAs' [hi | lo] halves
          const long long r = row0 + ((long long)(sl * units + h) * 2) * ROWS + 64 * rank;
          const uint32_t dst = s.w + stage * SLAB_BYTES;
          tma_load_2d_cg2(dst, wmap, 0, (int)r, full);                         // W_hi rows [64 rank, 64 rank + 64) of the unit
          tma_load_2d_cg2(dst + SLAB_BYTES / 2, wmap, 0, (int)r + ROWS, full); // W_lo, same rows
          }
        }
        __syncwarp();
        ++ps.slabs;
      }
}

__device__ __forceinline__ void mma_job_pair_split(const TcShared& s, uint32_t tmem_base, MmaState& m, int units,
                                                   bool a_new) {
  // Issue-side overheads matter at N = 128 (74-cycle issue floor per UMMA, r02z mma_pattern: every tcgen05.commit costs
  // ~76 cycles of tensor time and every wait + tcgen05.fence ~66), so this path commits once per unit (W_EMPTY; A_FREE is
  // never waited on by the pair kernel and D_READY is per half) and fences only after D_FREE -- operand visibility comes
  // from the mbarrier (TMA complete_tx / fence.proxy.async + arrive), not from a tcgen05 fence.
  const bool leader = (threadIdx.x & 31) == 0;
  const uint32_t idesc = make_idesc_split(2 * ROWS, UNIT_N);
  const uint32_t d = m.jobctr & 1, n = m.jobctr >> 1;
  mbar_wait(bar_at(s, BAR_D_FREE + d), (n + 1) & 1, 200);
  tc_fence_after();
  for (int g = 0; g < 2; ++g)
    for (int h = 0; h < units; ++h) {
      const uint32_t dcol = tmem_base + d * 256 + h * UNIT_N;
      for (int sl = 2 * g; sl < 2 * g + 2; ++sl) {
        if (a_new && h == 0) {
          mbar_wait(bar_at(s, BAR_A_READY + sl), (m.aready_bits >> sl) & 1, 210 + sl);
          m.aready_bits ^= 1u << sl;
          TC_TRACE(1000 + sl);          // issuer: operand slab sl available
        }
        const int stage = (int)(m.slabs & 3);
        const uint32_t a_hi = desc_lo(s.a_hi + sl * SLAB_BYTES), a_lo = desc_lo(s.a_lo + sl * SLAB_BYTES);
        const uint32_t b_hi = desc_lo(s.w + stage * SLAB_BYTES), b_lo = desc_lo(s.w + stage * SLAB_BYTES + SLAB_BYTES / 2);
        mbar_wait(bar_at(s, w_full_bar(stage)), (m.wbits >> stage) & 1, 220);
        m.wbits ^= 1u << stage;
        if (leader) {
          umma_slab2_hi(dcol, a_lo, a_hi, b_hi, idesc, sl != 0 ? 1u : 0u);
          umma_slab2_lo(dcol, a_hi, b_lo, idesc);
          umma_commit_cg2(bar_at(s, w_empty_bar(stage)), 3);
          if (g == 1 && sl == 3) umma_commit_cg2(bar_at(s, BAR_D_READY + 2 * d + h), 3);
        }
        __syncwarp();
        ++m.slabs;
      }
    }
  TC_TRACE(1010);                 // issuer: job fully issued
  ++m.jobctr;
}

// ---- row-thread helpers --------------------------------------------------------------------------------
// afree_bits / dready_bits: parity to wait on next, per operand slot / per (accumulator, column half)
struct EpiState { uint32_t jobctr; uint32_t afree_bits; uint32_t dready_bits; };

__device__ __forceinline__ void slab_begin(const TcShared& s, EpiState& e, int slot, bool wait_free) {
  if (wait_free) mbar_wait(bar_at(s, BAR_A_FREE + slot), (e.afree_bits >> slot) & 1, 300 + slot);
  e.afree_bits ^= 1u << slot;
}
// Publish operand slabs to the UMMA issuer.  Every writer thread makes its own st.shared visible to
// the async proxy (fence.proxy.async, ONE per call: it drains the thread's outstanding shared stores
// and is the most expensive instruction of the epilogue), the warp converges, and one lane arrives
// (A_READY counts warps, not threads: same-address mbarrier arrivals serialise).
// arrive on the barrier the UMMA issuer waits on: the CTA's own, or the leader's in CTA-pair mode
__device__ __forceinline__ void arrive_issuer(const TcShared& s, int idx) {
  if (s.pair_rank <= 0) mbar_arrive(bar_at(s, idx));
  else mbar_arrive_cluster(bar_at(s, idx), 0);
}
__device__ __forceinline__ void slab_done(const TcShared& s, int slot) {
  fence_proxy_async();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) arrive_issuer(s, BAR_A_READY + slot);
}
__device__ __forceinline__ void slabs_done2(const TcShared& s, int slot_a, int slot_b) {
  fence_proxy_async();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    arrive_issuer(s, BAR_A_READY + slot_a);
    arrive_issuer(s, BAR_A_READY + slot_b);
  }
}
// wait until column half `nh` of the current job's accumulator is complete; returns the accumulator index
__device__ __forceinline__ uint32_t epi_wait_half(const TcShared& s, EpiState& e, int nh) {
  const uint32_t d = e.jobctr & 1, b = 2 * d + nh;
  mbar_wait(bar_at(s, BAR_D_READY + b), (e.dready_bits >> b) & 1, 310 + b);
  e.dready_bits ^= 1u << b;
  tc_fence_after();
  if (threadIdx.x == EPI_T0) TC_TRACE(2000 + nh);      // rows: accumulator half nh observed complete
  return d;
}
__device__ __forceinline__ uint32_t epi_wait_d(const TcShared& s, EpiState& e, int units) {
  uint32_t d = 0;
  for (int nh = 0; nh < units; ++nh) d = epi_wait_half(s, e, nh);
  return d;
}
// barrier among the NEPI row threads only (named barrier 1)
template <int NEPI>
__device__ __forceinline__ void epi_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory");
}
__device__ __forceinline__ void epi_release_d(const TcShared& s, EpiState& e) {
  if (threadIdx.x == EPI_T0) TC_TRACE(2010);           // rows: accumulator drained
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) arrive_issuer(s, BAR_D_FREE + (e.jobctr & 1));
  ++e.jobctr;
}

// Hidden-layer epilogue: next A = relu(D + bias).  HALVES = threads per row: 1: this thread converts all 64 columns of
// every slab; 2: columns [32*half, 32*half + 32); 4: columns [16*half, 16*half + 16) (`half` is then the quarter index).
// Columns 0..127 are converted as soon as the first accumulator half is complete.
// (not inlined: it is instantiated once and called per layer -- the fully unrolled body is ~400
// instructions, and inlining it 5x per tile made instruction fetch 17 % of the row warps' stall time)
// (state by VALUE: the body is full of `asm volatile(... ::: "memory")`, after each of which everything that lives in
// memory -- as by-reference arguments of a non-inlined function do -- is re-loaded from the stack: 34 LDL per call sat in
// the middle of the per-slab latency chain; in registers the pipeline state costs nothing)
template <bool WAIT_FREE, int HALVES>
__device__ __noinline__ EpiState epi_hidden(const TcShared s, EpiState e, uint32_t lane_taddr, int row, int half,
                                               const float* __restrict__ bias_s) {
  constexpr int CW = HALVES == 4 ? 16 : 32;      // columns per chunk
  constexpr int NCH = HID / (CW * HALVES);       // chunks this thread handles
  constexpr int PER_HALF = NCH / 2;
  uint32_t buf[2][CW];
  auto col_of = [&](int c) { return HALVES == 1 ? (c * 32) : (c * 64 + half * CW); };
  uint32_t d = 0;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c % PER_HALF == 0) {                     // first chunk of an accumulator half
      d = epi_wait_half(s, e, c / PER_HALF);
      tmem_ldN_issue(lane_taddr + d * 256 + col_of(c), buf[c & 1]);
    }
    tmem_ldN_wait(buf[c & 1]);
    if ((c + 1) % PER_HALF != 0) tmem_ldN_issue(lane_taddr + d * 256 + col_of(c + 1), buf[(c + 1) & 1]);
    const int col = col_of(c), sl = col >> 6;
    if (HALVES >= 2 || (c & 1) == 0) slab_begin(s, e, sl, WAIT_FREE);
    float v[CW];
    if (!TC_DBG(1)) {
#pragma unroll
    for (int j = 0; j < CW / 4; ++j) {
      const float4 b4 = reinterpret_cast<const float4*>(bias_s + col)[j];
      // bias add on pairs (FADD2); the ReLU is part of the operand split (split2_relu)
      unpack2(add2(pack2(__uint_as_float(buf[c & 1][4 * j]), __uint_as_float(buf[c & 1][4 * j + 1])), pack2(b4.x, b4.y)), v[4 * j], v[4 * j + 1]);
      unpack2(add2(pack2(__uint_as_float(buf[c & 1][4 * j + 2]), __uint_as_float(buf[c & 1][4 * j + 3])), pack2(b4.z, b4.w)), v[4 * j + 2], v[4 * j + 3]);
    }
    a_storeN<true>(s.a_hi + sl * SLAB_BYTES, s.a_lo + sl * SLAB_BYTES, row, col & 63, v);
    }
    if (HALVES >= 2) {
      // single-CTA mode: per slab -- the next layer's first UMMAs start one slab earlier (-3 % kernel time).  CTA-pair
      // mode: the M = 256 UMMAs are ~3x cheaper per row and the ROW threads are the critical path (r02n: issuer idle
      // 63 %), so they publish two slabs per fence.proxy.async (the most expensive step of this epilogue) instead.
      // [r03] ... except the FIRST slabs: the issuer idles until slab 0 arrives (r02w trace: 3.0 k of every 12.0 k-cycle
      // layer period), so slabs 0 and 1 are published one by one and only 2 + 3 share a fence.
      if (s.pair_rank < 0 || sl < 2) slab_done(s, sl);
      else if (c & 1) slabs_done2(s, sl - 1, sl);
    } else if (c & 1) slab_done(s, sl);
    if (threadIdx.x == EPI_T0) TC_TRACE(2020 + sl);    // rows: this warp's part of slab sl written (published if odd / single mode)
  }
  epi_release_d(s, e);
  return e;
}

}  // namespace ciaosr
