// sm_100a primitives for the tcgen05 engine: mbarrier, bulk async copy (TMA engine,
// UBLKCP), TMEM allocation, UMMA descriptors / issue / commit, TMEM loads.
// Hand-written inline PTX; descriptor bit layouts follow the PTX ISA 8.6 "tcgen05"
// matrix / instruction descriptor tables.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace ciaosr {
namespace tc {

constexpr int ROWS = 128;            // UMMA M (one TMEM lane per row)
constexpr int KSLAB = 64;            // 16-bit elements per 128-byte swizzle row
constexpr int SLAB_BYTES = ROWS * 128;       // one [128 x 64] 16-bit operand slab (SW128, K-major)
constexpr int UNIT_N = 128;          // weight rows per ring stage
constexpr int UNIT_BYTES = 2 * SLAB_BYTES;   // hi slab + lo slab
constexpr int HID = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifdef CIAOSR_TC_TIMING
// Diagnostic build only (tools/wait_breakdown.py): cycles each warp spends in mbarrier waits, by tag / 10.
static __device__ unsigned long long g_wait_cycles[64];
static __device__ unsigned long long g_wait_count[64];
__device__ __forceinline__ void tc_time_add(int cls, long long t0) {
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&g_wait_cycles[cls & 63], (unsigned long long)(clock64() - t0));
    atomicAdd(&g_wait_count[cls & 63], 1ull);
  }
}
#endif
#ifdef CIAOSR_TC_TIMING
// timeline trace of CTA 0 (issuer warp and first row warp): (tag, clock) events, tools/trace_pair.py
static __device__ unsigned long long g_trace[2 * 8192];
static __device__ unsigned int g_trace_n;
static __device__ int g_trace_on;
static __device__ int g_trace_req;        // host request; the traced kernel copies it into g_trace_on
__device__ __forceinline__ void tc_trace(int tag) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && g_trace_on) {
    const unsigned int i = atomicAdd(&g_trace_n, 1u);
    if (i < 8192) { g_trace[2 * i] = (unsigned long long)tag; g_trace[2 * i + 1] = (unsigned long long)clock64(); }
  }
}
// (tag, %globaltimer) event: pairs with a TC_TRACE of the same place to read the SM clock rate during the kernel
__device__ __forceinline__ void tc_trace_ns(int tag) {
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && g_trace_on) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    const unsigned int i = atomicAdd(&g_trace_n, 1u);
    if (i < 8192) { g_trace[2 * i] = (unsigned long long)tag; g_trace[2 * i + 1] = ns; }
  }
}
#define TC_TRACE_NS(tag) ::ciaosr::tc::tc_trace_ns(tag)
#define TC_TRACE(tag) ::ciaosr::tc::tc_trace(tag)
// experiment switches of the diagnostic build (ciaosr_debug_flags): bit 0 = row threads skip their arithmetic and operand
// stores (waits / arrivals kept), bit 1 = the weight producer of the CTA-pair kernel signals stages without loading them
static __device__ int g_dbg_flags;
#define TC_DBG(bit) (::ciaosr::tc::g_dbg_flags & (bit))
#else
#define TC_DBG(bit) 0
#define TC_TRACE_NS(tag) do {} while (0)
#define TC_TRACE(tag) do {} while (0)
#endif
// Bounded wait: a protocol bug must surface as a CUDA error (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
#if defined(CIAOSR_TC_TIMING) && !defined(CIAOSR_TC_TRACE_ONLY)      // TRACE_ONLY: timeline events without the wait counters
  const long long tt = clock64();
  struct Rec { int tag; long long t; __device__ ~Rec() { tc_time_add(tag / 10, t); } } rec{tag, tt};
#endif
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
      printf("ciaosr tcgen05: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag,
             blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk async copy global -> shared (TMA engine, no tensor map) -------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// ---- TMEM --------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 32 lanes x 32 columns of fp32: thread i of the warp gets lane (base + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue the load, do other work, then wait (the wait names the registers so that no
// use of them can be scheduled above it)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// 16-column forms (four row threads per row: each thread owns a 16-column quarter of every 64-column slab)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])
               :
               : "memory");
}
// width-generic spellings used by the row-thread code (N = 16 or 32 columns per chunk)
__device__ __forceinline__ void tmem_ldN_issue(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldN_issue(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld16_issue(taddr, r); }
__device__ __forceinline__ void tmem_ldN_wait(uint32_t (&r)[32]) { tmem_ld32_wait(r); }
__device__ __forceinline__ void tmem_ldN_wait(uint32_t (&r)[16]) { tmem_ld16_wait(r); }

// ---- clusters ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy delivered to the same smem offset (and mbarrier offset) of every CTA in `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}

// ---- UMMA ----------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle:
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (ignored for swizzled K-major, 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Operand element type of the hi/lo split (see split2 below): fp16 by default, bf16 with -DCIAOSR_SPLIT_BF16.
#ifdef CIAOSR_SPLIT_BF16
using split_t = __nv_bfloat16;
constexpr uint32_t SPLIT_FMT = 1;      // kind::f16 operand format field: 1 = BF16
#else
using split_t = __half;
constexpr uint32_t SPLIT_FMT = 0;      //                                  0 = F16
#endif
// Instruction descriptor, kind::f16: D fp32, A/B fp16 (or bf16), both K-major, dense.
//   [4,6) D format = 1 (F32) | [7,10) A format | [10,13) B format
//   [15] A major = 0 (K) | [16] B major = 0 (K) | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_split(int M, int N) {
  return (1u << 4) | (SPLIT_FMT << 7) | (SPLIT_FMT << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when every UMMA this thread issued so far has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// same, arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// ---- CTA pairs (tcgen05 cta_group::2) ------------------------------------------------------------------------------
// One UMMA spans both SMs of a cluster of two: M = 256 rows (each CTA supplies its 128 A rows from its own shared
// memory and receives its 128 accumulator rows in its own TMEM), and each CTA stages only HALF of the B operand (N / 2
// weight rows) at the same shared-memory offset -- the instruction reads both halves.  Per SM the operand traffic of a
// 128 x 256 x 16 tile drops from 12 KB to 8 KB, which is what keeps a single-CTA N = 256 UMMA at ~189 cycles instead of
// the 128 of its math (measured, profiles/r01e).  Only the leader CTA (cluster rank 0) issues; completion is
// multicast to the mbarriers of both CTAs; the peer's row threads signal the leader's barriers with remote arrives.
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t slot_smem, uint32_t ncols) {   // the SAME warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {     // the same warp of both CTAs
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both]^T ; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_cg2(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t desc_hi32,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(desc_hi32)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` when all UMMAs issued so far by this thread completed
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// arrive on the mbarrier at local offset `bar` of CTA `rank` of this cluster.  RELAXED on purpose: a release at cluster
// scope compiles to MEMBAR + ERRBAR + CCTL.IVALL (it drains the thread's memory operations and invalidates the SM's L1,
// 11 % of all stall samples of the first pair-mode kernel, r02p), and nothing the waiter reads through the generic proxy
// depends on it: operand slabs are published to the tensor core by fence.proxy.async before the arrive, accumulator
// reads are complete (tcgen05.wait::ld) and fenced (tcgen05.fence::before_thread_sync) before it.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// acquire at cluster scope: pairs with remote arrivals (try_wait defaults to CTA scope)
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2, one issue slot for two IEEE fp32 operations) ------------------
// The row threads of the head kernels are the critical path once the UMMAs are cheap (CTA-pair mode, r02p: 147 k warp
// instructions per tile at IPC 1.4); their element-wise work -- bias adds, the layer-1 FMAs, the v - hi subtractions of
// the operand split -- runs on pairs.  Results are bit-identical to the scalar forms (same rounding, no contraction).
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t r, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// ---- hi/lo split of an fp32 value pair ----------------------------------------------------------
// x = hi + lo with hi = fp16(x), lo = fp16(x - hi): the pair carries 22 mantissa bits, and
// A*W ~= Ahi*Whi + Ahi*Wlo + Alo*Whi (3 UMMAs per product, fp32 accumulation in TMEM) is within ~2x of
// fp32's own rounding noise on this head -- measured 10x closer to the reference than the bf16 split
// (16 bits), at the same tensor-core rate.  The price is fp16's range: conversions saturate at +-65504
// (hi + lo then covers +-1.3e5) and resolve 6e-8 absolutely (subnormals), which is ample for normalised
// images, O(1) features and ReLU activations; bf16 (-DCIAOSR_SPLIT_BF16) keeps fp32's exponent range.
#ifdef CIAOSR_SPLIT_BF16
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v1 - h1), "f"(v0 - h0));
}
__device__ __forceinline__ void split2_relu(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  split2(fmaxf(v0, 0.0f), fmaxf(v1, 0.0f), hi, lo);
}
__host__ __device__ inline void split_scalar(float w, split_t& hi, split_t& lo) {
  hi = __float2bfloat16_rn(w);
  lo = __float2bfloat16_rn(w - __bfloat162float(hi));
}
#else
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float d0, d1;
  unpack2(sub2(pack2(v0, v1), pack2(h.x, h.y)), d0, d1);          // one FADD2 for both residuals
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}
// split of relu(v): the ReLU rides on the conversions (cvt ... .relu clamps negative results to +0), which removes one
// FMNMX per element from the row threads (the critical path of the pair-MLP kernel).  hi is rounded
// TOWARD ZERO so that the residual of a positive value is never negative (a .relu on the second conversion would
// otherwise clip it): |v - hi| < ulp_fp16(v) instead of <= ulp/2, i.e. the pair carries >= 21 instead of 22 mantissa
// bits; negative v gives hi = 0, residual v, lo = relu(v) = 0.
__device__ __forceinline__ void split2_relu(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float d0, d1;
  unpack2(sub2(pack2(v0, v1), pack2(h.x, h.y)), d0, d1);
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}
__device__ inline void split_scalar(float w, split_t& hi, split_t& lo) {
  hi = __float2half_rn(fminf(fmaxf(w, -65504.0f), 65504.0f));
  lo = __float2half_rn(fminf(fmaxf(w - __half2float(hi), -65504.0f), 65504.0f));
}
#endif

// Write 32 consecutive K-columns [c0, c0+32) (c0 % 32 == 0, within one 64-wide slab) of
// operand row `row` into the hi and lo slabs (SW128: 16-byte chunk j of a row lives at j ^ (row & 7)).
template <bool RELU = false, class T>
__device__ __forceinline__ void a_store32(uint32_t slab_hi, uint32_t slab_lo, int row, int c0,
                                          const T (&v)[32]) {
  const uint32_t rbase = (uint32_t)row * 128u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (RELU) split2_relu(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1], h[i], l[i]);
      else split2(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1], h[i], l[i]);
    }
    const uint32_t chunk = (uint32_t)(((c0 >> 3) + j) ^ (row & 7)) << 4;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_hi + rbase + chunk), "r"(h[0]),
                 "r"(h[1]), "r"(h[2]), "r"(h[3])
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_lo + rbase + chunk), "r"(l[0]),
                 "r"(l[1]), "r"(l[2]), "r"(l[3])
                 : "memory");
  }
}

// 16 consecutive K-columns [c0, c0+16) (c0 % 16 == 0): two 16-byte chunks per half
template <bool RELU = false, class T>
__device__ __forceinline__ void a_store16(uint32_t slab_hi, uint32_t slab_lo, int row, int c0,
                                          const T (&v)[16]) {
  const uint32_t rbase = (uint32_t)row * 128u;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (RELU) split2_relu(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1], h[i], l[i]);
      else split2(v[j * 8 + 2 * i], v[j * 8 + 2 * i + 1], h[i], l[i]);
    }
    const uint32_t chunk = (uint32_t)(((c0 >> 3) + j) ^ (row & 7)) << 4;
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_hi + rbase + chunk), "r"(h[0]),
                 "r"(h[1]), "r"(h[2]), "r"(h[3])
                 : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab_lo + rbase + chunk), "r"(l[0]),
                 "r"(l[1]), "r"(l[2]), "r"(l[3])
                 : "memory");
  }
}
template <bool RELU = false, class T>
__device__ __forceinline__ void a_storeN(uint32_t hi, uint32_t lo, int row, int c0, const T (&v)[32]) { a_store32<RELU>(hi, lo, row, c0, v); }
template <bool RELU = false, class T>
__device__ __forceinline__ void a_storeN(uint32_t hi, uint32_t lo, int row, int c0, const T (&v)[16]) { a_store16<RELU>(hi, lo, row, c0, v); }

// byte offset of element (n, k) inside a [128 x 64] SW128 slab (used by the weight packer)
__host__ __device__ inline uint32_t sw128_offset(int n, int k) {
  return (uint32_t)n * 128u + (uint32_t)(((k >> 3) ^ (n & 7)) << 4) + (uint32_t)(k & 7) * 2u;
}

}  // namespace tc
}  // namespace ciaosr
