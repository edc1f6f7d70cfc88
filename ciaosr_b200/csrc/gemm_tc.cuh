// Generic tcgen05 GEMM with functor-generated A rows:  C[m, n] = sum_k A(m, k) * B[n, k]
//
//   A  is produced on the fly, one thread per row (implicit im2col / patch extraction / pair
//      products), split into fp16 hi/lo and written into the SW128 operand slabs;
//   B  is a pre-swizzled unit blob ([128 N x 64 K] hi+lo, 32 KB each) in global memory: static
//      weights packed at plan time, or per-image operands packed by a small kernel just before;
//   C  leaves through an epilogue functor that sees 32 consecutive columns of one row.
//
// A job = (128-row tile, N-chunk of <= 256 columns); A is regenerated per job (K may exceed the
// 256 columns that fit the four operand slots, so slabs stream through them under A_FREE).
// Same pipeline, barriers and split arithmetic as the head kernels (tc_pipeline.cuh).
#pragma once
#include <type_traits>
#include "tc_pipeline.cuh"

namespace ciaosr {

struct GemmShape {
  long long M;             // rows
  int kslabs;              // ceil(K / 64)
  int nunits;              // ceil(N / 128): weight units per K-slab over the whole N
  long long rows_per_image;    // B operand switches every this many rows (M if shared)
  size_t blob_image_stride;    // bytes between per-image blobs (0 if shared)
  int kchunk = 0;              // K-slabs per accumulation chunk (0 = all of K in one TMEM accumulation), see below
  int ksplit = 0;              // > 1 (TMA-fed A only): the K range is cut into `ksplit` parts that run as INDEPENDENT jobs
                               // (few-row, long-K problems would otherwise occupy a handful of CTAs); part p stores its
                               // partial product through the epilogue at virtual row  p * (m_tiles * 128) + m,  i.e. the
                               // output is [ksplit][m_tiles * 128, N] and the caller sums the parts.
  int a_resident = 0;          // 1: K fits the four operand slots (kslabs <= 4): A is generated ONCE per 128-row tile and
                               // kept in shared memory while all N-chunks of that tile are issued back to back (the
                               // epilogue of chunk c overlaps the UMMAs of chunk c+1 through the two accumulators).
                               // Short-K layers (the SwinIR trunk's K = 180 / 360 Linears) are otherwise bound by
                               // regenerating A and refilling the pipeline per (tile, chunk) job.
};
// Long-K accumulation.  tcgen05.mma adds into its fp32 TMEM accumulator with TRUNCATION (measured,
// tools/ubench/mma_acc.cu: adding 0.94 ulp to 1.0 leaves 1.0; products inside one K16 instruction keep 2 guard
// bits), i.e. every accumulate step loses up to 1 ulp of the running sum, always towards zero: ~0.5 ulp x 3 K/16
// steps of systematic bias.  Harmless at K = 256 (24 ulp), not at K = 9216 (the P.V product of a 192x192 tile:
// ~900 ulp = 5e-5 relative).  With kchunk set, a job's K range is cut into chunks of `kchunk` slabs; each chunk
// is accumulated in TMEM, and the chunks are summed in fp32 (round-to-nearest) by the epilogue through
// Epi::accumulate (read-modify-write of the thread's own output elements).

// AGen:  struct Row;  __device__ Row row(long long m) const;
//        __device__ void fill(Row&, long long m, int k0, float (&v)[32]) const;     (k0 % 32 == 0)
//        optional: static constexpr bool kCombine = true;  __device__ float& partial(Row&) const;
//                  (a per-row scalar accumulated by fill(); the two threads of a row hold partial sums that are
//                   added through smem before the epilogue sees the row)
// Epi:   __device__ void store(const typename AGen::Row&, long long m, int n0, const float (&v)[32]) const;
//        optional (needed when GemmShape::kchunk is used): accumulate(...) with the same signature: C += v
//
// Threads: 4 control warps + 8 row warps.  Two threads serve each row: thread (half h, lane-quarter q, lane l)
// <-> row 32 q + l, columns [32 h, 32 h + 32) of every 64-column operand slab and the 32-column accumulator
// chunks 2 j + h (warps 4-7 are h = 0, warps 8-11 are h = 1; a warp may only touch TMEM lanes 32 (warp % 4) ..+31).
// Two warps per SM sub-partition overlap one thread's gather latency with the other's conversion work.
constexpr int TC_THREADS = 384;
constexpr int TC_NEPI = 256;

template <class AGen, class = void>
struct agen_combines : std::false_type {};
template <class AGen>
struct agen_combines<AGen, std::enable_if_t<AGen::kCombine>> : std::true_type {};

// AGen::kTma = true: A is not generated but loaded -- it already sits in HBM as two row-major 16-bit matrices
// (hi / lo halves, [M, K]) described by the two tensor maps passed to the launch; the producer warp lands the slabs
// (produce_job_tma_a) and the row threads only drain accumulators.  fill() is never called.
template <class AGen, class = void>
struct agen_tma : std::false_type {};
template <class AGen>
struct agen_tma<AGen, std::enable_if_t<AGen::kTma>> : std::true_type {};
struct TmaRowsGen {          // the generic "A comes through the tensor maps" generator
  static constexpr bool kTma = true;
  struct Row {};
  __device__ __forceinline__ Row row(long long) const { return Row{}; }
  __device__ __forceinline__ void fill(Row&, long long, int, float (&)[32]) const {}
};

// AGen::kPrefetch = true: fill() is split into  issue(Row&, m, k0, Raw&)  -- address arithmetic + the 8 float4 loads of
// a 32-column chunk, nothing that waits on them -- and  finish(Row&, m, k0, const Raw&, v)  -- masking / arithmetic on the
// loaded values.  The row threads then issue slab s+1's loads BEFORE converting and storing slab s, so the L2 / HBM
// latency of a gather overlaps the split + st.shared + fence of the previous slab instead of being exposed once per slab
// (ncu r02f: the short-K Linear kernels spent most of their time in long-scoreboard stalls on exactly these loads).
template <class AGen, class = void>
struct agen_prefetch : std::false_type {};
template <class AGen>
struct agen_prefetch<AGen, std::enable_if_t<AGen::kPrefetch>> : std::true_type {};

// Epi::kTile = true: the functor does not need the per-row state and offers
//   store4(long long m, int n, float4 v)   [and accumulate4(...) when K chunking is used]
// for 4 consecutive columns n .. n+3 of row m.  The kernel then transposes each warp's 32 x 32 accumulator chunk through
// shared memory (XOR-swizzled 16-byte pieces, conflict-free both ways) so that a warp-wide store4 covers 4 rows x 128
// contiguous bytes instead of 32 rows x 16 bytes: the thread-per-row pattern costs 32 memory transactions per
// instruction and made the short-K GEMMs epilogue-bound (r02i: the row warps of the trunk's Linears were busy 83 % of the
// kernel, the UMMA issuer 15 %).
template <class Epi, class = void>
struct epi_tile : std::false_type {};
template <class Epi>
struct epi_tile<Epi, std::enable_if_t<Epi::kTile>> : std::true_type {};
template <class Epi, class = void>
struct epi_tile_accumulates : std::false_type {};
template <class Epi>
struct epi_tile_accumulates<Epi, std::void_t<decltype(std::declval<const Epi&>().accumulate4(0LL, 0, float4{}))>>
    : std::true_type {};

template <class Epi, class Row, class = void>
struct epi_accumulates : std::false_type {};
template <class Epi, class Row>
struct epi_accumulates<Epi, Row, std::void_t<decltype(std::declval<const Epi&>().accumulate(
                                     std::declval<const Row&>(), 0LL, 0, std::declval<const float (&)[32]>()))>>
    : std::true_type {};

template <class AGen, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const GemmShape g, const uint8_t* __restrict__ blob, const AGen agen, const Epi epi,
               const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcShared s = tc_carve_gemm(smem);
  const uint32_t tmem_base = tc_prologue<1, TC_NEPI>(s, smem, GM_SLOT);
  const int warp = threadIdx.x >> 5;
  const int m_tiles = (int)((g.M + ROWS - 1) / ROWS);
  const int n_chunks = (g.nunits + 1) / 2;
  const bool resident = g.a_resident != 0;           // host guarantees: kslabs <= 4, no K chunking, no TMA-fed A
  // outer job = (tile, first chunk): one chunk per job normally, all chunks of the tile in resident mode
  const int cpj = resident ? n_chunks : 1;           // chunks per outer job
  const int ksplit = g.ksplit > 1 ? g.ksplit : 1;
  const int spp = (g.kslabs + ksplit - 1) / ksplit;  // K-slabs per part
  const long long jobs_per_part = resident ? (long long)m_tiles : (long long)m_tiles * n_chunks;
  const long long n_jobs = jobs_per_part * ksplit;
  // K range of a job: slabs [k0, k0 + klen), accumulated in chunks of `kchunk` slabs
  auto k_range = [&](long long job, int& k0, int& klen, int& kch, int& nkc) {
    const int part = (int)(job / jobs_per_part);
    k0 = part * spp;
    klen = max(0, min(spp, g.kslabs - k0));
    kch = (g.kchunk > 0 && g.kchunk < klen) ? g.kchunk : max(klen, 1);
    nkc = (klen + kch - 1) / kch;
  };

  if (warp == 0) {
    ProdState ps{0};
    uint32_t afree_bits = 0xFu;
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const long long jb = job % jobs_per_part;
      const int mt = (int)(resident ? jb : jb / n_chunks), nc0 = (int)(resident ? 0 : jb % n_chunks);
      const long long image = ((long long)mt * ROWS) / g.rows_per_image;
      int k0, klen, kchunk, n_kc;
      k_range(job, k0, klen, kchunk, n_kc);
      for (int nc = nc0; nc < nc0 + cpj; ++nc) {
        const int units = min(2, g.nunits - 2 * nc);
        // blob order: for chunk: for slab: for unit
        const uint8_t* src = blob + image * g.blob_image_stride + (size_t)nc * 2 * g.kslabs * UNIT_BYTES +
                             (size_t)k0 * units * UNIT_BYTES;
        for (int kc = 0; kc < n_kc; ++kc) {
          const int ns = min(kchunk, klen - kc * kchunk);
          if constexpr (agen_tma<AGen>::value)
            produce_job_tma_a<1, TC_NEPI, false>(s, ps, afree_bits, src + (size_t)kc * kchunk * units * UNIT_BYTES, ns,
                                                units, 0, &map_hi, &map_lo, k0 + kc * kchunk, mt * ROWS);
          else
            produce_job<1, false>(s, ps, src + (size_t)kc * kchunk * units * UNIT_BYTES, ns, units, 0);
        }
      }
    }
  } else if (warp == 1) {
    MmaState m{0, 0, 0};
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const long long jb = job % jobs_per_part;
      const int nc0 = (int)(resident ? 0 : jb % n_chunks);
      int k0, klen, kchunk, n_kc;
      k_range(job, k0, klen, kchunk, n_kc);
      for (int nc = nc0; nc < nc0 + cpj; ++nc)
        for (int kc = 0; kc < n_kc; ++kc)
          mma_job<1, false>(s, tmem_base, m, min(kchunk, klen - kc * kchunk), min(2, g.nunits - 2 * nc),
                           !resident || nc == 0, !resident || nc == n_chunks - 1);
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + (threadIdx.x & 31);
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu, 0};
    for (long long job = blockIdx.x; job < n_jobs; job += gridDim.x) {
      const long long jb = job % jobs_per_part;
      const int mt = (int)(resident ? jb : jb / n_chunks), nc0 = (int)(resident ? 0 : jb % n_chunks);
      const long long m = (long long)mt * ROWS + row;
      const bool valid = m < g.M;
      const long long m_out = m + (job / jobs_per_part) * ((long long)m_tiles * ROWS);     // part-stacked output rows
      int k0, klen, kchunk, n_kc;
      k_range(job, k0, klen, kchunk, n_kc);
      typename AGen::Row rs = agen.row(valid ? m : 0);
      for (int nc = nc0; nc < nc0 + cpj; ++nc) {
        const int units = min(2, g.nunits - 2 * nc);
        for (int kc = 0; kc < n_kc; ++kc) {
          const int sl_end = k0 + min((kc + 1) * kchunk, klen);
          if (!resident || nc == 0) {
            if constexpr (agen_prefetch<AGen>::value) {
              // software-pipelined: the loads of slab sl+1 are in flight while slab sl is converted and stored.
              // Two named buffers used alternately (a runtime-indexed array would live in local memory).
              typename AGen::Raw ra, rb;
              const int sl0 = k0 + kc * kchunk;
              auto process = [&](int sl, const typename AGen::Raw& raw) {
                const int slot = (sl - k0) & 3;
                float v[32];
                if (valid) agen.finish(rs, m, sl * 64 + half * 32, raw, v);
                else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                }
                slab_begin(s, e, slot, true);
                a_store32(s.a_hi + slot * SLAB_BYTES, s.a_lo + slot * SLAB_BYTES, row, half * 32, v);
                slab_done(s, slot);
              };
              if (valid && sl0 < sl_end) agen.issue(rs, m, sl0 * 64 + half * 32, ra);
#pragma unroll 1
              for (int sl = sl0; sl < sl_end; sl += 2) {
                if (valid && sl + 1 < sl_end) agen.issue(rs, m, (sl + 1) * 64 + half * 32, rb);
                process(sl, ra);
                if (sl + 1 < sl_end) {
                  if (valid && sl + 2 < sl_end) agen.issue(rs, m, (sl + 2) * 64 + half * 32, ra);
                  process(sl + 1, rb);
                }
              }
            } else {
#pragma unroll 1
              for (int sl = k0 + kc * kchunk; sl < sl_end; ++sl) {
                const int slot = (sl - k0) & 3;
                if constexpr (agen_tma<AGen>::value) { e.afree_bits ^= 1u << slot; continue; }   // producer-loaded slab
                float v[32];
                if (valid) agen.fill(rs, m, sl * 64 + half * 32, v);
                else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = 0.0f;
                }
                slab_begin(s, e, slot, true);
                a_store32(s.a_hi + slot * SLAB_BYTES, s.a_lo + slot * SLAB_BYTES, row, half * 32, v);
                slab_done(s, slot);
              }
            }
          }
          if constexpr (agen_combines<AGen>::value) {
            if (kc == n_kc - 1) {
              s.xchg[half * ROWS + row] = agen.partial(rs);
              epi_sync<TC_NEPI>();
              agen.partial(rs) = s.xchg[row] + s.xchg[ROWS + row];
            }
          }
          const uint32_t d = epi_wait_d(s, e, units);
          if constexpr (epi_tile<Epi>::value) {
            // coalesced epilogue: lane r holds row r's 32 columns -> stage [32 rows][8 x 16 B] with piece j of row r at
            // position j ^ (r & 7) -> lane l reads the 4 columns 4 (l & 7) .. of rows (l >> 3) + 4 i
            const int lane = threadIdx.x & 31;
            const uint32_t stg = smem_u32(smem) + GM_STAGE + (uint32_t)(warp - 4) * 4096u;
            const long long m_w0 = (long long)mt * ROWS + (warp & 3) * 32;              // first row of this warp
            const long long out_off = m_out - m;                                         // part stacking (ksplit)
#pragma unroll 1
            for (int cc = half; cc < units * 4; cc += 2) {
              float v[32];
              tmem_ld32(lane_taddr + d * 256 + cc * 32, v);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 128u +
                                                                               (uint32_t)((j ^ (lane & 7)) << 4)),
                             "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                             : "memory");
              __syncwarp();
              const int n = nc * 256 + cc * 32 + 4 * (lane & 7);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = (lane >> 3) + 4 * i;
                float4 q;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                             : "r"(stg + (uint32_t)rr * 128u + (uint32_t)(((lane & 7) ^ (rr & 7)) << 4))
                             : "memory");
                const long long mr = m_w0 + rr;
                if (mr < g.M) {
                  if (kc == 0) epi.store4(mr + out_off, n, q);
                  else if constexpr (epi_tile_accumulates<Epi>::value) epi.accumulate4(mr + out_off, n, q);
                }
              }
              __syncwarp();
            }
          } else {
#pragma unroll 1
            for (int cc = half; cc < units * 4; cc += 2) {
              float v[32];
              tmem_ld32(lane_taddr + d * 256 + cc * 32, v);
              if (valid) {
                if (kc == 0) epi.store(rs, m_out, nc * 256 + cc * 32, v);
                else if constexpr (epi_accumulates<Epi, typename AGen::Row>::value)
                  epi.accumulate(rs, m_out, nc * 256 + cc * 32, v);
              }
            }
          }
          epi_release_d(s, e);
        }
        if constexpr (agen_combines<AGen>::value) epi_sync<TC_NEPI>();     // xchg is rewritten by the next job
      }
    }
  }
  tc_teardown<1>(tmem_base);
}

inline int tc_grid_size(long long n_jobs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(n_jobs < sms ? n_jobs : sms);
}

template <class AGen, class Epi>
static int tc_gemm(const GemmShape& g, const uint8_t* blob, const AGen& agen, const Epi& epi, cudaStream_t st,
                   const CUtensorMap* map_hi = nullptr, const CUtensorMap* map_lo = nullptr) {
  if (g.M <= 0) return CIAOSR_OK;
  CIAOSR_REQUIRE(!agen_tma<AGen>::value || (map_hi && map_lo), CIAOSR_E_INVALID, "tc_gemm: tensor maps missing");
  static const CUtensorMap no_map = {};
  static DynSmemOptIn optin;         // one per template instantiation, per device (common.cuh)
  if (int rc = optin.ensure(tc_gemm_kernel<AGen, Epi>, GM_TOTAL)) return rc;
  const long long n_jobs = ((g.M + ROWS - 1) / ROWS) * (g.a_resident ? 1 : (g.nunits + 1) / 2) * (g.ksplit > 1 ? g.ksplit : 1);
  CIAOSR_REQUIRE(g.ksplit <= 1 || (agen_tma<AGen>::value && !g.a_resident), CIAOSR_E_INVALID,
                 "tc_gemm: K splitting needs a TMA-fed A operand");
  CIAOSR_REQUIRE(!g.a_resident || (g.kslabs <= 4 && !(g.kchunk > 0 && g.kchunk < g.kslabs) && !agen_tma<AGen>::value &&
                                   !agen_combines<AGen>::value),
                 CIAOSR_E_INVALID, "tc_gemm: A-resident mode needs K <= 256, one accumulation pass and a generated A");
  if (g.kchunk > 0 && g.kchunk < g.kslabs) {
    const bool can = epi_tile<Epi>::value ? epi_tile_accumulates<Epi>::value : epi_accumulates<Epi, typename AGen::Row>::value;
    CIAOSR_REQUIRE(can && g.kchunk % 4 == 0, CIAOSR_E_INVALID,
                   "tc_gemm: K chunking needs an accumulating epilogue and kchunk %% 4 == 0 (operand slots)");
  }
  CIAOSR_LAUNCH((tc_gemm_kernel<AGen, Epi>), tc_grid_size(n_jobs), TC_THREADS, GM_TOTAL, st, g, blob, agen, epi,
                map_hi ? *map_hi : no_map, map_lo ? *map_lo : no_map);
  return CIAOSR_OK;
}

// Pack a [N, K] operand into the unit blob: element (n, k) = src(n, k) (0 outside N x K).
// Blob order: chunk (256 N) -> slab (64 K) -> unit (128 N), each unit = [hi slab][lo slab].
template <class Src>
__global__ void tc_pack_operand_kernel(uint8_t* __restrict__ dst, int N, int K, int kslabs, int nunits,
                                       size_t image_stride, const Src src) {
  const long long per_image = (long long)kslabs * nunits * UNIT_N * KSLAB;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= per_image) return;
  const int image = blockIdx.y;
  const int k_in = (int)(i % KSLAB);
  const int n_in = (int)((i / KSLAB) % UNIT_N);
  const int rest = (int)(i / (KSLAB * UNIT_N));      // enumerates (unit_global, slab) pairs
  const int ug = rest % nunits, sl = rest / nunits;
  const int n = ug * UNIT_N + n_in, k = sl * KSLAB + k_in;
  const float w = (n < N && k < K) ? src(image, n, k) : 0.0f;
  split_t hi, lo;
  split_scalar(w, hi, lo);
  const int nc = ug / 2, u = ug % 2;
  const int units = min(2, nunits - 2 * nc);
  const size_t unit_index = (size_t)nc * 2 * kslabs + (size_t)sl * units + u;
  uint8_t* ub = dst + image * image_stride + unit_index * UNIT_BYTES;
  const uint32_t off = sw128_offset(n_in, k_in);
  *reinterpret_cast<split_t*>(ub + off) = hi;
  *reinterpret_cast<split_t*>(ub + SLAB_BYTES + off) = lo;
}

inline size_t tc_operand_blob_bytes(int kslabs, int nunits) { return (size_t)kslabs * nunits * UNIT_BYTES; }

template <class Src>
static int tc_pack_operand(uint8_t* dst, int images, int N, int K, size_t image_stride, const Src& src,
                           cudaStream_t st) {
  const int kslabs = (K + KSLAB - 1) / KSLAB, nunits = (N + UNIT_N - 1) / UNIT_N;
  const long long per_image = (long long)kslabs * nunits * UNIT_N * KSLAB;
  dim3 grid(cdiv(per_image, 256), images);
  CIAOSR_LAUNCH((tc_pack_operand_kernel<Src>), grid, 256, 0, st, dst, N, K, kslabs, nunits, image_stride, src);
  return CIAOSR_OK;
}

}  // namespace ciaosr
