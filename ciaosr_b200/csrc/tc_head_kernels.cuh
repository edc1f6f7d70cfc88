// The two fused tcgen05 kernels of the head (see head_tc.cu for the engine overview).
//
// Thread layout of both kernels: 384 threads = 4 control warps (0 weight producer, 1 UMMA issuer,
// 2 TMEM allocator, 3 idle) + 8 row warps.  Two threads serve each of the 128 rows of a tile:
// thread (half h, lane-quarter q, lane l) <-> row 32 q + l, columns [32 h, 32 h + 32) of every 64-column
// slab (warps 4-7 are h = 0, warps 8-11 are h = 1; a warp may only touch TMEM lanes 32 (warp % 4) ..+31).
// Two warps per SM sub-partition plus register prefetch (next tcgen05.ld / next gather issued before the
// current chunk is converted) hide the TMEM, L2 and shared-memory latencies behind each other.
#pragma once
#include "gemm_tc.cuh"
#include "pairs.cuh"
#include "tma.cuh"

namespace ciaosr {

constexpr int HEAD_THREADS = 384;
constexpr int NEPI = 256;

// =====================================================================================================
// pair kernel
// =====================================================================================================
struct PairParams {
  PairConsts pc;
  const float* coord; const float* cell;
  const float* featT; const float* nlT;
  int C, Cn, Dv, Dvp;
  const float* Pk; const float* Pv; const float* G; int ldg;
  const float* consts;        // 16 x 256 floats + bv5p[Dvp], see layout below
  const uint8_t* blob; int units_per_tile; int units5;
  split_t* x_hi; split_t* x_lo;   // [total_q, Dvp] attended values, already split for the query kernel's UMMAs
  long long total_rows; int n_tiles; int iters;     // iters = tiles per CTA (uniform over the grid)
  float softmax_scale;
  uint32_t terms;             // product terms of the UMMA jobs (TcShared::terms); 7 unless CIAOSR_TC_TERMS says otherwise
};
// consts layout (x256 floats): 0..3 rc_k, 4 b1_k, 5..7 b_k(layers 2..4), 8..11 rc_v, 12 b1_v, 13..15 b_v(2..4),
//                               then bv5p[Dvp] (last value Linear bias, tap-major, zero padded)

// layer 1 of imnet_k / imnet_v from the LR hoist:  relu(P[pix] + b1 + rc . [rel_y, rel_x, sc_y, sc_x])
// PARTS = row threads per row (2 or 4); `half` = this thread's part: columns [CW*half, CW*half + CW) of every slab
// WAITK (the value chain's layer 1): the operand slots still hold k.L4's input until its UMMAs are complete, so the
// gather and the arithmetic of slab 0 run first and the wait for k.L4's accumulator sits right before the first store:
// the issuer sees slab 0 ~2 k cycles earlier than with the wait in front.  Returns k.L4's accumulator index.
// WAITK = 2 (the NEXT tile's key layer 1, written before this tile's last value chunk is reduced): same, the wait is for
// the last accumulator of this tile (`wait_halves` column halves).
struct L1Result { EpiState e; uint32_t dk; };      // pipeline state by value, see epi_hidden
template <int PARTS, int WAITK>
__device__ __noinline__ L1Result gen_layer1(const TcShared s, EpiState e, int row, int half, const PairInfo p,
                                               const float* __restrict__ P, const float* __restrict__ rc_s,
                                               const float* __restrict__ b1_s, int wait_halves = 2) {
  uint32_t dk = 0;
  constexpr int CW = 64 / PARTS, NV = CW / 4;
  const float4* prow = p.pix >= 0 ? reinterpret_cast<const float4*>(P + (long long)p.pix * HID) + half * NV : nullptr;
  // the gather of this row's four slabs runs TWO slabs ahead of the arithmetic (r03s trace: with one slab of lookahead every
  // slab took 2.5 k cycles against 1.1 k for the same conversion work in epi_hidden -- L2 latency under the weight streams)
  float4 buf[2][NV];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int j = 0; j < NV; ++j) buf[q][j] = prow ? __ldg(prow + q * 16 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int sl = 0; sl < 4; ++sl) {
    float v[CW];
#pragma unroll
    for (int j = 0; j < NV; ++j) { v[4 * j] = buf[sl & 1][j].x; v[4 * j + 1] = buf[sl & 1][j].y; v[4 * j + 2] = buf[sl & 1][j].z; v[4 * j + 3] = buf[sl & 1][j].w; }
    if (sl + 2 < 4) {
#pragma unroll
      for (int j = 0; j < NV; ++j) buf[sl & 1][j] = prow ? __ldg(prow + (sl + 2) * 16 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int c0 = sl * 64 + half * CW;
    // same operations in the same order as the scalar form, two columns per instruction (FADD2 / FFMA2)
    const uint64_t ry = pack2(p.rel_y, p.rel_y), rx = pack2(p.rel_x, p.rel_x);
    const uint64_t sy = pack2(p.sc_y, p.sc_y), sx = pack2(p.sc_x, p.sc_x);
#pragma unroll
    for (int i = 0; i < CW; i += 4) {                  // four columns per step: one LDS.128 per constant array
      const int c = c0 + i;
      const float4 b1 = *reinterpret_cast<const float4*>(b1_s + c);
      const float4 r0 = *reinterpret_cast<const float4*>(rc_s + c), r1 = *reinterpret_cast<const float4*>(rc_s + HID + c);
      const float4 r2 = *reinterpret_cast<const float4*>(rc_s + 2 * HID + c), r3 = *reinterpret_cast<const float4*>(rc_s + 3 * HID + c);
      uint64_t t = add2(pack2(v[i], v[i + 1]), pack2(b1.x, b1.y));
      t = fma2(pack2(r0.x, r0.y), ry, t);
      t = fma2(pack2(r1.x, r1.y), rx, t);
      t = fma2(pack2(r2.x, r2.y), sy, t);
      t = fma2(pack2(r3.x, r3.y), sx, t);
      unpack2(t, v[i], v[i + 1]);                      // the ReLU is part of the operand split (split2_relu)
      uint64_t u = add2(pack2(v[i + 2], v[i + 3]), pack2(b1.z, b1.w));
      u = fma2(pack2(r0.z, r0.w), ry, u);
      u = fma2(pack2(r1.z, r1.w), rx, u);
      u = fma2(pack2(r2.z, r2.w), sy, u);
      u = fma2(pack2(r3.z, r3.w), sx, u);
      unpack2(u, v[i + 2], v[i + 3]);
    }
    if (WAITK && sl == 0) {
      dk = epi_wait_half(s, e, 0);
      if (wait_halves == 2) epi_wait_half(s, e, 1);
      if (WAITK == 1 && threadIdx.x == EPI_T0) TC_TRACE(3002);       // k.L4 complete
    }
    slab_begin(s, e, sl, false);
    if (!TC_DBG(1)) a_storeN<true>(s.a_hi + sl * SLAB_BYTES, s.a_lo + sl * SLAB_BYTES, row, half * CW, v);
    if (s.pair_rank < 0 || sl < 2) slab_done(s, sl);     // CTA-pair mode: slabs 2 + 3 share a fence (see epi_hidden)
    else if (sl & 1) slabs_done2(s, sl - 1, sl);
  }
  return L1Result{e, dk};
}

// Row-thread work of ONE pair tile (128 (query, neighbour) rows = 32 queries): layer 1 of both chains from the LR hoists,
// the hidden-layer epilogues, logits + softmax over the 4 neighbours, the value gather and the weighted sum -> x rows
// [x_row0, x_row0 + 32) of P.x_hi / P.x_lo.  `cst` = the pair constants in shared memory, `bv5` = the padded last bias.
// PARTS = row threads per row (2: 32-column chunks, 4: 16-column chunks); `half` = this thread's part index.
// Software pipeline across tiles [r03]: `p` is this tile's pair geometry, computed during the PREVIOUS tile (compute_pair is
// ~3.6 k cycles of dependent global loads and divisions); on return it holds the geometry of `next_tile` (< 0: none),
// computed here while the value chain's UMMAs run.  Layer 1 of the next tile's key chain is written as soon as the last
// UMMA of this tile is complete -- BEFORE the last value chunk is reduced -- so the tensor pipe works on the next tile
// while this tile's x rows are finished (r02w trace: 12.7 k of every 104 k-cycle tile period had it idle at the boundary).
__device__ __forceinline__ PairInfo pair_info_of(const PairParams& P, long long tile, int row) {
  const long long R = tile * ROWS + row;
  PairInfo p;
  if (tile >= 0 && tile < P.n_tiles && R < P.total_rows) p = compute_pair(P.pc, P.coord, P.cell, R >> 2, (int)(R & 3));
  else { p.pix = -1; p.gidx = -1; p.rel_y = p.rel_x = p.sc_y = p.sc_x = 0.0f; }
  return p;
}
template <int PARTS>
__device__ __forceinline__ void pair_tile_rows(const TcShared& s, EpiState& e, const PairParams& P, const float* cst,
                                               const float* bv5, uint32_t lane_taddr, int row, int half, int lane,
                                               long long tile, long long x_row0, PairInfo& p, bool first, long long next_tile) {
  const int C = P.C, H = P.pc.H, W = P.pc.W;
  constexpr int CW = 64 / PARTS, NV = CW / 4;    // columns / float4s per chunk
  const bool tap32 = (C % CW) == 0;              // a chunk never straddles a tap
  const int nchunks5 = (P.units5 + 1) / 2;
  {
      const long long R = tile * ROWS + row;
      const bool valid = tile < P.n_tiles && R < P.total_rows;

      // ---- key chain -----------------------------------------------------------------------
      if (threadIdx.x == EPI_T0) { TC_TRACE(3000); TC_TRACE_NS(9000); }       // tile start (+ wall clock in the trace build)
      if (first) e = gen_layer1<PARTS, 0>(s, e, row, half, p, P.Pk, cst, cst + 4 * HID).e;   // else: written by the previous tile
      if (threadIdx.x == EPI_T0) TC_TRACE(3001);       // k.L1 written
      e = epi_hidden<false, PARTS>(s, e, lane_taddr, row, half, cst + 5 * HID);
      e = epi_hidden<false, PARTS>(s, e, lane_taddr, row, half, cst + 6 * HID);
      // k.L4 is complete once both accumulator halves are; that also frees the operand slabs, so the value
      // chain's layer 1 is built FIRST: the UMMAs of v.L2 then run while the logits are reduced from D.
      const L1Result l1v = gen_layer1<PARTS, 1>(s, e, row, half, p, P.Pv, cst + 8 * HID, cst + 12 * HID);
      e = l1v.e;
      const uint32_t dk = l1v.dk;
      if (threadIdx.x == EPI_T0) TC_TRACE(3003);       // v.L1 written
      float logit = 0.0f;
      {
        const uint32_t d = dk;
        const float* bias_s = cst + 7 * HID;
        const float4* grow = p.gidx >= 0 ? reinterpret_cast<const float4*>(P.G + (long long)p.gidx * P.ldg) + half * NV : nullptr;
        uint32_t buf[2][CW];
        float4 gb[NV];
        tmem_ldN_issue(lane_taddr + d * 256 + half * CW, buf[0]);
#pragma unroll
        for (int j = 0; j < NV; ++j) gb[j] = grow ? __ldg(grow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          tmem_ldN_wait(buf[sl & 1]);
          if (sl + 1 < 4) tmem_ldN_issue(lane_taddr + d * 256 + (sl + 1) * 64 + half * CW, buf[(sl + 1) & 1]);
          const int c0 = sl * 64 + half * CW;
#pragma unroll
          for (int j = 0; j < NV; ++j) {               // bias adds on pairs; the logit sum keeps its sequential order
            const float4 g = gb[j];
            const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * j);
            float h0, h1, h2, h3;
            unpack2(add2(pack2(__uint_as_float(buf[sl & 1][4 * j]), __uint_as_float(buf[sl & 1][4 * j + 1])), pack2(b4.x, b4.y)), h0, h1);
            unpack2(add2(pack2(__uint_as_float(buf[sl & 1][4 * j + 2]), __uint_as_float(buf[sl & 1][4 * j + 3])), pack2(b4.z, b4.w)), h2, h3);
            logit = fmaf(fmaxf(h0, 0.0f), g.x, logit);
            logit = fmaf(fmaxf(h1, 0.0f), g.y, logit);
            logit = fmaf(fmaxf(h2, 0.0f), g.z, logit);
            logit = fmaf(fmaxf(h3, 0.0f), g.w, logit);
          }
          if (sl + 1 < 4) {
#pragma unroll
            for (int j = 0; j < NV; ++j) gb[j] = grow ? __ldg(grow + (sl + 1) * 16 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (grow && half == 0) logit += __ldg(P.G + (long long)p.gidx * P.ldg + HID);
        epi_release_d(s, e);
      }
      // combine the two column halves of every row, then softmax over the 4 neighbours (4 adjacent lanes)
      float a;
      {
        s.xchg[half * ROWS + row] = logit;
        epi_sync<128 * PARTS>();
        const float lsum = PARTS == 2 ? s.xchg[row] + s.xchg[ROWS + row]
                                      : (s.xchg[row] + s.xchg[ROWS + row]) + (s.xchg[2 * ROWS + row] + s.xchg[3 * ROWS + row]);
        const float l = __fdiv_rn(lsum, P.softmax_scale);
        float mx = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float ex = expf(l - mx);
        float sum = ex + __shfl_xor_sync(0xffffffffu, ex, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        a = valid ? __fdiv_rn(ex, sum) : 0.0f;
      }

      if (threadIdx.x == EPI_T0) TC_TRACE(3004);       // softmax done
      // everything below only needs pix of this tile's geometry; the next tile's is computed under the value chain's UMMAs
      const int pix = p.pix;
      p = pair_info_of(P, next_tile, row);
      // ---- value chain (layer 1 was built above) ----------------------------------------------------
      e = epi_hidden<false, PARTS>(s, e, lane_taddr, row, half, cst + 13 * HID);
      e = epi_hidden<false, PARTS>(s, e, lane_taddr, row, half, cst + 14 * HID);
      e = epi_hidden<false, PARTS>(s, e, lane_taddr, row, half, cst + 15 * HID);   // h4v -> operand of the last Linear

      // geometry of this row's latent code for the value gather
      int py = 0, px = 0;
      const float* fbase = nullptr;
      const float* nbase = nullptr;
      if (pix >= 0) {
        const int hw = pix % (H * W);
        py = hw / W; px = hw % W;
        fbase = P.featT + (long long)pix * C;
        if (P.nlT) nbase = P.nlT + (long long)pix * P.Cn;
      }
      // value[cp .. cp+3] in tap-major order (cp % 4 == 0); zeros outside the image / past Dv
      auto value4 = [&](int cp) -> float4 {
        if (fbase == nullptr || cp >= P.Dv) return make_float4(0.f, 0.f, 0.f, 0.f);
        if (cp >= 9 * C) return __ldg(reinterpret_cast<const float4*>(nbase + (cp - 9 * C)));
        const int t = cp / C, ch = cp - t * C;
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        if (py + dy < 0 || py + dy >= H || px + dx < 0 || px + dx >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return __ldg(reinterpret_cast<const float4*>(fbase + (dy * W + dx) * C + ch));
      };
      // NV float4 = the CW values of chunk [cp0, cp0+CW)
      auto load_values = [&](int cp0, float4 (&vb)[NV]) {
        if (tap32) {
          const float* src = nullptr;
          if (fbase != nullptr && cp0 < P.Dv) {
            if (cp0 >= 9 * C) src = nbase + (cp0 - 9 * C);
            else {
              const int t = cp0 / C, ch = cp0 - t * C;
              const int dy = t / 3 - 1, dx = t % 3 - 1;
              if (py + dy >= 0 && py + dy < H && px + dx >= 0 && px + dx < W) src = fbase + (dy * W + dx) * C + ch;
            }
          }
#pragma unroll
          for (int g = 0; g < NV; ++g)
            vb[g] = src ? __ldg(reinterpret_cast<const float4*>(src) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
#pragma unroll
          for (int g = 0; g < NV; ++g) vb[g] = value4(cp0 + 4 * g);
        }
      };
      const long long q = x_row0 + (row >> 2);    // row of x for this (query, neighbour) row's query
      const int sub = (lane & 1) * (CW / 2) + ((lane >> 1) & 1) * (CW / 4);   // columns of a chunk this lane ends up owning
      for (int c = 0; c < nchunks5; ++c) {
        const int units = min(2, P.units5 - 2 * c);
        const int nsub = units * 2;                            // chunks per thread (interleaved with the other parts')
        uint32_t d = 0;
        // last chunk: once its accumulator is complete every UMMA of this tile is, and the operand slots are free:
        // the next tile's k.L1 goes in first, then this chunk is reduced while the tensor pipe runs the next tile's k.L2
        const bool ahead = c == nchunks5 - 1 && next_tile >= 0;
        if (ahead) {
          const L1Result l1n = gen_layer1<PARTS, 2>(s, e, row, half, p, P.Pk, cst, cst + 4 * HID, units);
          e = l1n.e;
          d = l1n.dk;
        }
        uint32_t buf[2][CW];
        float4 vb[NV];
        load_values(c * 256 + half * CW, vb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k >= nsub) break;                                // nsub is 2 or 4, warp-uniform
          const int cc = PARTS * k + half;
          const int cp0 = c * 256 + cc * CW;
          float v[CW];
          if ((k & 1) == 0) {                                  // first chunk of an accumulator half
            if (!ahead) d = epi_wait_half(s, e, k >> 1);
            tmem_ldN_issue(lane_taddr + d * 256 + cc * CW, buf[k & 1]);
          }
          tmem_ldN_wait(buf[k & 1]);
          if ((k & 1) == 0) tmem_ldN_issue(lane_taddr + d * 256 + (cc + PARTS) * CW, buf[(k + 1) & 1]);
          const uint64_t aa = pack2(a, a);
#pragma unroll
          for (int g = 0; g < NV; ++g) {               // (a * value) * (acc + bias), two columns per instruction
            const float4 val = vb[g];
            const int o = cp0 + 4 * g;
            const float4 b4 = *reinterpret_cast<const float4*>(bv5 + o);
            unpack2(mul2(mul2(aa, pack2(val.x, val.y)),
                         add2(pack2(__uint_as_float(buf[k & 1][4 * g]), __uint_as_float(buf[k & 1][4 * g + 1])),
                              pack2(b4.x, b4.y))),
                    v[4 * g], v[4 * g + 1]);
            unpack2(mul2(mul2(aa, pack2(val.z, val.w)),
                         add2(pack2(__uint_as_float(buf[k & 1][4 * g + 2]), __uint_as_float(buf[k & 1][4 * g + 3])),
                              pack2(b4.z, b4.w))),
                    v[4 * g + 2], v[4 * g + 3]);
          }
          if (k + 1 < nsub) load_values(cp0 + 64, vb);
          // sum over the 4 neighbour rows (lanes 4q..4q+3), leaving each lane with CW/4 of the chunk's columns
          float r16[CW / 2], r8[CW / 4];
#pragma unroll
          for (int i = 0; i < CW / 2; ++i) {
            const float keep = (lane & 1) ? v[CW / 2 + i] : v[i];
            const float send = (lane & 1) ? v[i] : v[CW / 2 + i];
            r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            const float keep = (lane & 2) ? r16[CW / 4 + i] : r16[i];
            const float send = (lane & 2) ? r16[i] : r16[CW / 4 + i];
            r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          if (valid) {                                         // one 16-byte (8-byte when PARTS = 4) store per half
            if constexpr (PARTS == 2) {
              uint4 h, l;
              split2(r8[0], r8[1], h.x, l.x); split2(r8[2], r8[3], h.y, l.y);
              split2(r8[4], r8[5], h.z, l.z); split2(r8[6], r8[7], h.w, l.w);
              *reinterpret_cast<uint4*>(P.x_hi + q * P.Dvp + cp0 + sub) = h;
              *reinterpret_cast<uint4*>(P.x_lo + q * P.Dvp + cp0 + sub) = l;
            } else {
              uint2 h, l;
              split2(r8[0], r8[1], h.x, l.x); split2(r8[2], r8[3], h.y, l.y);
              *reinterpret_cast<uint2*>(P.x_hi + q * P.Dvp + cp0 + sub) = h;
              *reinterpret_cast<uint2*>(P.x_lo + q * P.Dvp + cp0 + sub) = l;
            }
          }
        }
        epi_release_d(s, e);
      }
  }
}

template <int CL>
__global__ void __launch_bounds__(HEAD_THREADS, 1) pair_mlp_kernel(const PairParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s = tc_carve(smem);
  s.terms = P.terms;
  for (int i = threadIdx.x; i < 16 * HID + P.Dvp; i += HEAD_THREADS) s.consts[i] = P.consts[i];
#ifdef CIAOSR_TC_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) tc::g_trace_on = tc::g_trace_req;
#endif
  const uint32_t tmem_base = tc_prologue<CL, NEPI>(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks5 = (P.units5 + 1) / 2;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  // tile of iteration i: clusters take CL consecutive tiles; every CTA runs `iters` iterations
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;

  if (warp == 0) {
    ProdState ps{0};
#define PAIR_PRODUCE(b, ns, un) produce_job<CL>(s, ps, b, ns, un, cta_rank)
    for (int it = 0; it < P.iters; ++it) {
      const uint8_t* b = P.blob;
      for (int j = 0; j < 6; ++j) { PAIR_PRODUCE(b, 4, 2); b += (size_t)8 * UNIT_BYTES; }
      for (int c = 0; c < nchunks5; ++c) {
        const int units = min(2, P.units5 - 2 * c);
        PAIR_PRODUCE(b, 4, units);
        b += (size_t)4 * units * UNIT_BYTES;
      }
    }
  } else if (warp == 1) {
    MmaState m{0, 0, 0};
// a_release = false: the row threads of this kernel never wait on A_FREE (every operand slot is rewritten only after the
// accumulator that read it was observed complete), so it is not committed either (synccheck: arrivals nobody waits for)
#define PAIR_MMA(ns, un, an) mma_job<CL>(s, tmem_base, m, ns, un, an, false)
    for (int it = 0; it < P.iters; ++it) {
      for (int j = 0; j < 6; ++j) PAIR_MMA(4, 2, true);
      for (int c = 0; c < nchunks5; ++c) PAIR_MMA(4, min(2, P.units5 - 2 * c), c == 0);
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0, 0};
    const float* cst = s.consts;
    const float* bv5 = s.consts + 16 * HID;
    auto tile_of = [&](int it) { return ((long long)it * n_clusters + cluster_id) * CL + cta_rank; };
    PairInfo pinfo = pair_info_of(P, tile_of(0), row);
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = tile_of(it);
      pair_tile_rows<2>(s, e, P, cst, bv5, lane_taddr, row, half, lane, tile, tile * (ROWS / 4), pinfo, it == 0,
                        it + 1 < P.iters ? tile_of(it + 1) : -1);
    }
  }
#ifdef CIAOSR_TC_TIMING
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) tc::g_trace_on = 0;
#endif
  tc_teardown<CL>(tmem_base);
}

// CTA-pair variant (cta_group::2, tc_pipeline.cuh "CTA-pair mode"): the two CTAs of a cluster work on 256 consecutive rows
// with ONE M = 256 UMMA stream issued by the leader; each CTA stages half of every weight operand.  Row-thread work is
// the same function as above (its barrier arrivals go to the leader through TcShared::pair_rank).
// PARTS = row threads per row: 2 (384 threads, default) or 4 (640 threads: 16 row warps, four per SM sub-partition; an
// experiment -- measured no faster, see tc_row_parts() in head_tc_engine.cu).
// NSPLIT: the N-split issue schedule of tc_pipeline.cuh (two N = 128 column halves per layer, epilogue of the first
// half under the UMMAs of the second) instead of one N = 256 stream per layer.
template <int PARTS, bool NSPLIT>
__global__ void __launch_bounds__(128 + 128 * PARTS, 1)
pair_mlp_pair_kernel(const PairParams P, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s = tc_carve(smem);
  s.pair_rank = (int)cluster_ctarank();
  s.terms = P.terms;
  for (int i = threadIdx.x; i < 16 * HID + P.Dvp; i += 128 + 128 * PARTS) s.consts[i] = P.consts[i];
#ifdef CIAOSR_TC_TIMING
  if (blockIdx.x == 0 && threadIdx.x == 0) tc::g_trace_on = tc::g_trace_req;
#endif
  const uint32_t tmem_base = tc_prologue_pair<128 * PARTS>(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks5 = (P.units5 + 1) / 2;
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;

  if (warp == 0) {
    ProdState ps{0};
    if (lane == 0) tma_prefetch_desc(&wmap);
    for (int it = 0; it < P.iters; ++it) {
      long long r = 0;                                   // 128-byte row of the blob
      for (int j = 0; j < 6; ++j) {
        if (NSPLIT) produce_job_pair_split(s, ps, &wmap, r, 2);
        else produce_job_pair(s, ps, &wmap, r, 4, 2);
        r += 4 * 2 * 2 * ROWS;
      }
      for (int c = 0; c < nchunks5; ++c) {
        const int units = min(2, P.units5 - 2 * c);
        if (NSPLIT) produce_job_pair_split(s, ps, &wmap, r, units);
        else produce_job_pair(s, ps, &wmap, r, 4, units);
        r += 4 * units * 2 * ROWS;
      }
    }
  } else if (warp == 1) {
    if (s.pair_rank == 0) {
      MmaState m{0, 0, 0};
      for (int it = 0; it < P.iters; ++it) {
        for (int j = 0; j < 6; ++j) {
          if (NSPLIT) mma_job_pair_split(s, tmem_base, m, 2, true);
          else mma_job_pair(s, tmem_base, m, 4, 2, true, false);          // a_release = false: see PAIR_MMA above
        }
        for (int c = 0; c < nchunks5; ++c) {
          if (NSPLIT) mma_job_pair_split(s, tmem_base, m, min(2, P.units5 - 2 * c), c == 0);
          else mma_job_pair(s, tmem_base, m, 4, min(2, P.units5 - 2 * c), c == 0, false);
        }
      }
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0, 0};
    const float* cst = s.consts;
    const float* bv5 = s.consts + 16 * HID;
    auto tile_of = [&](int it) { return ((long long)it * n_clusters + cluster_id) * 2 + s.pair_rank; };
    PairInfo pinfo = pair_info_of(P, tile_of(0), row);
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = tile_of(it);
      pair_tile_rows<PARTS>(s, e, P, cst, bv5, lane_taddr, row, half, lane, tile, tile * (ROWS / 4), pinfo, it == 0,
                            it + 1 < P.iters ? tile_of(it + 1) : -1);
    }
  }
#ifdef CIAOSR_TC_TIMING
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) tc::g_trace_on = 0;
#endif
  tc_teardown_pair(tmem_base);
}

// =====================================================================================================
// query kernel: imnet_q + residual
// =====================================================================================================
struct QueryParams {
  int Dvp;                                 // x = x_hi + x_lo [total_q, Dvp] comes through the two tensor maps
  const float* consts;                     // x256 floats: 0..3 b_q(layers 1..4), 4..6 W5 rows, 7: b5 in [0..3)
  const uint8_t* blob; int units_per_tile; int slabs1;
  const float* lr; const float* coord; int H, W, Q;
  float* out; long long total_q; int n_tiles; int iters;
  uint32_t terms;
};

// Row-thread work of ONE query tile (128 queries): drain layer 1 (its operand slabs are TMA-loaded by the producer warp),
// hidden layers 2..4, the 256 -> 3 Linear and the bilinear residual on CUDA cores -> P.out rows of tile `tile`.
__device__ __forceinline__ void query_tile_rows(const TcShared& s, EpiState& e, const QueryParams& P, const float* cst,
                                                uint32_t lane_taddr, int row, int half, long long tile) {
  {
      const long long g = tile * ROWS + row;
      const bool valid = tile < P.n_tiles && g < P.total_q;
      // layer-1 operand slabs are written by the producer warp (TMA); keep the slot parity bookkeeping in step
      for (int sl = 0; sl < P.slabs1; ++sl) e.afree_bits ^= 1u << (sl & 3);
      e = epi_hidden<true, 2>(s, e, lane_taddr, row, half, cst);
      e = epi_hidden<true, 2>(s, e, lane_taddr, row, half, cst + HID);
      e = epi_hidden<true, 2>(s, e, lane_taddr, row, half, cst + 2 * HID);
      // last hidden layer + the 256 -> 3 Linear on CUDA cores (each thread: its 32 columns of every slab)
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      {
        uint32_t d = 0;
        const float* bias_s = cst + 3 * HID;
        uint32_t buf[2][32];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          if ((sl & 1) == 0) {
            d = epi_wait_half(s, e, sl >> 1);
            tmem_ld32_issue(lane_taddr + d * 256 + sl * 64 + half * 32, buf[sl & 1]);
          }
          tmem_ld32_wait(buf[sl & 1]);
          if ((sl & 1) == 0) tmem_ld32_issue(lane_taddr + d * 256 + (sl + 1) * 64 + half * 32, buf[(sl + 1) & 1]);
          const int c0 = sl * 64 + half * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float h = fmaxf(__uint_as_float(buf[sl & 1][i]) + bias_s[c0 + i], 0.0f);
            o0 = fmaf(h, cst[4 * HID + c0 + i], o0);
            o1 = fmaf(h, cst[5 * HID + c0 + i], o1);
            o2 = fmaf(h, cst[6 * HID + c0 + i], o2);
          }
        }
        epi_release_d(s, e);
      }
      if (half == 1) { s.xchg[row * 3] = o0; s.xchg[row * 3 + 1] = o1; s.xchg[row * 3 + 2] = o2; }
      epi_sync<NEPI>();
      if (half == 0 && valid) {
        o0 += s.xchg[row * 3] + cst[7 * HID];
        o1 += s.xchg[row * 3 + 1] + cst[7 * HID + 1];
        o2 += s.xchg[row * 3 + 2] + cst[7 * HID + 2];
        if (P.lr) {
          const int b = (int)(g / P.Q);
          const float cy = P.coord[g * 2], cx = P.coord[g * 2 + 1];
          const float* img = P.lr + (long long)b * 3 * P.H * P.W;
          o0 += bilinear_border(img, P.H, P.W, cy, cx);
          o1 += bilinear_border(img + P.H * P.W, P.H, P.W, cy, cx);
          o2 += bilinear_border(img + 2 * P.H * P.W, P.H, P.W, cy, cx);
        }
        P.out[g * 3] = o0; P.out[g * 3 + 1] = o1; P.out[g * 3 + 2] = o2;
      }
      epi_sync<NEPI>();        // xchg is rewritten next tile
  }
}

// Layer 1's A operand is x, which pair_mlp_kernel left in HBM already split into fp16 hi / lo: the producer
// warp lands each [128 queries x 64 columns] slab pair straight in the operand slots with two 2-D TMA tile
// loads (128B swizzle = the UMMA layout; rows past total_q read as zero), so the row threads touch layer 1
// only to drain its accumulator.  A_READY counts NEPI/32 arrivals for the row-warp-written layers; for a TMA
// slab the producer supplies all of them itself (one with the transaction byte count).
template <int CL>
__global__ void __launch_bounds__(HEAD_THREADS, 1)
query_mlp_kernel(const QueryParams P, const __grid_constant__ CUtensorMap map_hi,
                 const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s = tc_carve(smem);
  s.terms = P.terms;
  for (int i = threadIdx.x; i < 8 * HID; i += HEAD_THREADS) s.consts[i] = P.consts[i];
  const uint32_t tmem_base = tc_prologue<CL, NEPI>(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;

  if (warp == 0) {
    ProdState ps{0};
    uint32_t afree_bits = 0xFu;            // parity to wait on next, per operand slot (same bookkeeping as the row threads)
    if (lane == 0) { tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo); }
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = ((long long)it * n_clusters + cluster_id) * CL + cta_rank;
      const int row0 = (int)min(tile * ROWS, (long long)0x7FFFFF00);     // past the end: all rows read as zero
      const uint8_t* b = P.blob;
      // layer 1: A slabs (TMA) interleaved with their weight slabs, so neither ring starves the other
      produce_job_tma_a<CL, NEPI>(s, ps, afree_bits, b, P.slabs1, 2, cta_rank, &map_hi, &map_lo, 0, row0);
      b += (size_t)P.slabs1 * 2 * UNIT_BYTES;
      for (int j = 0; j < 3; ++j) { produce_job<CL>(s, ps, b, 4, 2, cta_rank); b += (size_t)8 * UNIT_BYTES; }
      afree_bits ^= 0xFu; afree_bits ^= 0xFu; afree_bits ^= 0xFu;        // layers 2..4: each slot written once by the row threads
    }
  } else if (warp == 1) {
    MmaState m{0, 0, 0};
    for (int it = 0; it < P.iters; ++it) {
      mma_job<CL>(s, tmem_base, m, P.slabs1, 2, true);
      for (int j = 0; j < 3; ++j) mma_job<CL>(s, tmem_base, m, 4, 2, true);
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu, 0};                  // A_free waits start at parity 1 (fresh barrier passes)
    const float* cst = s.consts;
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = ((long long)it * n_clusters + cluster_id) * CL + cta_rank;
      query_tile_rows(s, e, P, cst, lane_taddr, row, half, tile);
    }
  }
  tc_teardown<CL>(tmem_base);
}


// CTA-pair variant of the query kernel (cta_group::2): two query tiles per cluster, M = 256 UMMAs issued by the leader, each
// CTA stages half of every weight operand (the single-CTA kernel pays 189 cycles per N = 256 UMMA -- 12 KB of shared-memory
// operand reads each -- the pair ~150).  Row-thread work is query_tile_rows unchanged.
__global__ void __launch_bounds__(HEAD_THREADS, 1)
query_mlp_pair_kernel(const QueryParams P, const __grid_constant__ CUtensorMap map_hi,
                      const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s = tc_carve(smem);
  s.pair_rank = (int)cluster_ctarank();
  s.terms = P.terms;
  for (int i = threadIdx.x; i < 8 * HID; i += HEAD_THREADS) s.consts[i] = P.consts[i];
  const uint32_t tmem_base = tc_prologue_pair<NEPI>(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;

  if (warp == 0) {
    ProdState ps{0};
    uint32_t afree_bits = 0xFu;
    if (lane == 0) { tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo); tma_prefetch_desc(&wmap); }
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = ((long long)it * n_clusters + cluster_id) * 2 + s.pair_rank;
      const int row0 = (int)min(tile * ROWS, (long long)0x7FFFFF00);     // past the end: all rows read as zero
      long long r = 0;                                                   // 128-byte row of the weight blob
      produce_job_tma_a_pair<NEPI>(s, ps, afree_bits, &wmap, r, P.slabs1, 2, &map_hi, &map_lo, row0);
      r += (long long)P.slabs1 * 2 * 2 * ROWS;
      for (int j = 0; j < 3; ++j) { produce_job_pair(s, ps, &wmap, r, 4, 2); r += 4 * 2 * 2 * ROWS; }
      afree_bits ^= 0xFu; afree_bits ^= 0xFu; afree_bits ^= 0xFu;        // layers 2..4: each slot written once by the row threads
    }
  } else if (warp == 1) {
    if (s.pair_rank == 0) {
      MmaState m{0, 0, 0};
      for (int it = 0; it < P.iters; ++it) {
        mma_job_pair(s, tmem_base, m, P.slabs1, 2, true);
        for (int j = 0; j < 3; ++j) mma_job_pair(s, tmem_base, m, 4, 2, true);
      }
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu, 0};
    const float* cst = s.consts;
    for (int it = 0; it < P.iters; ++it) {
      const long long tile = ((long long)it * n_clusters + cluster_id) * 2 + s.pair_rank;
      query_tile_rows(s, e, P, cst, lane_taddr, row, half, tile);
    }
  }
  tc_teardown_pair(tmem_base);
}

// =====================================================================================================
// fused head kernel: per 128 queries, 4 pair tiles + 1 query tile in ONE persistent CTA  (opt-in)
// =====================================================================================================
// The two kernels above hand the attended values x [total_q, Dvp] through HBM (1.5 GB written and read back at the
// bench size: 53x the head's algorithmic traffic, and 1.5 GB of workspace; 11.9 GB for an un-tiled x8 frame).  Here
// every CTA alternates: four pair tiles (4 x 32 queries) write their x rows into the CTA's OWN 128-row scratch block
// (148 x 128 rows x Dvp: 48 MB at C = 64, written and re-read by the same SM a few microseconds later, so it lives in
// L2), then the query MLP of those 128 queries runs with its layer-1 operand TMA-loaded from that block.  Same barriers,
// same job protocol; one more mbarrier (X_READY) orders the row threads' global stores of x before the producer's TMA
// loads of them.
// Measured (r02j / r02k, bench workload): 8.86 ms against 7.18 + 1.08 ms for the two kernels -- every phase switch
// drains the operand / weight pipeline that each separate kernel keeps full across tiles, and the x loads cannot be
// issued before x is stored.  (A variant that delays the query phase by one super-tile so that its loads are issued
// while the current x is still being stored needed a second scratch block and a PAIR_DONE barrier and was slower
// still: 9.65 ms.)  The x round trip of the two-kernel path costs ~0.5 ms of HBM time that overlaps with tensor work,
// so the fused kernel is NOT the default: it is selected with CIAOSR_HEAD_FUSED=1 when workspace matters more than
// the last 7 % of speed (its workspace is O(1) in the number of queries).
// smem: A slots and weight ring as in tc_carve, then barriers / TMEM slot / xchg, then BOTH kernels' constants
// (variable length, last):  [16 x 256 + Dvp pair constants][8 x 256 query constants].
constexpr int FU_BAR = SM_W + W_STAGES * SLAB_BYTES;
constexpr int FU_NBARS = N_BARS + 1;
constexpr int BAR_X_READY = N_BARS;
constexpr int FU_SLOT = FU_BAR + (FU_NBARS * 8 + 15) / 16 * 16;   // keeps everything after it 16-byte aligned
constexpr int FU_XCHG = FU_SLOT + 16;                   // 512 floats
constexpr int FU_CONST = FU_XCHG + 512 * 4;
__host__ __device__ constexpr int fused_query_const_off(int Dvp) { return 16 * HID + ((Dvp + 3) / 4) * 4; }
__host__ __device__ constexpr int fused_smem_bytes(int Dvp) { return FU_CONST + (fused_query_const_off(Dvp) + 8 * HID) * 4; }

template <int CL>
__global__ void __launch_bounds__(HEAD_THREADS, 1)
head_fused_kernel(const PairParams P, const QueryParams Q, const __grid_constant__ CUtensorMap map_hi,
                  const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s;
  {
    const uint32_t base = smem_u32(smem);
    s.a_hi = base + SM_A_HI; s.a_lo = base + SM_A_LO; s.w = base + SM_W;
    s.bar = base + FU_BAR;
    s.consts = reinterpret_cast<float*>(smem + FU_CONST);
    s.xchg = reinterpret_cast<float*>(smem + FU_XCHG);
    s.terms = P.terms;
  }
  const int qoff = fused_query_const_off(P.Dvp);
  for (int i = threadIdx.x; i < 16 * HID + P.Dvp; i += HEAD_THREADS) s.consts[i] = P.consts[i];
  for (int i = threadIdx.x; i < 8 * HID; i += HEAD_THREADS) s.consts[qoff + i] = Q.consts[i];
  if (threadIdx.x == 0) mbar_init(bar_at(s, BAR_X_READY), NEPI / 32);       // fenced + synced by the prologue below
  const uint32_t tmem_base = tc_prologue<CL, NEPI>(s, smem, FU_SLOT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks5 = (P.units5 + 1) / 2;
  const uint32_t cta_rank = CL > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int scratch_row0 = blockIdx.x * ROWS;            // this CTA's block of x rows

  if (warp == 0) {
    ProdState ps{0};
    uint32_t afree_bits = 0xFu;
    if (lane == 0) { tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo); }
    for (int it = 0; it < Q.iters; ++it) {
      for (int sub = 0; sub < 4; ++sub) {
        const uint8_t* b = P.blob;
        for (int j = 0; j < 6; ++j) { produce_job<CL>(s, ps, b, 4, 2, cta_rank); b += (size_t)8 * UNIT_BYTES; }
        for (int c = 0; c < nchunks5; ++c) {
          const int units = min(2, P.units5 - 2 * c);
          produce_job<CL>(s, ps, b, 4, units, cta_rank);
          b += (size_t)4 * units * UNIT_BYTES;
        }
      }
      // x of this super-tile is complete (and visible to the async proxy) once every row warp has arrived.  All of the
      // pair phase never commits A_FREE, so the slot parities below continue exactly as in query_mlp_kernel.
      mbar_wait(bar_at(s, BAR_X_READY), it & 1, 130);
      const uint8_t* b = Q.blob;
      produce_job_tma_a<CL, NEPI>(s, ps, afree_bits, b, Q.slabs1, 2, cta_rank, &map_hi, &map_lo, 0, scratch_row0);
      b += (size_t)Q.slabs1 * 2 * UNIT_BYTES;
      for (int j = 0; j < 3; ++j) { produce_job<CL>(s, ps, b, 4, 2, cta_rank); b += (size_t)8 * UNIT_BYTES; }
      afree_bits ^= 0xFu; afree_bits ^= 0xFu; afree_bits ^= 0xFu;        // layers 2..4: each slot written once by the row threads
    }
  } else if (warp == 1) {
    MmaState m{0, 0, 0};
    for (int it = 0; it < Q.iters; ++it) {
      for (int sub = 0; sub < 4; ++sub) {
        // pair phase: A_FREE is neither waited on nor committed (see PAIR_MMA); the query phase below uses it
        for (int j = 0; j < 6; ++j) mma_job<CL>(s, tmem_base, m, 4, 2, true, false);
        for (int c = 0; c < nchunks5; ++c) mma_job<CL>(s, tmem_base, m, 4, min(2, P.units5 - 2 * c), c == 0, false);
      }
      mma_job<CL>(s, tmem_base, m, Q.slabs1, 2, true);
      for (int j = 0; j < 3; ++j) mma_job<CL>(s, tmem_base, m, 4, 2, true);
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu, 0};                  // the pair phase toggles every slot parity an even number of times
    const float* cst = s.consts;
    const float* bv5 = s.consts + 16 * HID;
    const float* qcst = s.consts + qoff;
    PairInfo pinfo;
    for (int it = 0; it < Q.iters; ++it) {
      const long long st = ((long long)it * n_clusters + cluster_id) * CL + cta_rank;      // super-tile = 128 queries
      for (int sub = 0; sub < 4; ++sub) {
        // pipelined across the four pair tiles of a super-tile; the query tile in between uses the operand slots
        if (sub == 0) pinfo = pair_info_of(P, st * 4, row);
        pair_tile_rows<2>(s, e, P, cst, bv5, lane_taddr, row, half, lane, st * 4 + sub, scratch_row0 + sub * (ROWS / 4), pinfo,
                          sub == 0, sub < 3 ? st * 4 + sub + 1 : -1);
      }
      // publish x: global stores (generic proxy) -> visible device-wide -> visible to the TMA engine (async proxy)
      __threadfence();
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_at(s, BAR_X_READY));
      query_tile_rows(s, e, Q, qcst, lane_taddr, row, half, st);
    }
  }
  tc_teardown<CL>(tmem_base);
}

// CTA-pair form of the fused kernel [r2b]: the same alternation of four pair tiles and one query tile per 128 queries, with
// the M = 256 cta_group::2 UMMAs of pair_mlp_pair_kernel / query_mlp_pair_kernel: both CTAs of a cluster walk the job
// sequence in lock-step on their own queries and their own x scratch block (X_READY stays local to each CTA; the layer-1
// operand loads of both CTAs count on the leader's A_READY).  Default form of the fused path when CTA pairs are on.
__global__ void __launch_bounds__(HEAD_THREADS, 1)
head_fused_pair_kernel(const PairParams P, const QueryParams Q, const __grid_constant__ CUtensorMap map_hi,
                       const __grid_constant__ CUtensorMap map_lo, const __grid_constant__ CUtensorMap wmap_p,
                       const __grid_constant__ CUtensorMap wmap_q) {
  extern __shared__ __align__(1024) uint8_t smem[];
  TcShared s;
  {
    const uint32_t base = smem_u32(smem);
    s.a_hi = base + SM_A_HI; s.a_lo = base + SM_A_LO; s.w = base + SM_W;
    s.bar = base + FU_BAR;
    s.consts = reinterpret_cast<float*>(smem + FU_CONST);
    s.xchg = reinterpret_cast<float*>(smem + FU_XCHG);
    s.terms = P.terms;
    s.pair_rank = (int)cluster_ctarank();
  }
  const int qoff = fused_query_const_off(P.Dvp);
  for (int i = threadIdx.x; i < 16 * HID + P.Dvp; i += HEAD_THREADS) s.consts[i] = P.consts[i];
  for (int i = threadIdx.x; i < 8 * HID; i += HEAD_THREADS) s.consts[qoff + i] = Q.consts[i];
  if (threadIdx.x == 0) mbar_init(bar_at(s, BAR_X_READY), NEPI / 32);       // fenced + synced by the prologue below
  const uint32_t tmem_base = tc_prologue_pair<NEPI>(s, smem, FU_SLOT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks5 = (P.units5 + 1) / 2;
  const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
  const int scratch_row0 = blockIdx.x * ROWS;            // this CTA's block of x rows

  if (warp == 0) {
    ProdState ps{0};
    uint32_t afree_bits = 0xFu;
    if (lane == 0) { tma_prefetch_desc(&map_hi); tma_prefetch_desc(&map_lo); tma_prefetch_desc(&wmap_p); tma_prefetch_desc(&wmap_q); }
    for (int it = 0; it < Q.iters; ++it) {
      for (int sub = 0; sub < 4; ++sub) {
        long long r = 0;
        for (int j = 0; j < 6; ++j) { produce_job_pair(s, ps, &wmap_p, r, 4, 2); r += 4 * 2 * 2 * ROWS; }
        for (int c = 0; c < nchunks5; ++c) {
          const int units = min(2, P.units5 - 2 * c);
          produce_job_pair(s, ps, &wmap_p, r, 4, units);
          r += 4 * units * 2 * ROWS;
        }
      }
      mbar_wait(bar_at(s, BAR_X_READY), it & 1, 130);   // this CTA's x block is complete and visible to the async proxy
      long long r = 0;
      produce_job_tma_a_pair<NEPI>(s, ps, afree_bits, &wmap_q, r, Q.slabs1, 2, &map_hi, &map_lo, scratch_row0);
      r += (long long)Q.slabs1 * 2 * 2 * ROWS;
      for (int j = 0; j < 3; ++j) { produce_job_pair(s, ps, &wmap_q, r, 4, 2); r += 4 * 2 * 2 * ROWS; }
      afree_bits ^= 0xFu; afree_bits ^= 0xFu; afree_bits ^= 0xFu;
    }
  } else if (warp == 1) {
    if (s.pair_rank == 0) {
      MmaState m{0, 0, 0};
      for (int it = 0; it < Q.iters; ++it) {
        for (int sub = 0; sub < 4; ++sub) {
          for (int j = 0; j < 6; ++j) mma_job_pair(s, tmem_base, m, 4, 2, true, false);
          for (int c = 0; c < nchunks5; ++c) mma_job_pair(s, tmem_base, m, 4, min(2, P.units5 - 2 * c), c == 0, false);
        }
        mma_job_pair(s, tmem_base, m, Q.slabs1, 2, true);
        for (int j = 0; j < 3; ++j) mma_job_pair(s, tmem_base, m, 4, 2, true);
      }
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu, 0};
    const float* cst = s.consts;
    const float* bv5 = s.consts + 16 * HID;
    const float* qcst = s.consts + qoff;
    PairInfo pinfo;
    for (int it = 0; it < Q.iters; ++it) {
      const long long st = ((long long)it * n_clusters + cluster_id) * 2 + s.pair_rank;      // super-tile = 128 queries
      for (int sub = 0; sub < 4; ++sub) {
        if (sub == 0) pinfo = pair_info_of(P, st * 4, row);
        pair_tile_rows<2>(s, e, P, cst, bv5, lane_taddr, row, half, lane, st * 4 + sub, scratch_row0 + sub * (ROWS / 4), pinfo,
                          sub == 0, sub < 3 ? st * 4 + sub + 1 : -1);
      }
      __threadfence();
      asm volatile("fence.proxy.async.global;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_at(s, BAR_X_READY));
      query_tile_rows(s, e, Q, qcst, lane_taddr, row, half, st);
    }
  }
  tc_teardown_pair(tmem_base);
}

}  // namespace ciaosr
