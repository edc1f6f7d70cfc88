// tcgen05 engine for the implicit attention head (CIAOSR_ENGINE_TCGEN05), sm_100a.
//
// Two persistent, warp-specialised kernels run the MLP stacks on the 5th-gen tensor cores
// with fp32-grade accuracy (bf16 hi/lo split, 3 UMMAs per product, fp32 accumulation in TMEM):
//
//   pair_mlp_kernel   per 128 (query, neighbour) rows = 32 queries x 4 neighbours:
//       layer 1 of imnet_k / imnet_v from the LR-resolution hoist (gather + 4 FMAs, ciaosr_net.py:195-205)
//       hidden layers 2..4 of both MLPs on UMMA (M=128, N=256, K=256 each)
//       key side: logit = h4k . G[query px, offset] (+c0), softmax over the 4 neighbours (:203,:214-215)
//       value side: last Linear (256 -> Dv) on UMMA in N-chunks, fused with  sum_n a_n * value_n * W_v  (:206,:215)
//       -> x[query, Dv]
//   query_mlp_kernel  per 128 queries: imnet_q (Dv -> 256 x4 -> 3) + bilinear residual (:221, :107-108)
//
// Pipeline inside a CTA (1 CTA / SM, all 512 TMEM columns = two 128x256 fp32 accumulators):
//   warp 0  weight producer: cp.async.bulk (TMA engine) of pre-swizzled 32 KB weight units into a 2-stage ring
//   warp 1  UMMA issuer (one lane): waits operand slabs / weight units, issues tcgen05.mma, commits to mbarriers
//   warp 2  TMEM allocator
//   warps 4-7  one thread per row: build layer-1 operands, drain accumulators (tcgen05.ld),
//              bias+ReLU, bf16 hi/lo split, write the next layer's A operand straight into the
//              128B-swizzled K-major smem slabs the next UMMA reads.  Slab-granular mbarriers let layer
//              l+1's UMMAs start as soon as the first 64 columns of layer l are converted, while the
//              second accumulator absorbs them.
// Hidden activations never leave the SM.  Weights stream from L2 (2.2 MB per 128 rows at C=64).
#include "kernels.cuh"
#include "pairs.cuh"
#include "gemm_tc.cuh"

namespace ciaosr {
using namespace tc;

// =====================================================================================================
// pair kernel
// =====================================================================================================
struct PairParams {
  PairConsts pc;
  const float* coord; const float* cell;
  const float* featT; const float* nlT;
  int C, Cn, Dv, Dvp;
  const float* Pk; const float* Pv; const float* G; int ldg;
  const float* consts;        // 16 x 256 floats, see pair_consts layout below
  const float* bv5p;          // [Dvp] last-layer value bias, tap-major, zero padded
  const uint8_t* blob; int units_per_tile; int units5;
  float* x;                   // [total_q, Dvp]
  long long total_rows; int n_tiles;
  float softmax_scale;
};
// consts layout (x256 floats): 0..3 rc_k, 4 b1_k, 5..7 b_k(layers 2..4), 8..11 rc_v, 12 b1_v, 13..15 b_v(2..4)

__device__ __forceinline__ void gen_layer1(const TcShared& s, EpiState& e, int row, const PairInfo& p,
                                           const float* __restrict__ P, const float* __restrict__ rc_s,
                                           const float* __restrict__ b1_s) {
  const float4* prow = p.pix >= 0 ? reinterpret_cast<const float4*>(P + (long long)p.pix * HID) : nullptr;
#pragma unroll 1
  for (int sl = 0; sl < 4; ++sl) {
    slab_begin(s, e, sl, false);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      const int c0 = sl * 64 + half * 32;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = prow ? __ldg(prow + (c0 >> 2) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int c = c0 + i;
        float t = v[i] + b1_s[c];
        t = fmaf(rc_s[c], p.rel_y, t);
        t = fmaf(rc_s[HID + c], p.rel_x, t);
        t = fmaf(rc_s[2 * HID + c], p.sc_y, t);
        t = fmaf(rc_s[3 * HID + c], p.sc_x, t);
        v[i] = fmaxf(t, 0.0f);
      }
      a_store32(s.a_hi + sl * SLAB_BYTES, s.a_lo + sl * SLAB_BYTES, row, half * 32, v);
    }
    slab_done(s, sl);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) pair_mlp_kernel(const PairParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcShared s = tc_carve(smem);
  for (int i = threadIdx.x; i < CONST_FLOATS; i += TC_THREADS) s.consts[i] = P.consts[i];
  const uint32_t tmem_base = tc_prologue(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks5 = (P.units5 + 1) / 2;

  if (warp == 0) {
    producer_loop(s, P.blob, P.units_per_tile, P.n_tiles);
  } else if (warp == 1) {
    MmaState m{0, 0, 0, 0};
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      for (int j = 0; j < 6; ++j) mma_job(s, tmem_base, m, 4, 2, true);
      for (int c = 0; c < nchunks5; ++c) mma_job(s, tmem_base, m, 4, min(2, P.units5 - 2 * c), c == 0);
    }
  } else if (warp >= 4) {
    const int row = threadIdx.x - EPI_T0;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0};
    const float* cst = s.consts;
    const int C = P.C, H = P.pc.H, W = P.pc.W;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const long long R = (long long)tile * ROWS + row;
      const bool valid = R < P.total_rows;
      PairInfo p;
      if (valid) p = compute_pair(P.pc, P.coord, P.cell, R >> 2, (int)(R & 3));
      else { p.pix = -1; p.gidx = -1; p.rel_y = p.rel_x = p.sc_y = p.sc_x = 0.0f; }

      // ---- key chain -----------------------------------------------------------------------
      gen_layer1(s, e, row, p, P.Pk, cst, cst + 4 * HID);
      epi_hidden<false>(s, e, lane_taddr, row, cst + 5 * HID);
      epi_hidden<false>(s, e, lane_taddr, row, cst + 6 * HID);
      float logit = 0.0f;
      {
        const uint32_t d = epi_wait_d(s, e);
        const float* bias_s = cst + 7 * HID;
        const float4* grow = p.gidx >= 0 ? reinterpret_cast<const float4*>(P.G + (long long)p.gidx * P.ldg) : nullptr;
#pragma unroll 1
        for (int c0 = 0; c0 < HID; c0 += 32) {
          float v[32];
          tmem_ld32(lane_taddr + d * 256 + c0, v);
          if (grow) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g = __ldg(grow + (c0 >> 2) + j);
              logit = fmaf(fmaxf(v[4 * j] + bias_s[c0 + 4 * j], 0.0f), g.x, logit);
              logit = fmaf(fmaxf(v[4 * j + 1] + bias_s[c0 + 4 * j + 1], 0.0f), g.y, logit);
              logit = fmaf(fmaxf(v[4 * j + 2] + bias_s[c0 + 4 * j + 2], 0.0f), g.z, logit);
              logit = fmaf(fmaxf(v[4 * j + 3] + bias_s[c0 + 4 * j + 3], 0.0f), g.w, logit);
            }
          }
        }
        if (grow) logit += __ldg(P.G + (long long)p.gidx * P.ldg + HID);
        epi_release_d(s, e);
      }
      // softmax over the 4 neighbours of this query (4 adjacent lanes)
      float a;
      {
        const float l = __fdiv_rn(logit, P.softmax_scale);
        float mx = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float ex = expf(l - mx);
        float sum = ex + __shfl_xor_sync(0xffffffffu, ex, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        a = valid ? __fdiv_rn(ex, sum) : 0.0f;
      }

      // ---- value chain -----------------------------------------------------------------------
      gen_layer1(s, e, row, p, P.Pv, cst + 8 * HID, cst + 12 * HID);
      epi_hidden<false>(s, e, lane_taddr, row, cst + 13 * HID);
      epi_hidden<false>(s, e, lane_taddr, row, cst + 14 * HID);
      epi_hidden<false>(s, e, lane_taddr, row, cst + 15 * HID);   // h4v -> operand of the last Linear

      // geometry of this row's latent code for the value gather
      int py = 0, px = 0;
      const float* fbase = nullptr;
      const float* nbase = nullptr;
      if (p.pix >= 0) {
        const int hw = p.pix % (H * W);
        py = hw / W; px = hw % W;
        fbase = P.featT + (long long)p.pix * C;
        if (P.nlT) nbase = P.nlT + (long long)p.pix * P.Cn;
      }
      const long long q = R >> 2;
      const int sub = (lane & 1) * 16 + ((lane >> 1) & 1) * 8;   // columns of a 32-chunk this lane ends up owning
      for (int c = 0; c < nchunks5; ++c) {
        const int units = min(2, P.units5 - 2 * c);
        const uint32_t d = epi_wait_d(s, e);
#pragma unroll 1
        for (int cc = 0; cc < units * 4; ++cc) {
          const int cp0 = c * 256 + cc * 32;
          float v[32];
          tmem_ld32(lane_taddr + d * 256 + cc * 32, v);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int cp = cp0 + 4 * g;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (fbase != nullptr && cp < P.Dv) {
              if (cp < 9 * C) {
                const int t = cp / C, ch = cp - t * C;
                const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                  val = __ldg(reinterpret_cast<const float4*>(fbase + ((t / 3 - 1) * W + (t % 3 - 1)) * C + ch));
              } else {
                val = __ldg(reinterpret_cast<const float4*>(nbase + (cp - 9 * C)));
              }
            }
            const float4 b = __ldg(reinterpret_cast<const float4*>(P.bv5p + cp));
            v[4 * g] = a * val.x * (v[4 * g] + b.x);
            v[4 * g + 1] = a * val.y * (v[4 * g + 1] + b.y);
            v[4 * g + 2] = a * val.z * (v[4 * g + 2] + b.z);
            v[4 * g + 3] = a * val.w * (v[4 * g + 3] + b.w);
          }
          // sum over the 4 neighbour rows (lanes 4q..4q+3), leaving each lane with 8 of the 32 columns
          float r16[16], r8[8];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float keep = (lane & 1) ? v[16 + i] : v[i];
            const float send = (lane & 1) ? v[i] : v[16 + i];
            r16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float keep = (lane & 2) ? r16[8 + i] : r16[i];
            const float send = (lane & 2) ? r16[i] : r16[8 + i];
            r8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(P.x + q * P.Dvp + cp0 + sub);
            dst[0] = make_float4(r8[0], r8[1], r8[2], r8[3]);
            dst[1] = make_float4(r8[4], r8[5], r8[6], r8[7]);
          }
        }
        epi_release_d(s, e);
      }
    }
  }
  tc_epilogue_dealloc(tmem_base);
}

// =====================================================================================================
// query kernel: imnet_q + residual
// =====================================================================================================
struct QueryParams {
  const float* x; int Dvp;                 // [total_q, Dvp]
  const float* consts;                     // x256 floats: 0..3 b_q(layers 1..4), 4..6 W5 rows, 7: b5 in [0..3)
  const uint8_t* blob; int units_per_tile; int slabs1;
  const float* lr; const float* coord; int H, W, Q;
  float* out; long long total_q; int n_tiles;
};

__global__ void __launch_bounds__(TC_THREADS, 1) query_mlp_kernel(const QueryParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcShared s = tc_carve(smem);
  for (int i = threadIdx.x; i < 8 * HID; i += TC_THREADS) s.consts[i] = P.consts[i];
  const uint32_t tmem_base = tc_prologue(s, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    producer_loop(s, P.blob, P.units_per_tile, P.n_tiles);
  } else if (warp == 1) {
    MmaState m{0, 0, 0, 0};
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      mma_job(s, tmem_base, m, P.slabs1, 2, true);
      for (int j = 0; j < 3; ++j) mma_job(s, tmem_base, m, 4, 2, true);
    }
  } else if (warp >= 4) {
    const int row = threadIdx.x - EPI_T0;
    const uint32_t lane_taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    EpiState e{0, 0xFu};                     // A_free waits start at parity 1 (fresh barrier passes)
    const float* cst = s.consts;
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      const long long g = (long long)tile * ROWS + row;
      const bool valid = g < P.total_q;
      const float4* xrow = valid ? reinterpret_cast<const float4*>(P.x + g * P.Dvp) : nullptr;
      // layer-1 operand: x (fp32) -> bf16 hi/lo slabs, streamed through the 4 slots
#pragma unroll 1
      for (int sl = 0; sl < P.slabs1; ++sl) {
        const int slot = sl & 3;
        slab_begin(s, e, slot, true);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = xrow ? __ldg(xrow + sl * 16 + half * 8 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
          }
          a_store32(s.a_hi + slot * SLAB_BYTES, s.a_lo + slot * SLAB_BYTES, row, half * 32, v);
        }
        slab_done(s, slot);
      }
      epi_hidden<true>(s, e, lane_taddr, row, cst);
      epi_hidden<true>(s, e, lane_taddr, row, cst + HID);
      epi_hidden<true>(s, e, lane_taddr, row, cst + 2 * HID);
      // last hidden layer + the 256 -> 3 Linear on CUDA cores
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      {
        const uint32_t d = epi_wait_d(s, e);
        const float* bias_s = cst + 3 * HID;
#pragma unroll 1
        for (int c0 = 0; c0 < HID; c0 += 32) {
          float v[32];
          tmem_ld32(lane_taddr + d * 256 + c0, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float h = fmaxf(v[i] + bias_s[c0 + i], 0.0f);
            o0 = fmaf(h, cst[4 * HID + c0 + i], o0);
            o1 = fmaf(h, cst[5 * HID + c0 + i], o1);
            o2 = fmaf(h, cst[6 * HID + c0 + i], o2);
          }
        }
        epi_release_d(s, e);
      }
      if (valid) {
        o0 += cst[7 * HID]; o1 += cst[7 * HID + 1]; o2 += cst[7 * HID + 2];
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
    }
  }
  tc_epilogue_dealloc(tmem_base);
}

// =====================================================================================================
// plan-time packing
// =====================================================================================================
// blob layout (bytes): [pair units][query units]; a unit = 128 weight rows x 64 K as two SW128 slabs (hi, lo)
struct TcLayout {
  int Dvp, units5, pair_units, slabs1, query_units;
  int slabs_k, slabs_v;              // K-slabs of the LR-resolution GEMMs (9C and Dv)
  size_t pair_blob, query_blob;      // byte offsets in the blob
  size_t k1_blob, v1_blob, kfin_blob;   // layer-1 hoists and the key fold, N = 256 each
  size_t pair_consts, bv5p, query_consts;   // byte offsets (float arrays)
  size_t total;
};

static TcLayout tc_layout(int C, int Cn) {
  TcLayout t;
  const int Dv = 9 * C + Cn;
  t.Dvp = (Dv + 127) / 128 * 128;
  t.units5 = t.Dvp / 128;
  t.pair_units = 6 * 4 * 2 + 4 * t.units5;
  t.slabs1 = t.Dvp / 64;
  t.query_units = t.slabs1 * 2 + 3 * 4 * 2;
  size_t off = 0;
  t.pair_blob = off; off += (size_t)t.pair_units * UNIT_BYTES;
  t.query_blob = off; off += (size_t)t.query_units * UNIT_BYTES;
  t.slabs_k = (9 * C + KSLAB - 1) / KSLAB;
  t.slabs_v = (Dv + KSLAB - 1) / KSLAB;
  t.k1_blob = off; off += (size_t)t.slabs_k * 2 * UNIT_BYTES;
  t.v1_blob = off; off += (size_t)t.slabs_v * 2 * UNIT_BYTES;
  t.kfin_blob = off; off += (size_t)t.slabs_k * 2 * UNIT_BYTES;
  t.pair_consts = off; off += CONST_FLOATS * 4;
  t.bv5p = off; off += (size_t)t.Dvp * 4;
  t.query_consts = off; off += 8 * HID * 4;
  t.total = (off + 255) / 256 * 256;
  return t;
}

bool tc_shapes_ok(const ciaosr_head_desc* d) {
  if (d->local_size != 2 || d->channels % 4 != 0) return false;
  const ciaosr_mlp_desc* ms[3] = {&d->imnet_q, &d->imnet_k, &d->imnet_v};
  for (auto m : ms) {
    if (m->n_layers != 5) return false;
    for (int l = 1; l <= 4; ++l) if (m->dims[l] != HID) return false;
  }
  return true;
}

size_t tc_blob_bytes(const ciaosr_head_desc* d) {
  const int Cn = d->non_local_attn ? d->channels * d->cs_attn.n_scales : 0;
  return tc_layout(d->channels, Cn).total;
}

// One job's weights -> units.  Element (n, k) of the job = W[rowmap(n0 + n)][colmap(k)] (0 outside).
//   W row-major [n_valid_src, ld];  perm_rows / perm_cols: tap-major -> reference channel order.
__global__ void tc_pack_job_kernel(uint8_t* __restrict__ dst, const float* __restrict__ W, int ld,
                                   int n_valid, int k_valid, int nslabs, int units, int n0, int perm_rows,
                                   int perm_cols, int C, int trans) {
  const long long total = (long long)nslabs * units * UNIT_N * KSLAB;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k_in = (int)(i % KSLAB);
  const int n_in = (int)((i / KSLAB) % UNIT_N);
  const int unit = (int)(i / (KSLAB * UNIT_N));          // = sl * units + u
  const int sl = unit / units, u = unit % units;
  const int n = n0 + u * UNIT_N + n_in, k = sl * KSLAB + k_in;
  float w = 0.0f;
  if (n < n_valid && k < k_valid) {
    const int Dk = 9 * C;
    const int ns = (perm_rows && n < Dk) ? (n % C) * 9 + n / C : n;
    const int ks = (perm_cols && k < Dk) ? (k % C) * 9 + k / C : k;
    w = trans ? W[(long long)ks * ld + ns] : W[(long long)ns * ld + ks];   // trans: W is [K, N]
  }
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  uint8_t* ub = dst + (size_t)unit * UNIT_BYTES;
  const uint32_t off = sw128_offset(n_in, k_in);
  *reinterpret_cast<__nv_bfloat16*>(ub + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(ub + SLAB_BYTES + off) = lo;
}

__global__ void tc_pack_consts_kernel(float* __restrict__ pc, float* __restrict__ bv5p, float* __restrict__ qc,
                                      const float* __restrict__ plan, PlanLayout L, int Dvp,
                                      const float* __restrict__ q5w, const float* __restrict__ q5b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * HID) { pc[i] = plan[L.k.rc + i]; pc[8 * HID + i] = plan[L.v.rc + i]; }
  if (i < HID) {
    pc[4 * HID + i] = plan[L.k.bias[0] + i];
    pc[12 * HID + i] = plan[L.v.bias[0] + i];
    for (int l = 1; l <= 3; ++l) {
      pc[(4 + l) * HID + i] = plan[L.k.bias[l] + i];
      pc[(12 + l) * HID + i] = plan[L.v.bias[l] + i];
      qc[l * HID + i] = plan[L.q.bias[l] + i];
    }
    qc[i] = plan[L.q.bias[0] + i];
    for (int c = 0; c < 3; ++c) qc[(4 + c) * HID + i] = q5w[c * HID + i];
    qc[7 * HID + i] = i < 3 ? q5b[i] : 0.0f;
  }
  if (i < Dvp) bv5p[i] = i < L.Dv ? plan[L.v.bias[4] + i] : 0.0f;
}

int tc_pack(const ciaosr_head_desc* d, const PlanLayout& L, float* plan, cudaStream_t st) {
  const TcLayout t = tc_layout(L.C, L.Cn);
  uint8_t* blob = reinterpret_cast<uint8_t*>(plan + L.tc_blob);
  auto pack = [&](uint8_t* dst, const float* W, int ld, int n_valid, int k_valid, int nslabs, int units,
                  int n0, int pr, int pcol, int trans = 0) -> int {
    const long long total = (long long)nslabs * units * UNIT_N * KSLAB;
    CIAOSR_LAUNCH(tc_pack_job_kernel, cdiv(total, 256), 256, 0, st, dst, W, ld, n_valid, k_valid, nslabs,
                  units, n0, pr, pcol, L.C, trans);
    return CIAOSR_OK;
  };
  int rc;
  uint8_t* p = blob + t.pair_blob;
  for (int l = 1; l <= 3; ++l) {    // imnet_k hidden layers 2..4
    if ((rc = pack(p, d->imnet_k.weight[l], HID, HID, HID, 4, 2, 0, 0, 0))) return rc;
    p += (size_t)8 * UNIT_BYTES;
  }
  for (int l = 1; l <= 3; ++l) {    // imnet_v hidden layers 2..4
    if ((rc = pack(p, d->imnet_v.weight[l], HID, HID, HID, 4, 2, 0, 0, 0))) return rc;
    p += (size_t)8 * UNIT_BYTES;
  }
  for (int c = 0; c * 2 < t.units5; ++c) {   // imnet_v last layer, N-chunks of <= 256 tap-major outputs
    const int units = t.units5 - 2 * c < 2 ? t.units5 - 2 * c : 2;
    if ((rc = pack(p, d->imnet_v.weight[4], HID, L.Dv, HID, 4, units, c * 256, 1, 0))) return rc;
    p += (size_t)4 * units * UNIT_BYTES;
  }
  p = blob + t.query_blob;
  if ((rc = pack(p, d->imnet_q.weight[0], L.Dv, HID, L.Dv, t.slabs1, 2, 0, 0, 1))) return rc;
  p += (size_t)t.slabs1 * 2 * UNIT_BYTES;
  for (int l = 1; l <= 3; ++l) {
    if ((rc = pack(p, d->imnet_q.weight[l], HID, HID, HID, 4, 2, 0, 0, 0))) return rc;
    p += (size_t)8 * UNIT_BYTES;
  }
  // LR-resolution GEMM operands: layer-1 hoists W1[:, :D] (tap-major K) and the key fold W5k^T
  if ((rc = pack(blob + t.k1_blob, d->imnet_k.weight[0], L.Dk + 4, HID, L.Dk, t.slabs_k, 2, 0, 0, 1))) return rc;
  if ((rc = pack(blob + t.v1_blob, d->imnet_v.weight[0], L.Dv + 4, HID, L.Dv, t.slabs_v, 2, 0, 0, 1))) return rc;
  if ((rc = pack(blob + t.kfin_blob, d->imnet_k.weight[4], HID, HID, L.Dk, t.slabs_k, 2, 0, 0, 1, 1))) return rc;
  const int n = t.Dvp > 4 * HID ? t.Dvp : 4 * HID;
  CIAOSR_LAUNCH(tc_pack_consts_kernel, cdiv(n, 256), 256, 0, st,
                reinterpret_cast<float*>(blob + t.pair_consts), reinterpret_cast<float*>(blob + t.bv5p),
                reinterpret_cast<float*>(blob + t.query_consts), plan, L, t.Dvp, d->imnet_q.weight[4],
                d->imnet_q.bias[4]);
  return CIAOSR_OK;
}

// =====================================================================================================
// LR-resolution precompute on the tensor cores (replaces run_lr_precompute's CUDA-core GEMMs)
// =====================================================================================================
struct UnfoldGen {       // A[pix, kp] = tap-major 3x3 unfold of the NHWC feature (+ non-local channels)
  const float* f; const float* nl; int H, W, C, Cn, K;
  struct Row { int y, x; };
  __device__ __forceinline__ Row row(long long m) const {
    const int hw = (int)(m % ((long long)H * W));
    return Row{hw / W, hw % W};
  }
  __device__ __forceinline__ void fill(Row& r, long long m, int k0, float (&v)[32]) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < 9 * C) {
        const int t = k / C, ch = k - t * C;
        const int dy = t / 3 - 1, dx = t % 3 - 1;
        if (r.y + dy >= 0 && r.y + dy < H && r.x + dx >= 0 && r.x + dx < W)
          q = __ldg(reinterpret_cast<const float4*>(f + (m + dy * W + dx) * C + ch));
      } else if (k < K) {
        q = __ldg(reinterpret_cast<const float4*>(nl + m * Cn + (k - 9 * C)));
      }
      v[4 * g] = q.x; v[4 * g + 1] = q.y; v[4 * g + 2] = q.z; v[4 * g + 3] = q.w;
    }
  }
};
struct PairProdGen {     // A[pix*9 + d, kp] = U[pix, kp] * U[pix + d, kp];  also c0 = sum_kp A * b5[kp]
  const float* f; const float* fin; int H, W, C, ldfin;     // fin[kp * ldfin + 256] = permuted last-layer bias
  struct Row { int y, x, dy, dx; long long pix; bool ok; float c0; };
  __device__ __forceinline__ Row row(long long m) const {
    Row r;
    r.pix = m / 9;
    const int d = (int)(m % 9), hw = (int)(r.pix % ((long long)H * W));
    r.y = hw / W; r.x = hw % W; r.dy = d / 3 - 1; r.dx = d % 3 - 1;
    r.ok = r.y + r.dy >= 0 && r.y + r.dy < H && r.x + r.dx >= 0 && r.x + r.dx < W;
    r.c0 = 0.0f;
    return r;
  }
  __device__ __forceinline__ void fill(Row& r, long long, int k0, float (&v)[32]) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r.ok && k < 9 * C) {
        const int t = k / C, ch = k - t * C;
        const int ty = t / 3 - 1, tx = t % 3 - 1;
        const int ay = r.y + ty, ax = r.x + tx, by = ay + r.dy, bx = ax + r.dx;
        if (ay >= 0 && ay < H && ax >= 0 && ax < W && by >= 0 && by < H && bx >= 0 && bx < W) {
          const float* pa = f + (r.pix + ty * W + tx) * C + ch;
          const float4 a = __ldg(reinterpret_cast<const float4*>(pa));
          const float4 b = __ldg(reinterpret_cast<const float4*>(pa + (r.dy * W + r.dx) * C));
          q = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
          r.c0 = fmaf(q.x, __ldg(fin + (long long)k * ldfin + HID), r.c0);
          r.c0 = fmaf(q.y, __ldg(fin + (long long)(k + 1) * ldfin + HID), r.c0);
          r.c0 = fmaf(q.z, __ldg(fin + (long long)(k + 2) * ldfin + HID), r.c0);
          r.c0 = fmaf(q.w, __ldg(fin + (long long)(k + 3) * ldfin + HID), r.c0);
        }
      }
      v[4 * g] = q.x; v[4 * g + 1] = q.y; v[4 * g + 2] = q.z; v[4 * g + 3] = q.w;
    }
  }
};
template <class Row>
struct StoreRowsEpi {    // C[m, n0 .. n0+32) = v   (ld % 4 == 0)
  float* c; int ld;
  __device__ __forceinline__ void store(const Row&, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(c + m * ld + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
};
struct StoreGEpi {       // G rows: 256 folded weights + the folded bias term in column 256
  float* g; int ld;
  __device__ __forceinline__ void store(const PairProdGen::Row& r, long long m, int n0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(g + m * ld + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    if (n0 == HID - 32) g[m * ld + HID] = r.c0;
  }
};

static int run_lr_precompute_tc(const PlanLayout& L, const float* plan, const HeadArgs& a, float* Pk, float* Pv,
                                float* G, int ldg, cudaStream_t st) {
  const TcLayout t = tc_layout(L.C, L.Cn);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(plan + L.tc_blob);
  const long long npix = (long long)a.B * a.H * a.W;
  int rc;
  UnfoldGen uk{a.featT, a.nlT, a.H, a.W, L.C, L.Cn, L.Dk};
  UnfoldGen uv{a.featT, a.nlT, a.H, a.W, L.C, L.Cn, L.Dv};
  if ((rc = tc_gemm(GemmShape{npix, t.slabs_k, 2, npix, 0}, blob + t.k1_blob, uk,
                    StoreRowsEpi<UnfoldGen::Row>{Pk, HID}, st))) return rc;
  if ((rc = tc_gemm(GemmShape{npix, t.slabs_v, 2, npix, 0}, blob + t.v1_blob, uv,
                    StoreRowsEpi<UnfoldGen::Row>{Pv, HID}, st))) return rc;
  PairProdGen pg{a.featT, plan + L.k.fin, a.H, a.W, L.C, HID + 1};
  if ((rc = tc_gemm(GemmShape{npix * 9, t.slabs_k, 2, npix * 9, 0}, blob + t.kfin_blob, pg, StoreGEpi{G, ldg}, st)))
    return rc;
  return CIAOSR_OK;
}

// =====================================================================================================
// host orchestration
// =====================================================================================================
struct TcBufs { float *Pk, *Pv, *G, *x; };
static int tc_ldg() { return HID + 4; }

static TcBufs tc_carve_ws(Arena& a, const PlanLayout& L, int B, int H, int W, int Q) {
  TcBufs s;
  const TcLayout t = tc_layout(L.C, L.Cn);
  const size_t npix = (size_t)B * H * W;
  s.Pk = a.take<float>(npix * HID);
  s.Pv = a.take<float>(npix * HID);
  s.G = a.take<float>(npix * 9 * tc_ldg());
  s.x = a.take<float>((size_t)B * Q * t.Dvp);
  return s;
}

size_t head_tc_workspace(const PlanLayout& L, int B, int H, int W, int Q) {
  Arena a(nullptr, 0);
  tc_carve_ws(a, L, B, H, W, Q);
  return a.used();
}

static int tc_grid(int n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return n_tiles < sms ? n_tiles : sms;
}

int run_head_tc(const PlanLayout& L, const float* plan, const HeadArgs& a, void* ws, size_t ws_bytes,
                cudaStream_t st) {
  Arena ar(ws, ws_bytes);
  TcBufs b = tc_carve_ws(ar, L, a.B, a.H, a.W, a.Q);
  CIAOSR_REQUIRE(ar.ok, CIAOSR_E_WORKSPACE, "head (tcgen05) workspace too small: need %zu, have %zu",
                 ar.used(), ws_bytes);
  const TcLayout t = tc_layout(L.C, L.Cn);
  const uint8_t* blob = reinterpret_cast<const uint8_t*>(plan + L.tc_blob);
  int rc;
  {
    StageScope sc(2, st);
    if ((rc = run_lr_precompute_tc(L, plan, a, b.Pk, b.Pv, b.G, tc_ldg(), st))) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(pair_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    CIAOSR_CUDA_OK(cudaFuncSetAttribute(query_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    attr_set = true;
  }
  const long long total_q = (long long)a.B * a.Q;
  {
    StageScope sc(3, st);
    PairParams P;
    P.pc = PairConsts{a.H, a.W, a.Q, a.eval_bsize, L.local_size, a.cy0, a.cy1, a.cx0, a.cx1};
    P.coord = a.coord; P.cell = a.cell; P.featT = a.featT; P.nlT = a.nlT;
    P.C = L.C; P.Cn = L.Cn; P.Dv = L.Dv; P.Dvp = t.Dvp;
    P.Pk = b.Pk; P.Pv = b.Pv; P.G = b.G; P.ldg = tc_ldg();
    P.consts = reinterpret_cast<const float*>(blob + t.pair_consts);
    P.bv5p = reinterpret_cast<const float*>(blob + t.bv5p);
    P.blob = blob + t.pair_blob; P.units_per_tile = t.pair_units; P.units5 = t.units5;
    P.x = b.x; P.total_rows = total_q * 4; P.n_tiles = (int)((P.total_rows + ROWS - 1) / ROWS);
    P.softmax_scale = L.softmax_scale;
    CIAOSR_LAUNCH(pair_mlp_kernel, tc_grid(P.n_tiles), TC_THREADS, SM_TOTAL, st, P);
  }
  {
    StageScope sc(4, st);
    QueryParams Qp;
    Qp.x = b.x; Qp.Dvp = t.Dvp;
    Qp.consts = reinterpret_cast<const float*>(blob + t.query_consts);
    Qp.blob = blob + t.query_blob; Qp.units_per_tile = t.query_units; Qp.slabs1 = t.slabs1;
    Qp.lr = a.lr; Qp.coord = a.coord; Qp.H = a.H; Qp.W = a.W; Qp.Q = a.Q;
    Qp.out = a.out; Qp.total_q = total_q; Qp.n_tiles = (int)((total_q + ROWS - 1) / ROWS);
    CIAOSR_LAUNCH(query_mlp_kernel, tc_grid(Qp.n_tiles), TC_THREADS, SM_TOTAL, st, Qp);
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr
