// tcgen05 engine for the implicit attention head (CIAOSR_ENGINE_TCGEN05), sm_100a.
//
// Persistent, warp-specialised kernels run every dense contraction of the head on the 5th-gen tensor
// cores with fp32-grade accuracy (fp16 hi/lo split, 3 UMMAs per product, fp32 accumulation in TMEM):
//
//   LR precompute     tc_gemm_kernel (gemm_tc.cuh): layer-1 hoists Pk, Pv and the key fold G
//   pair_mlp_kernel   per 128 (query, neighbour) rows = 32 queries x 4 neighbours:
//       layer 1 of imnet_k / imnet_v from the hoist (gather + 4 FMAs, ciaosr_net.py:195-205)
//       hidden layers 2..4 of both MLPs on UMMA (M=128, N=256, K=256 each)
//       key side: logit = h4k . G[query px, offset] (+c0), softmax over the 4 neighbours (:203,:214-215)
//       value side: last Linear (256 -> Dv) on UMMA in N-chunks, fused with  sum_n a_n * value_n * W_v  (:206,:215)
//       -> x[query, Dv]
//   query_mlp_kernel  per 128 queries: imnet_q (Dv -> 256 x4 -> 3) + bilinear residual (:221, :107-108)
//
// Pipeline (tc_pipeline.cuh): 1 CTA / SM, all 512 TMEM columns = two 128x256 fp32 accumulators; warp 0
// streams pre-swizzled 16 KB weight slabs through a 4-stage ring with cp.async.bulk (TMA engine), warp 1
// issues tcgen05.mma / tcgen05.commit, row threads drain accumulators with tcgen05.ld, apply bias+ReLU,
// split to fp16 hi/lo and write the next layer's A operand straight into the 128B-swizzled K-major smem
// slabs the next UMMA reads.  Slab-granular mbarriers let layer l+1's UMMAs start as soon as the first 64
// columns of layer l are converted, while the second accumulator absorbs them.  Hidden activations never
// leave the SM.  In the two head kernels CTA pairs (2-CTA clusters) share one weight stream: each CTA
// loads half of every unit and multicasts it to both, halving L2 reads.
#include "kernels.cuh"
#include "tc_head_kernels.cuh"

namespace ciaosr {

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
  t.pair_consts = off; off += 16 * HID * 4;     // immediately followed by bv5p: one contiguous const block
  t.bv5p = off; off += (size_t)t.Dvp * 4;
  t.query_consts = off; off += 8 * HID * 4;
  t.total = (off + 255) / 256 * 256;
  return t;
}

bool tc_shapes_ok(const ciaosr_head_desc* d) {
  if (d->local_size != 2 || d->channels % 4 != 0) return false;
  const int Cn = d->non_local_attn ? d->channels * d->cs_attn.n_scales : 0;
  if (9 * d->channels + Cn > 2048 - 127) return false;        // bv5p lives in the 2048-float smem const tail
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
  split_t hi, lo;
  split_scalar(w, hi, lo);
  uint8_t* ub = dst + (size_t)unit * UNIT_BYTES;
  const uint32_t off = sw128_offset(n_in, k_in);
  *reinterpret_cast<split_t*>(ub + off) = hi;
  *reinterpret_cast<split_t*>(ub + SLAB_BYTES + off) = lo;
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
// Both generators gather 8 float4 per 32-column chunk.  All addresses are formed first and every load is
// unconditional (out-of-image taps read the row's own pixel and are masked afterwards), so the 8 (or 16)
// loads of a chunk are in flight together instead of one L2 round trip per branch.
struct UnfoldGen {       // A[pix, kp] = tap-major 3x3 unfold of the NHWC feature (+ non-local channels)
  static constexpr bool kPrefetch = true;     // loads of the next slab overlap the split / store of this one (gemm_tc.cuh)
  const float* f; const float* nl; int H, W, C, Cn, K;
  struct Row { int y, x; };
  struct Raw { float4 q[8]; uint32_t ok; };
  __device__ __forceinline__ Row row(long long m) const {
    const int hw = (int)(m % ((long long)H * W));
    return Row{hw / W, hw % W};
  }
  __device__ __forceinline__ void issue(Row& r, long long m, int k0, Raw& w) const {
    const float* src[8];
    const float* self = f + m * C;
    w.ok = 0;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      src[g] = self;
      if (k < 9 * C) {
        const int t = k / C, ch = k - t * C;
        const int dy = t / 3 - 1, dx = t - (dy + 1) * 3 - 1;
        if (r.y + dy >= 0 && r.y + dy < H && r.x + dx >= 0 && r.x + dx < W) {
          src[g] = self + (dy * W + dx) * C + ch;
          w.ok |= 1u << g;
        }
      } else if (k < K) {
        w.ok |= 1u << g;
        src[g] = nl + m * Cn + (k - 9 * C);
      }
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) w.q[g] = __ldg(reinterpret_cast<const float4*>(src[g]));
  }
  __device__ __forceinline__ void finish(Row&, long long, int, const Raw& w, float (&v)[32]) const {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const bool ok = (w.ok >> g) & 1;
      v[4 * g] = ok ? w.q[g].x : 0.f; v[4 * g + 1] = ok ? w.q[g].y : 0.f;
      v[4 * g + 2] = ok ? w.q[g].z : 0.f; v[4 * g + 3] = ok ? w.q[g].w : 0.f;
    }
  }
};
struct PairProdGen {     // A[pix*9 + d, kp] = U[pix, kp] * U[pix + d, kp];  also c0 = sum_kp A * b5[kp]
  static constexpr bool kCombine = true;
  const float* f; const float* fin; int H, W, C, ldfin;     // fin[kp * ldfin + 256] = permuted last-layer bias
  struct Row { int y, x, dy, dx; long long pix; bool ok; float c0; };
  __device__ __forceinline__ float& partial(Row& r) const { return r.c0; }
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
    const float* self = f + r.pix * C;
    const int doff = r.ok ? (r.dy * W + r.dx) * C : 0;
    const float* src[8];
    const float* bsrc[8];
    bool ok[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int k = k0 + 4 * g;
      src[g] = self; ok[g] = false; bsrc[g] = fin + HID;
      if (r.ok && k < 9 * C) {
        const int t = k / C, ch = k - t * C;
        const int ty = t / 3 - 1, tx = t - (ty + 1) * 3 - 1;
        const int ay = r.y + ty, ax = r.x + tx, by = ay + r.dy, bx = ax + r.dx;
        ok[g] = ay >= 0 && ay < H && ax >= 0 && ax < W && by >= 0 && by < H && bx >= 0 && bx < W;
        if (ok[g]) { src[g] = self + (ty * W + tx) * C + ch; bsrc[g] = fin + (long long)k * ldfin + HID; }
      }
    }
    float4 a[8], b[8];
    float w5[8][4];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      a[g] = __ldg(reinterpret_cast<const float4*>(src[g]));
      b[g] = __ldg(reinterpret_cast<const float4*>(src[g] + (ok[g] ? doff : 0)));
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int i = 0; i < 4; ++i) w5[g][i] = __ldg(bsrc[g] + (ok[g] ? (long long)i * ldfin : 0));
    }
    float c0 = 0.0f, c1 = 0.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float m0 = ok[g] ? a[g].x * b[g].x : 0.f, m1 = ok[g] ? a[g].y * b[g].y : 0.f;
      const float m2 = ok[g] ? a[g].z * b[g].z : 0.f, m3 = ok[g] ? a[g].w * b[g].w : 0.f;
      v[4 * g] = m0; v[4 * g + 1] = m1; v[4 * g + 2] = m2; v[4 * g + 3] = m3;
      c0 = fmaf(m0, w5[g][0], c0); c1 = fmaf(m1, w5[g][1], c1);
      c0 = fmaf(m2, w5[g][2], c0); c1 = fmaf(m3, w5[g][3], c1);
    }
    r.c0 += c0 + c1;
  }
};
template <class Row>
struct StoreRowsEpi {    // C[m, n .. n+4) = v   (ld % 4 == 0; every column the GEMM produces exists); tile functor (gemm_tc.cuh)
  static constexpr bool kTile = true;
  float* c; int ld;
  __device__ __forceinline__ void store4(long long m, int n, float4 v) const {
    *reinterpret_cast<float4*>(c + m * ld + n) = v;
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
struct TcBufs { float *Pk, *Pv, *G; split_t *x_hi, *x_lo; long long x_rows; };
static int tc_ldg() { return HID + 4; }

// CTAs sharing one weight stream in the head kernels (CIAOSR_TC_CLUSTER=1 disables the multicast)
static int tc_cluster_size() {
  static int cl = -1;
  if (cl < 0) {
    const char* e = getenv("CIAOSR_TC_CLUSTER");
    cl = (e && atoi(e) == 1) ? 1 : 2;
  }
  return cl;
}
static int tc_grid(int n_tiles, int CL) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (const char* e = getenv("CIAOSR_DBG_MAXSMS")) {          // experiment: fewer persistent CTAs (power / shared-resource probe)
    const int m = atoi(e);
    if (m > 0 && m < sms) sms = m;
  }
  sms = sms / CL * CL;
  const int want = (n_tiles + CL - 1) / CL * CL;
  return want < sms ? want : sms;
}
// The pair-MLP stage runs as CTA pairs with cta_group::2 UMMAs (pair_mlp_pair_kernel) unless CIAOSR_HEAD_PAIR=0 (read at
// every call) selects the single-CTA kernel with multicast weights.
static bool tc_use_pair_umma() {
  const char* e = getenv("CIAOSR_HEAD_PAIR");
  return !(e && atoi(e) == 0);
}
// Experiment switches of the CTA-pair kernel, both measured slower than the default on the bench workload:
// CIAOSR_HEAD_ROWPARTS=4: four row threads per row (16 row warps) instead of two -- a slab's conversion is a chain of
// latencies every warp walks in lock-step (barrier, tcgen05.ld, convert, st.shared, proxy fence, arrive), so twice the
// warps with half the columns each convert a slab in the same 1.5 k cycles (6.91 vs 6.83 ms);
// CIAOSR_HEAD_NSPLIT=1: two N = 128 column halves per layer with the first half's epilogue under the second half's
// UMMAs -- fewer bubbles, but N = 128 UMMAs cost 106 cycles in the kernel against 2 x 74 ideal (6.91 vs 6.75 ms).
// CIAOSR_QUERY_PAIR=0: the query MLP on the single-CTA kernel even when the pair-MLP stage runs as CTA pairs
static bool tc_query_single() {
  const char* e = getenv("CIAOSR_QUERY_PAIR");
  return e && atoi(e) == 0;
}
static int tc_row_parts() {
  const char* e = getenv("CIAOSR_HEAD_ROWPARTS");
  return (e && atoi(e) == 4) ? 4 : 2;
}
static bool tc_nsplit() {
  const char* e = getenv("CIAOSR_HEAD_NSPLIT");
  return e && atoi(e) == 1;
}
// CIAOSR_TC_TERMS (read at every call): product terms of the pair / query MLP jobs.  7 (default) = fp16 hi/lo split with three
// UMMAs per product (fp32-grade, the only mode inside the parity tolerance); 3 = A.W_hi (weights at 11 bits); 2 = A_hi.W_hi
// (single fp16 pass).  The reduced modes exist to document the cost / accuracy frontier on hardware; they
// apply to the pair / query MLP kernels only (the LR precompute, cross-scale attention and encoder keep all terms).
static unsigned tc_terms() {
  const char* e = getenv("CIAOSR_TC_TERMS");
  const int v = e ? atoi(e) : 7;
  return (v == 2 || v == 3) ? (unsigned)v : 7u;
}
// CIAOSR_HEAD_FUSED=1 (read at every call) selects head_fused_kernel: x stays in a per-CTA, L2-resident scratch block and
// the workspace no longer grows with the number of queries, at ~5 % more time than the two pipelined kernels (CTA-pair
// form head_fused_pair_kernel: 7.5 vs 7.1 ms on the bench workload; see the kernels' headers); ignored when its constants do not fit beside the operand slabs (very wide heads).
// Without the variable the fused kernel is chosen automatically when the two-kernel path's x buffer would exceed
// CIAOSR_X_WORKSPACE_GIB (default 16 GiB): an un-tiled x8 call producing a 4K frame needs 21 GB of x, an 8K frame 85 GB --
// where the reference's eval_bsize loop keeps memory bounded, this engine switches to its O(1)-workspace kernel instead of
// failing with an out-of-memory error.  CIAOSR_HEAD_FUSED=0 forces the two-kernel path.
static bool tc_use_fused(int Dvp, long long total_q) {
  if (fused_smem_bytes(Dvp) > 227 * 1024) return false;
  if (const char* e = getenv("CIAOSR_HEAD_FUSED")) return atoi(e) == 1;
  double limit_gib = 16.0;
  if (const char* e = getenv("CIAOSR_X_WORKSPACE_GIB")) limit_gib = atof(e);
  return (double)total_q * Dvp * 2.0 * sizeof(split_t) > limit_gib * 1073741824.0;
}

static TcBufs tc_carve_ws(Arena& a, const PlanLayout& L, int B, int H, int W, int Q) {
  TcBufs s;
  const TcLayout t = tc_layout(L.C, L.Cn);
  const size_t npix = (size_t)B * H * W;
  s.Pk = a.take<float>(npix * HID);
  s.Pv = a.take<float>(npix * HID);
  s.G = a.take<float>(npix * 9 * tc_ldg());
  // attended values, fp16 hi / lo halves: the whole call for the two-kernel path, one 128-row block per CTA when fused
  const long long total_q = (long long)B * Q;
  s.x_rows = total_q;
  if (tc_use_fused(t.Dvp, total_q)) {
    const int n_super = (int)((total_q + ROWS - 1) / ROWS);
    s.x_rows = (long long)tc_grid(n_super, tc_use_pair_umma() ? 2 : tc_cluster_size()) * ROWS;   // one block per CTA (head_fused_*kernel)
  }
  s.x_hi = a.take<split_t>((size_t)s.x_rows * t.Dvp);
  s.x_lo = a.take<split_t>((size_t)s.x_rows * t.Dvp);
  return s;
}

size_t head_tc_workspace(const PlanLayout& L, int B, int H, int W, int Q) {
  Arena a(nullptr, 0);
  tc_carve_ws(a, L, B, H, W, Q);
  return a.used();
}

template <class Kernel, class... Args>
static int launch_clustered_t(Kernel kernel, int grid, int CL, int threads, int smem_bytes, cudaStream_t st,
                              const Args&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, args...);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  if (err != cudaSuccess) {
    set_error("launch of a clustered tcgen05 kernel failed: %s", cudaGetErrorString(err));
    return CIAOSR_E_CUDA;
  }
  return CIAOSR_OK;
}

template <class Kernel, class... Args>
static int launch_clustered(Kernel kernel, int grid, int CL, int smem_bytes, cudaStream_t st, const Args&... args) {
  return launch_clustered_t(kernel, grid, CL, HEAD_THREADS, smem_bytes, st, args...);
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
  const int CL = tc_cluster_size();
  const long long total_q = (long long)a.B * a.Q;
  PairParams P;
  P.pc = PairConsts{a.H, a.W, a.Q, a.eval_bsize, L.local_size, a.cy0, a.cy1, a.cx0, a.cx1};
  P.coord = a.coord; P.cell = a.cell; P.featT = a.featT; P.nlT = a.nlT;
  P.C = L.C; P.Cn = L.Cn; P.Dv = L.Dv; P.Dvp = t.Dvp;
  P.Pk = b.Pk; P.Pv = b.Pv; P.G = b.G; P.ldg = tc_ldg();
  P.consts = reinterpret_cast<const float*>(blob + t.pair_consts);
  P.blob = blob + t.pair_blob; P.units_per_tile = t.pair_units; P.units5 = t.units5;
  P.x_hi = b.x_hi; P.x_lo = b.x_lo; P.total_rows = total_q * 4; P.n_tiles = (int)((P.total_rows + ROWS - 1) / ROWS);
  P.softmax_scale = L.softmax_scale;
  P.terms = tc_terms();
  QueryParams Qp;
  Qp.Dvp = t.Dvp;
  Qp.consts = reinterpret_cast<const float*>(blob + t.query_consts);
  Qp.blob = blob + t.query_blob; Qp.units_per_tile = t.query_units; Qp.slabs1 = t.slabs1;
  Qp.lr = a.lr; Qp.coord = a.coord; Qp.H = a.H; Qp.W = a.W; Qp.Q = a.Q;
  Qp.terms = tc_terms();
  Qp.out = a.out; Qp.total_q = total_q; Qp.n_tiles = (int)((total_q + ROWS - 1) / ROWS);
  CUtensorMap map_hi, map_lo;
  if ((rc = tma_make_map_2d(&map_hi, b.x_hi, b.x_rows, t.Dvp)) || (rc = tma_make_map_2d(&map_lo, b.x_lo, b.x_rows, t.Dvp)))
    return rc;
  if (tc_use_fused(t.Dvp, total_q)) {
    // pair tiles and the query tile of the same 128 queries in one persistent CTA; x through an L2-resident scratch block.
    // The stage timer attributes the whole kernel to the pair stage (the query MLP is ~9 % of its tensor work).
    StageScope sc(3, st);
    static DynSmemOptIn optin[2];
    const int smem_bytes = fused_smem_bytes(t.Dvp);
    if ((rc = optin[0].ensure(head_fused_kernel<1>, smem_bytes)) || (rc = optin[1].ensure(head_fused_kernel<2>, smem_bytes)))
      return rc;
    if (tc_use_pair_umma()) {                 // CTA pairs: cta_group::2 UMMAs in both phases
      static DynSmemOptIn optin_fp;
      if ((rc = optin_fp.ensure(head_fused_pair_kernel, smem_bytes))) return rc;
      CUtensorMap wmap_p, wmap_q;
      if ((rc = tma_make_map_linear_rows(&wmap_p, blob + t.pair_blob, (long long)t.pair_units * 2 * ROWS)) ||
          (rc = tma_make_map_linear_rows(&wmap_q, blob + t.query_blob, (long long)t.query_units * 2 * ROWS)))
        return rc;
      const int grid = tc_grid(Qp.n_tiles, 2);
      Qp.iters = (Qp.n_tiles + grid - 1) / grid;
      P.iters = Qp.iters * 4;
      return launch_clustered(head_fused_pair_kernel, grid, 2, smem_bytes, st, P, Qp, map_hi, map_lo, wmap_p, wmap_q);
    }
    const int grid = tc_grid(Qp.n_tiles, CL);
    Qp.iters = (Qp.n_tiles + grid - 1) / grid;
    P.iters = Qp.iters * 4;
    return CL == 2 ? launch_clustered(head_fused_kernel<2>, grid, 2, smem_bytes, st, P, Qp, map_hi, map_lo)
                   : launch_clustered(head_fused_kernel<1>, grid, 1, smem_bytes, st, P, Qp, map_hi, map_lo);
  }
  static DynSmemOptIn optin[4];       // per kernel, per device (common.cuh)
  if ((rc = optin[0].ensure(pair_mlp_kernel<1>, SM_TOTAL)) || (rc = optin[1].ensure(pair_mlp_kernel<2>, SM_TOTAL)) ||
      (rc = optin[2].ensure(query_mlp_kernel<1>, SM_TOTAL)) || (rc = optin[3].ensure(query_mlp_kernel<2>, SM_TOTAL)))
    return rc;
  if (tc_use_pair_umma()) {
    StageScope sc(3, st);
    static DynSmemOptIn optin_pair[4];
    if ((rc = optin_pair[0].ensure(pair_mlp_pair_kernel<2, false>, SM_TOTAL)) ||
        (rc = optin_pair[1].ensure(pair_mlp_pair_kernel<4, false>, SM_TOTAL)) ||
        (rc = optin_pair[2].ensure(pair_mlp_pair_kernel<2, true>, SM_TOTAL)) ||
        (rc = optin_pair[3].ensure(pair_mlp_pair_kernel<4, true>, SM_TOTAL)))
      return rc;
    CUtensorMap wmap;
    if ((rc = tma_make_map_linear_rows(&wmap, blob + t.pair_blob, (long long)t.pair_units * 2 * ROWS))) return rc;
    const int grid = tc_grid(P.n_tiles, 2);
    P.iters = (P.n_tiles + grid - 1) / grid;
    if (tc_nsplit())
      rc = tc_row_parts() == 4 ? launch_clustered_t(pair_mlp_pair_kernel<4, true>, grid, 2, 640, SM_TOTAL, st, P, wmap)
                               : launch_clustered_t(pair_mlp_pair_kernel<2, true>, grid, 2, 384, SM_TOTAL, st, P, wmap);
    else
      rc = tc_row_parts() == 4 ? launch_clustered_t(pair_mlp_pair_kernel<4, false>, grid, 2, 640, SM_TOTAL, st, P, wmap)
                               : launch_clustered_t(pair_mlp_pair_kernel<2, false>, grid, 2, 384, SM_TOTAL, st, P, wmap);
    if (rc) return rc;
  } else {
    StageScope sc(3, st);
    const int grid = tc_grid(P.n_tiles, CL);
    P.iters = (P.n_tiles + grid - 1) / grid;
    int rc2 = CL == 2 ? launch_clustered(pair_mlp_kernel<2>, grid, 2, SM_TOTAL, st, P)
                      : launch_clustered(pair_mlp_kernel<1>, grid, 1, SM_TOTAL, st, P);
    if (rc2) return rc2;
  }
  if (tc_use_pair_umma() && !tc_query_single()) {
    StageScope sc(4, st);
    static DynSmemOptIn optin_qpair;
    if ((rc = optin_qpair.ensure(query_mlp_pair_kernel, SM_TOTAL))) return rc;
    CUtensorMap qwmap;
    if ((rc = tma_make_map_linear_rows(&qwmap, blob + t.query_blob, (long long)t.query_units * 2 * ROWS))) return rc;
    const int grid = tc_grid(Qp.n_tiles, 2);
    Qp.iters = (Qp.n_tiles + grid - 1) / grid;
    if ((rc = launch_clustered(query_mlp_pair_kernel, grid, 2, SM_TOTAL, st, Qp, map_hi, map_lo, qwmap))) return rc;
  } else {
    StageScope sc(4, st);
    const int grid = tc_grid(Qp.n_tiles, CL);
    Qp.iters = (Qp.n_tiles + grid - 1) / grid;
    int rc2 = CL == 2 ? launch_clustered(query_mlp_kernel<2>, grid, 2, SM_TOTAL, st, Qp, map_hi, map_lo)
                      : launch_clustered(query_mlp_kernel<1>, grid, 1, SM_TOTAL, st, Qp, map_hi, map_lo);
    if (rc2) return rc2;
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr

#ifdef CIAOSR_TC_TIMING
extern "C" int ciaosr_debug_trace(int on, unsigned long long* out, unsigned int* n) {
  cudaDeviceSynchronize();
  if (out) {
    cudaMemcpyFromSymbol(n, ciaosr::tc::g_trace_n, 4);
    cudaMemcpyFromSymbol(out, ciaosr::tc::g_trace, 2 * 8192 * 8);
  }
  unsigned int z = 0;
  cudaMemcpyToSymbol(ciaosr::tc::g_trace_n, &z, 4);
  cudaMemcpyToSymbol(ciaosr::tc::g_trace_req, &on, 4);
  return 0;
}
extern "C" int ciaosr_debug_flags(int flags) {
  cudaDeviceSynchronize();
  cudaMemcpyToSymbol(ciaosr::tc::g_dbg_flags, &flags, 4);
  return 0;
}
// diagnostic build only: cycles spent in mbarrier waits by the kernels of this translation unit
extern "C" int ciaosr_debug_wait_read_head(unsigned long long* cycles, unsigned long long* counts, int reset) {
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
