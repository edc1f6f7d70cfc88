// CUDA-core fp32 engine for the implicit attention head (CIAOSR_ENGINE_SIMT).
//
// Any-shape path and on-device fp32 cross-check of the tcgen05 engine.  The
// MLP layers run through the generic functor GEMM with activations in the
// workspace (HBM); gathers, concatenations and the unfold are index math inside
// the loaders.  Stages (reference lines: ciaosr_net.py)
//   LR precompute   layer-1 hoist of imnet_k / imnet_v (:195-205 first Linear), key fold (:203,214)
//   pair_layer1     local-ensemble encoding (:159-193) + first ReLU
//   hidden GEMMs    MLPRefiner hidden layers (mlp_refiner.py:74-86)
//   pair_logits     query . (key * W_k)  (:203, :214)
//   value GEMM      last Linear of imnet_v (:205)
//   attend          softmax over neighbours and weighted sum of value * W_v (:206, :215)
//   query GEMMs     imnet_q (:221) + bilinear residual (:107-108)
#include "gemm_simt.cuh"
#include "kernels.cuh"
#include "pairs.cuh"

namespace ciaosr {

// ---- LR-resolution loaders ------------------------------------------------------
struct UnfoldA {       // A[pix, kp]: tap-major unfolded feature (+ non-local channels)
  const float* f; const float* nl; int H, W, C, Cn;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return value_at(f, nl, m, k, H, W, C, Cn);
  }
};
struct PairProdA {     // A[pix*9 + d, kp] = U[pix, kp] * U[pix + d, kp]   (kp < 9C)
  const float* f; int H, W, C;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int pix = m / 9, d = m % 9;
    const int hw = pix % (H * W);
    const int y = hw / W + d / 3 - 1, x = hw % W + d % 3 - 1;
    if (y < 0 || y >= H || x < 0 || x >= W) return 0.0f;
    const int npix = pix + (d / 3 - 1) * W + (d % 3 - 1);
    return value_at(f, nullptr, pix, k, H, W, C, 0) * value_at(f, nullptr, npix, k, H, W, C, 0);
  }
};

int run_lr_precompute(const PlanLayout& L, const float* plan, const HeadArgs& a, float* Pk,
                      float* Pv, float* G, int ldg, cudaStream_t st) {
  const int npix = a.B * a.H * a.W;
  const int H1k = L.k.dims[1], H1v = L.v.dims[1], Hlk = L.k.dims[L.k.n_layers - 1];
  int rc;
  UnfoldA ua{a.featT, a.nlT, a.H, a.W, L.C, L.Cn};
  if ((rc = gemm_simt(npix, H1k, L.Dk, ua, RowMajorB{plan + L.k.wt[0], H1k},
                      EpiBiasAct{Pk, H1k, nullptr, 0}, st))) return rc;
  if ((rc = gemm_simt(npix, H1v, L.Dv, ua, RowMajorB{plan + L.v.wt[0], H1v},
                      EpiBiasAct{Pv, H1v, nullptr, 0}, st))) return rc;
  if ((rc = gemm_simt(npix * 9, Hlk + 1, L.Dk, PairProdA{a.featT, a.H, a.W, L.C},
                      RowMajorB{plan + L.k.fin, Hlk + 1}, EpiBiasAct{G, ldg, nullptr, 0}, st)))
    return rc;
  return CIAOSR_OK;
}

// ---- per-chunk kernels ------------------------------------------------------------
constexpr int L1_ROWS = 16;

// rows [r0, r0+nrows) of the global (query, neighbour) list
__global__ void __launch_bounds__(256)
pair_layer1_kernel(PairConsts pc, const float* __restrict__ coord, const float* __restrict__ cell,
                   long long g0, int nn, int nrows,
                   const float* __restrict__ Pk, const float* __restrict__ Pv,
                   const float* __restrict__ rck, const float* __restrict__ bk, int H1k,
                   const float* __restrict__ rcv, const float* __restrict__ bv, int H1v,
                   float* __restrict__ hk, float* __restrict__ hv, int* __restrict__ pair_pix,
                   int* __restrict__ pair_g) {
  __shared__ PairInfo info[L1_ROWS];
  const int rb = blockIdx.x * L1_ROWS;
  if (threadIdx.x < L1_ROWS && rb + threadIdx.x < nrows) {
    const int r = rb + threadIdx.x;
    const PairInfo p = compute_pair(pc, coord, cell, g0 + r / nn, r % nn);
    info[threadIdx.x] = p;
    pair_pix[r] = p.pix;
    pair_g[r] = p.gidx;
  }
  __syncthreads();
  const int rows = min(L1_ROWS, nrows - rb);
  for (int i = 0; i < rows; ++i) {
    const PairInfo p = info[i];
    const long long r = rb + i;
    for (int h = threadIdx.x; h < H1k; h += blockDim.x) {
      float v = (p.pix >= 0 ? Pk[(long long)p.pix * H1k + h] : 0.0f) + bk[h];
      v = fmaf(rck[h], p.rel_y, v);
      v = fmaf(rck[H1k + h], p.rel_x, v);
      v = fmaf(rck[2 * H1k + h], p.sc_y, v);
      v = fmaf(rck[3 * H1k + h], p.sc_x, v);
      hk[r * H1k + h] = fmaxf(v, 0.0f);
    }
    for (int h = threadIdx.x; h < H1v; h += blockDim.x) {
      float v = (p.pix >= 0 ? Pv[(long long)p.pix * H1v + h] : 0.0f) + bv[h];
      v = fmaf(rcv[h], p.rel_y, v);
      v = fmaf(rcv[H1v + h], p.rel_x, v);
      v = fmaf(rcv[2 * H1v + h], p.sc_y, v);
      v = fmaf(rcv[3 * H1v + h], p.sc_x, v);
      hv[r * H1v + h] = fmaxf(v, 0.0f);
    }
  }
}

// one warp per row: logit = h . G[gidx, :Hl] + G[gidx, Hl]
__global__ void pair_logits_kernel(const float* __restrict__ hk, const float* __restrict__ G,
                                   const int* __restrict__ pair_g, int nrows, int Hl, int ldg,
                                   float* __restrict__ logits) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nrows) return;
  const int lane = threadIdx.x & 31;
  const int g = pair_g[r];
  float s = 0.0f;
  if (g >= 0) {
    const float* gr = G + (long long)g * ldg;
    const float* hr = hk + (long long)r * Hl;
    for (int h = lane; h < Hl; h += 32) s = fmaf(hr[h], gr[h], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += gr[Hl];
  }
  if (lane == 0) logits[r] = s;
}

// one block per query: softmax over the nn logits, x = sum_n a_n * value_n * Wv_n
__global__ void __launch_bounds__(128)
attend_kernel(const float* __restrict__ logits, const float* __restrict__ wv,
              const int* __restrict__ pair_pix, const float* __restrict__ featT,
              const float* __restrict__ nlT, int nn, int H, int W, int C, int Cn, int Dv,
              float softmax_scale, float* __restrict__ x) {
  const long long qi = blockIdx.x;
  float a[9];
  int pix[9];
  float mx = -INFINITY;
  for (int n = 0; n < nn; ++n) {
    a[n] = __fdiv_rn(logits[qi * nn + n], softmax_scale);
    pix[n] = pair_pix[qi * nn + n];
    mx = fmaxf(mx, a[n]);
  }
  float sum = 0.0f;
  for (int n = 0; n < nn; ++n) { a[n] = expf(a[n] - mx); sum += a[n]; }
  for (int n = 0; n < nn; ++n) a[n] = __fdiv_rn(a[n], sum);
  for (int cp = threadIdx.x; cp < Dv; cp += blockDim.x) {
    float acc = 0.0f;
    for (int n = 0; n < nn; ++n) {
      const float v = value_at(featT, nlT, pix[n], cp, H, W, C, Cn);
      acc = fmaf(a[n], v * wv[(qi * nn + n) * Dv + cp], acc);
    }
    x[qi * Dv + cp] = acc;
  }
}

struct EpiRgbOut {     // out[g, n] = acc + b[n] (+ bilinear residual of the LR image)
  float* out; const float* bias; const float* lr; const float* coord; long long g0;
  int H, W, Q;
  __device__ __forceinline__ void operator()(int m, int n, float acc) const {
    const long long g = g0 + m;
    float v = acc + bias[n];
    if (lr) {
      const int b = (int)(g / Q);
      v += bilinear_border(lr + ((long long)b * 3 + n) * H * W, H, W, coord[g * 2], coord[g * 2 + 1]);
    }
    out[g * 3 + n] = v;
  }
};

// ---- host orchestration -----------------------------------------------------------
static int simt_chunk(long long total) { return (int)(total < 32768 ? total : 32768); }

static int simt_ldg(const PlanLayout& L) { return (L.k.dims[L.k.n_layers - 1] + 1 + 3) / 4 * 4; }

static int max_hidden(const MlpPlan& p) {
  int m = 0;
  for (int l = 1; l < p.n_layers; ++l) m = m > p.dims[l] ? m : p.dims[l];
  return m;
}

struct SimtBufs {
  float *Pk, *Pv, *G, *bufA, *bufB, *bufC, *logits, *wv, *x;
  int *pair_pix, *pair_g;
};

static SimtBufs simt_carve(Arena& a, const PlanLayout& L, int B, int H, int W, int Q) {
  SimtBufs s;
  const size_t npix = (size_t)B * H * W;
  const int QC = simt_chunk((long long)B * Q);
  const size_t rows = (size_t)QC * L.nn;
  int hm = max_hidden(L.k);
  hm = hm > max_hidden(L.v) ? hm : max_hidden(L.v);
  hm = hm > max_hidden(L.q) ? hm : max_hidden(L.q);
  s.Pk = a.take<float>(npix * L.k.dims[1]);
  s.Pv = a.take<float>(npix * L.v.dims[1]);
  s.G = a.take<float>(npix * 9 * simt_ldg(L));
  s.bufA = a.take<float>(rows * hm);
  s.bufB = a.take<float>(rows * hm);
  s.bufC = a.take<float>(rows * hm);
  s.logits = a.take<float>(rows);
  s.wv = a.take<float>(rows * L.Dv);
  s.x = a.take<float>((size_t)QC * L.Dv);
  s.pair_pix = a.take<int>(rows);
  s.pair_g = a.take<int>(rows);
  return s;
}

size_t head_simt_workspace(const PlanLayout& L, int B, int H, int W, int Q) {
  Arena a(nullptr, 0);
  simt_carve(a, L, B, H, W, Q);
  return a.used();
}

int run_head_simt(const PlanLayout& L, const float* plan, const HeadArgs& a, void* ws,
                  size_t ws_bytes, cudaStream_t st) {
  Arena ar(ws, ws_bytes);
  SimtBufs s = simt_carve(ar, L, a.B, a.H, a.W, a.Q);
  CIAOSR_REQUIRE(ar.ok, CIAOSR_E_WORKSPACE, "head (SIMT) workspace too small: need %zu, have %zu",
                 ar.used(), ws_bytes);
  int rc;
  {
    StageScope sc(2, st);
    if ((rc = run_lr_precompute(L, plan, a, s.Pk, s.Pv, s.G, simt_ldg(L), st))) return rc;
  }

  const long long total = (long long)a.B * a.Q;
  const int QC = simt_chunk(total);
  const int H1k = L.k.dims[1], H1v = L.v.dims[1], Hlk = L.k.dims[L.k.n_layers - 1];
  PairConsts pc{a.H, a.W, a.Q, a.eval_bsize, L.local_size, a.cy0, a.cy1, a.cx0, a.cx1};
  for (long long g0 = 0; g0 < total; g0 += QC) {
    const int nq = (int)((total - g0) < QC ? (total - g0) : QC);
    const int rows = nq * L.nn;
    StageScope* sc = new StageScope(3, st);
    struct ScopeGuard { StageScope*& p; ~ScopeGuard() { delete p; } } guard{sc};
    float* hk = s.bufA;
    float* hv = s.bufB;
    // layer 1 of both chains is produced together: hk in bufA, hv in bufB (live until the value chain)
    CIAOSR_LAUNCH(pair_layer1_kernel, cdiv(rows, L1_ROWS), 256, 0, st, pc, a.coord, a.cell, g0, L.nn,
                  rows, s.Pk, s.Pv, plan + L.k.rc, plan + L.k.bias[0], H1k, plan + L.v.rc,
                  plan + L.v.bias[0], H1v, hk, hv, s.pair_pix, s.pair_g);
    // key chain: hidden layers 1 .. n-2 (ping-pong bufA <-> bufC)
    float* cur = hk;
    float* other = s.bufC;
    for (int l = 1; l < L.k.n_layers - 1; ++l) {
      if ((rc = gemm_simt(rows, L.k.dims[l + 1], L.k.dims[l], RowMajorA{cur, L.k.dims[l]},
                          RowMajorB{plan + L.k.wt[l], L.k.dims[l + 1]},
                          EpiBiasAct{other, L.k.dims[l + 1], plan + L.k.bias[l], 1}, st))) return rc;
      float* t = cur; cur = other; other = t;
    }
    CIAOSR_LAUNCH(pair_logits_kernel, cdiv(rows, 8), 256, 0, st, cur, s.G, s.pair_g, rows, Hlk, simt_ldg(L), s.logits);
    // value chain: hidden layers in bufB <-> bufA, last layer -> wv
    cur = hv;
    other = s.bufA;
    for (int l = 1; l < L.v.n_layers - 1; ++l) {
      if ((rc = gemm_simt(rows, L.v.dims[l + 1], L.v.dims[l], RowMajorA{cur, L.v.dims[l]},
                          RowMajorB{plan + L.v.wt[l], L.v.dims[l + 1]},
                          EpiBiasAct{other, L.v.dims[l + 1], plan + L.v.bias[l], 1}, st))) return rc;
      float* t = cur; cur = other; other = t;
    }
    {
      const int l = L.v.n_layers - 1;
      if ((rc = gemm_simt(rows, L.Dv, L.v.dims[l], RowMajorA{cur, L.v.dims[l]},
                          RowMajorB{plan + L.v.wt[l], L.Dv},
                          EpiBiasAct{s.wv, L.Dv, plan + L.v.bias[l], 0}, st))) return rc;
    }
    CIAOSR_LAUNCH(attend_kernel, nq, 128, 0, st, s.logits, s.wv, s.pair_pix, a.featT, a.nlT, L.nn,
                  a.H, a.W, L.C, L.Cn, L.Dv, L.softmax_scale, s.x);
    // query MLP
    delete sc;
    sc = new StageScope(4, st);
    cur = s.x;
    float* pp[2] = {s.bufA, s.bufB};
    for (int l = 0; l < L.q.n_layers - 1; ++l) {
      float* dst = pp[l & 1];
      if ((rc = gemm_simt(nq, L.q.dims[l + 1], L.q.dims[l], RowMajorA{cur, L.q.dims[l]},
                          RowMajorB{plan + L.q.wt[l], L.q.dims[l + 1]},
                          EpiBiasAct{dst, L.q.dims[l + 1], plan + L.q.bias[l], 1}, st))) return rc;
      cur = dst;
    }
    {
      const int l = L.q.n_layers - 1;
      if ((rc = gemm_simt(nq, 3, L.q.dims[l], RowMajorA{cur, L.q.dims[l]},
                          RowMajorB{plan + L.q.wt[l], 3},
                          EpiRgbOut{a.out, plan + L.q.bias[l], a.lr, a.coord, g0, a.H, a.W, a.Q}, st)))
        return rc;
    }
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr
