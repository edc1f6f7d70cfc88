// Plan layout + weight packing kernels (see plan.cuh for the algebra).
#include "plan.cuh"

namespace ciaosr {

static size_t take(size_t& off, size_t n) {
  off = (off + 63) / 64 * 64;          // 256-byte alignment in floats
  size_t r = off;
  off += n;
  return r;
}

static int check_mlp(const ciaosr_mlp_desc& m, const char* name, int in_dim, int out_dim) {
  CIAOSR_REQUIRE(m.n_layers >= 2 && m.n_layers <= CIAOSR_MAX_LAYERS, CIAOSR_E_INVALID,
                 "%s: n_layers=%d unsupported (need 2..%d, i.e. a non-empty hidden_list)", name,
                 m.n_layers, CIAOSR_MAX_LAYERS);
  CIAOSR_REQUIRE(m.dims[0] == in_dim, CIAOSR_E_INVALID,
                 "%s: in_dim=%d but the head geometry needs %d (ciaosr_net.py:61-76)", name,
                 m.dims[0], in_dim);
  CIAOSR_REQUIRE(m.dims[m.n_layers] == out_dim, CIAOSR_E_INVALID,
                 "%s: out_dim=%d but the head geometry needs %d", name, m.dims[m.n_layers], out_dim);
  for (int l = 0; l < m.n_layers; ++l) {
    CIAOSR_REQUIRE(m.dims[l + 1] > 0, CIAOSR_E_INVALID, "%s: layer %d has width %d", name, l,
                   m.dims[l + 1]);
    CIAOSR_REQUIRE(m.weight[l] && m.bias[l], CIAOSR_E_INVALID, "%s: layer %d weight/bias is NULL",
                   name, l);
  }
  return CIAOSR_OK;
}

int plan_layout(const ciaosr_head_desc* d, PlanLayout* L) {
  CIAOSR_REQUIRE(d != nullptr, CIAOSR_E_INVALID, "desc is NULL");
  CIAOSR_REQUIRE(d->abi_version == CIAOSR_ABI_VERSION, CIAOSR_E_INVALID,
                 "desc.abi_version=%d, library is %d", d->abi_version, CIAOSR_ABI_VERSION);
  CIAOSR_REQUIRE(d->channels > 0 && d->channels <= 1024, CIAOSR_E_INVALID, "channels=%d",
                 d->channels);
  CIAOSR_REQUIRE(d->feat_unfold == 1, CIAOSR_E_INVALID,
                 "feat_unfold=False is not implemented (every reference config sets it True)");
  CIAOSR_REQUIRE(d->local_size >= 1 && d->local_size <= 3, CIAOSR_E_INVALID,
                 "local_size=%d unsupported (1, 2 or 3)", d->local_size);
  CIAOSR_REQUIRE(d->softmax_scale != 0.0f, CIAOSR_E_INVALID, "softmax_scale is 0");
  *L = PlanLayout();
  L->C = d->channels;
  L->non_local = d->non_local_attn ? 1 : 0;
  L->local_size = d->local_size;
  L->softmax_scale = d->softmax_scale;
  L->nn = d->local_size == 1 ? 1 : (d->local_size == 2 ? 4 : 9);
  L->Cn = 0;
  if (L->non_local) {
    const ciaosr_cs_attn_desc& a = d->cs_attn;
    CIAOSR_REQUIRE(a.channels == d->channels, CIAOSR_E_INVALID, "cs_attn.channels=%d != %d",
                   a.channels, d->channels);
    CIAOSR_REQUIRE(a.n_scales == 1 && a.scales[0] == 2, CIAOSR_E_INVALID,
                   "cs_attn: only multi_scale=[2] is implemented on device (got %d scales, first=%d)",
                   a.n_scales, a.scales[0]);
    CIAOSR_REQUIRE(d->channels % 2 == 0, CIAOSR_E_INVALID, "cs_attn needs even channels");
    CIAOSR_REQUIRE(a.match1_w && a.match1_b && a.match1_slope && a.match2_w && a.match2_b &&
                       a.match2_slope && a.assembly_w && a.assembly_b && a.assembly_slope &&
                       a.down_w && a.down_b && a.escape_nan,
                   CIAOSR_E_INVALID, "cs_attn: NULL parameter pointer");
    L->Cn = d->channels * a.n_scales;
    L->cs_softmax_scale = a.softmax_scale;
  }
  L->Dk = 9 * L->C;
  L->Dv = L->Dk + L->Cn;
  int rc;
  if ((rc = check_mlp(d->imnet_k, "imnet_k", L->Dk + 4, L->Dk))) return rc;
  if ((rc = check_mlp(d->imnet_v, "imnet_v", L->Dv + 4, L->Dv))) return rc;
  if ((rc = check_mlp(d->imnet_q, "imnet_q", L->Dv, 3))) return rc;

  size_t off = 0;
  auto lay = [&](const ciaosr_mlp_desc& m, MlpPlan& p, int kind) {
    p.n_layers = m.n_layers;
    for (int l = 0; l <= m.n_layers; ++l) p.dims[l] = m.dims[l];
    for (int l = 0; l < m.n_layers; ++l) {
      int kin = m.dims[l];
      if (l == 0 && kind != 2) kin -= 4;                    // rel/cell columns live in `rc`
      p.wt[l] = take(off, (size_t)kin * m.dims[l + 1]);
      p.bias[l] = take(off, m.dims[l + 1]);
    }
    p.rc = kind != 2 ? take(off, 4 * (size_t)m.dims[1]) : 0;
    p.fin = kind == 0 ? take(off, (size_t)L->Dk * (m.dims[m.n_layers - 1] + 1)) : 0;
  };
  lay(d->imnet_k, L->k, 0);
  lay(d->imnet_v, L->v, 1);
  lay(d->imnet_q, L->q, 2);
  if (L->non_local) {
    const int C = L->C, Ch = C / 2;
    L->m1_wt = take(off, (size_t)C * Ch);  L->m1_b = take(off, Ch);
    L->m2_wt = take(off, (size_t)C * Ch);  L->m2_b = take(off, Ch);
    L->as_wt = take(off, (size_t)C * C);   L->as_b = take(off, C);
    L->down_wt = take(off, (size_t)9 * C * C);  L->down_b = take(off, C);
    L->scalars = take(off, 4);
  }
  L->tc_ok = tc_shapes_ok(d) ? 1 : 0;
  L->tc_blob_bytes = L->tc_ok ? tc_blob_bytes(d) : 0;
  L->tc_blob = take(off, (L->tc_blob_bytes + 3) / 4);
  L->total_floats = (off + 63) / 64 * 64;
  return CIAOSR_OK;
}

// ---- packing kernels --------------------------------------------------------
// source index on the unfolded axis for internal (tap-major) index kp
__device__ __forceinline__ int src_channel(int kp, int C) {
  const int Dk = 9 * C;
  return kp < Dk ? (kp % C) * 9 + kp / C : kp;
}

// dst[kp, n] = src[n, src_channel(kp)]            (src row-major [N, ld]); perm off when C == 0
__global__ void pack_transpose_kernel(float* dst, const float* src, int Kp, int N, int ld, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Kp * N) return;
  const int kp = (int)(i / N), n = (int)(i % N);
  const int ks = C > 0 ? src_channel(kp, C) : kp;
  dst[i] = src[(long long)n * ld + ks];
}
// dst[j, n] = src[n, ld - 4 + j]
__global__ void pack_rc_kernel(float* dst, const float* src, int N, int ld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * N) return;
  dst[i] = src[(long long)(i % N) * ld + (ld - 4 + i / N)];
}
// key-side fold: dst[kp, h] = W[src_channel(kp), h] (h < Hl), dst[kp, Hl] = b[src_channel(kp)]
__global__ void pack_kfin_kernel(float* dst, const float* W, const float* b, int Dk, int Hl, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Dk * (Hl + 1)) return;
  const int kp = (int)(i / (Hl + 1)), h = (int)(i % (Hl + 1));
  const int ks = src_channel(kp, C);
  dst[i] = h < Hl ? W[(long long)ks * Hl + h] : b[ks];
}
// value-side last layer: dst[h, cp] = W[src_channel(cp), h]; db[cp] = b[src_channel(cp)]
__global__ void pack_vfin_kernel(float* dst, float* db, const float* W, const float* b, int Dv,
                                 int Hl, int C) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)Dv * Hl) return;
  const int h = (int)(i / Dv), cp = (int)(i % Dv);
  const int cs = src_channel(cp, C);
  dst[i] = W[(long long)cs * Hl + h];
  if (h == 0) db[cp] = b[cs];
}
__global__ void pack_copy_kernel(float* dst, const float* src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}
// 3x3 stride-2 conv weight [co, ci, u, v] -> [(u*3+v)*C + ci, co]
__global__ void pack_down_kernel(float* dst, const float* w, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * C * C) return;
  const int co = i % C, ci = (i / C) % C, uv = i / (C * C);
  dst[i] = w[((long long)co * C + ci) * 9 + uv];
}
__global__ void pack_scalars_kernel(float* dst, const float* s1, const float* s2, const float* sa,
                                    const float* esc) {
  if (threadIdx.x == 0) { dst[0] = *s1; dst[1] = *s2; dst[2] = *sa; dst[3] = *esc; }
}

static int ew_grid(long long n) { return cdiv(n, 256); }

static int pack_mlp(const ciaosr_mlp_desc& m, const MlpPlan& p, int kind, const PlanLayout& L,
                    float* plan, cudaStream_t st) {
  const int n = m.n_layers;
  for (int l = 0; l < n; ++l) {
    const int out = m.dims[l + 1], ld = m.dims[l];
    int kin = ld, permC = 0;
    if (l == 0) { permC = L.C; if (kind != 2) kin = ld - 4; }
    const bool is_fin = (l == n - 1);
    if (is_fin && kind == 0) {
      // imnet_k last layer is consumed through `fin`; wt[l] stays unused but keep bias
      CIAOSR_LAUNCH(pack_kfin_kernel, ew_grid((long long)L.Dk * (ld + 1)), 256, 0, st,
                    plan + p.fin, m.weight[l], m.bias[l], L.Dk, ld, L.C);
    } else if (is_fin && kind == 1) {
      CIAOSR_LAUNCH(pack_vfin_kernel, ew_grid((long long)L.Dv * ld), 256, 0, st, plan + p.wt[l],
                    plan + p.bias[l], m.weight[l], m.bias[l], L.Dv, ld, L.C);
      continue;
    } else {
      CIAOSR_LAUNCH(pack_transpose_kernel, ew_grid((long long)kin * out), 256, 0, st,
                    plan + p.wt[l], m.weight[l], kin, out, ld, permC);
    }
    CIAOSR_LAUNCH(pack_copy_kernel, ew_grid(out), 256, 0, st, plan + p.bias[l], m.bias[l], out);
    if (l == 0 && kind != 2)
      CIAOSR_LAUNCH(pack_rc_kernel, ew_grid(4 * out), 256, 0, st, plan + p.rc, m.weight[l], out, ld);
  }
  return CIAOSR_OK;
}

int plan_pack(const ciaosr_head_desc* d, const PlanLayout& L, float* plan, cudaStream_t st) {
  int rc;
  if ((rc = pack_mlp(d->imnet_k, L.k, 0, L, plan, st))) return rc;
  if ((rc = pack_mlp(d->imnet_v, L.v, 1, L, plan, st))) return rc;
  if ((rc = pack_mlp(d->imnet_q, L.q, 2, L, plan, st))) return rc;
  if (L.non_local) {
    const ciaosr_cs_attn_desc& a = d->cs_attn;
    const int C = L.C, Ch = C / 2;
    CIAOSR_LAUNCH(pack_transpose_kernel, ew_grid(C * Ch), 256, 0, st, plan + L.m1_wt, a.match1_w, C, Ch, C, 0);
    CIAOSR_LAUNCH(pack_copy_kernel, ew_grid(Ch), 256, 0, st, plan + L.m1_b, a.match1_b, Ch);
    CIAOSR_LAUNCH(pack_transpose_kernel, ew_grid(C * Ch), 256, 0, st, plan + L.m2_wt, a.match2_w, C, Ch, C, 0);
    CIAOSR_LAUNCH(pack_copy_kernel, ew_grid(Ch), 256, 0, st, plan + L.m2_b, a.match2_b, Ch);
    CIAOSR_LAUNCH(pack_transpose_kernel, ew_grid(C * C), 256, 0, st, plan + L.as_wt, a.assembly_w, C, C, C, 0);
    CIAOSR_LAUNCH(pack_copy_kernel, ew_grid(C), 256, 0, st, plan + L.as_b, a.assembly_b, C);
    CIAOSR_LAUNCH(pack_down_kernel, ew_grid(9 * C * C), 256, 0, st, plan + L.down_wt, a.down_w, C);
    CIAOSR_LAUNCH(pack_copy_kernel, ew_grid(C), 256, 0, st, plan + L.down_b, a.down_b, C);
    CIAOSR_LAUNCH(pack_scalars_kernel, 1, 32, 0, st, plan + L.scalars, a.match1_slope,
                  a.match2_slope, a.assembly_slope, a.escape_nan);
  }
  if (L.tc_ok) {
    if ((rc = tc_pack(d, L, plan, st))) return rc;
  }
  return CIAOSR_OK;
}

}  // namespace ciaosr
