// Internal entry points shared between translation units.
#pragma once
#include "common.cuh"
#include "plan.cuh"

namespace ciaosr {

// cs_attn.cu
int transpose_batched(const float* src, float* dst, int batch, int rows, int cols, cudaStream_t st);
size_t cs_attn_workspace(const PlanLayout& L, int H, int W);
int run_cs_attn(const PlanLayout& L, const float* plan, const float* featT, int B, int H, int W,
                float* out_nhwc, int ldo, float* out_nchw, void* ws, size_t ws_bytes,
                cudaStream_t st);

// cs_attn_tc.cu (tensor-core path; needs C % 4 == 0)
bool cs_attn_tc_ok(const PlanLayout& L);
size_t cs_attn_tc_workspace(const PlanLayout& L, int B, int H, int W);
int run_cs_attn_tc(const PlanLayout& L, const float* plan, const float* featT, int B, int H, int W,
                   float* out_nhwc, int ldo, float* out_nchw, void* ws, size_t ws_bytes, cudaStream_t st);

// Geometry of one head call, shared by both engines.
struct HeadArgs {
  int B, H, W, Q, eval_bsize;
  const float* featT;      // [B,H,W,C]
  const float* nlT;        // [B,H,W,Cn] or nullptr
  const float* coord;      // [B,Q,2]
  const float* cell;       // [B,Q,2]
  const float* lr;         // [B,3,H,W] or nullptr
  float* out;              // [B,Q,3]
  // make_coord constants (double -> fp32 on the host, like the reference's python scalars)
  float cy0, cy1, cx0, cx1;   // centre(i) = c0 + c1 * i
};

// head_simt.cu
size_t head_simt_workspace(const PlanLayout& L, int B, int H, int W, int Q);
int run_head_simt(const PlanLayout& L, const float* plan, const HeadArgs& a, void* ws,
                  size_t ws_bytes, cudaStream_t st);
// LR-resolution precompute shared by both engines: Pk, Pv (layer-1 hoist, no bias) and G (key fold)
// G is [B*H*W*9, ldg] with ldg >= H_last + 1 (column H_last holds the folded bias term)
int run_lr_precompute(const PlanLayout& L, const float* plan, const HeadArgs& a, float* Pk,
                      float* Pv, float* G, int ldg, cudaStream_t st);

// head_tc.cu
size_t head_tc_workspace(const PlanLayout& L, int B, int H, int W, int Q);
int run_head_tc(const PlanLayout& L, const float* plan, const HeadArgs& a, void* ws,
                size_t ws_bytes, cudaStream_t st);

}  // namespace ciaosr
