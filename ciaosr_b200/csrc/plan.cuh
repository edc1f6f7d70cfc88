// Plan = the head's weights re-laid-out for the kernels (device memory owned by
// the caller).  The layout is a pure function of the descriptor, recomputed on
// the host at every call; nothing but floats lives in the plan buffer.
//
// Internal channel order.  The reference's unfolded feature uses channel
// c = ch*9 + t (F.unfold, ciaosr_net.py:132; t = ki*3 + kj).  Kernels here use
// the tap-major order c' = t*C + ch, so that one tap of one LR pixel is a
// contiguous run of C floats in the NHWC feature map; non-local channels keep
// their place after the 9C unfolded ones.  Every weight matrix that touches
// that axis is permuted once, here.
//
// Algebra done at plan time (exact in real arithmetic; fp32 reassociation only)
//  * layer-1 hoist:  Linear_1(cat[unfold(f)[p], rel, cell])
//                  = (W1[:, :D] . unfold(f))[p] + W1[:, D:] . [rel, cell] + b1
//    so W1 is split into `w1` ([D', H1], applied once per LR pixel) and
//    `rc` ([4, H1], applied per (query, neighbour)).
//  * key-side last layer folded into the logit:
//      logit = sum_c q_c k_c (W_L h + b_L)_c = h . (W_L^T (q.k)) + b_L . (q.k)
//    `kfin` is [D', H_last + 1]: W_L rows in tap-major order with b_L appended
//    as one more column; it is applied once per (LR pixel, neighbour offset).
#pragma once
#include "common.cuh"

namespace ciaosr {

struct MlpPlan {
  int n_layers;                      // Linear layers
  int dims[CIAOSR_MAX_LAYERS + 1];
  // offsets (in floats) into the plan buffer
  size_t wt[CIAOSR_MAX_LAYERS];      // [dims[l], dims[l+1]] (k-major rows, n contiguous)
  size_t bias[CIAOSR_MAX_LAYERS];    // [dims[l+1]]
  size_t rc;                         // [4, dims[1]]  (k and v only)
  size_t fin;                        // k only: [Dk, H_last + 1]
};

struct PlanLayout {
  int C, Cn, nn, Dk, Dv;             // channels, non-local channels, neighbours
  int local_size, non_local;
  float softmax_scale;
  MlpPlan k, v, q;
  // cross-scale attention
  size_t m1_wt, m1_b, m2_wt, m2_b, as_wt, as_b, down_wt, down_b, scalars;  // scalars: slope1, slope2, slopeA, escape
  float cs_softmax_scale;
  // tcgen05 blobs (head_tc.cu); 0 when the shapes do not admit that engine
  int tc_ok;
  size_t tc_blob;                    // offset in floats (16-byte aligned)
  size_t tc_blob_bytes;
  size_t total_floats;
};

int plan_layout(const ciaosr_head_desc* d, PlanLayout* L);
int plan_pack(const ciaosr_head_desc* d, const PlanLayout& L, float* plan, cudaStream_t st);

// tcgen05 engine hooks (head_tc.cu)
bool tc_shapes_ok(const ciaosr_head_desc* d);
size_t tc_blob_bytes(const ciaosr_head_desc* d);
int tc_pack(const ciaosr_head_desc* d, const PlanLayout& L, float* plan, cudaStream_t st);

}  // namespace ciaosr
