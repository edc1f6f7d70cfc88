/*
 * ciaosr_b200.h -- C ABI of the B200-native CiaoSR implicit attention head.
 *
 * This is the drop-in boundary for the ONE hot path this repository
 * accelerates: everything the reference does between
 * "encoder feature is available" and "[B,Q,3] prediction is returned".
 * All pointers are plain DEVICE pointers (fp32 unless stated), all sizes are
 * plain ints, streams are passed as `void*` (a `cudaStream_t`).  No torch
 * types, no C++ types.  Paths below are relative to the reference root.
 *
 * Reference interface each entry point replaces
 * ---------------------------------------------
 *  ciaosr_cross_scale_attn_forward
 *      CrossScaleAttention.forward(input[B,C,H,W]) -> [B,C*len(scale),H,W]
 *      mmedited/models/common/arch_csnln.py:430-532
 *      (called from mmedited/models/backbones/sr_backbones/ciaosr_net.py:135)
 *  ciaosr_query_rgb_forward
 *      LocalImplicitSRNet.query_rgb(features, coord, cell) -> [B,q,3]
 *      ciaosr_net.py:113-224, plus the query-axis chunk loop
 *      LocalImplicitSRNet.batched_predict (ciaosr_net.py:226-248) when
 *      eval_bsize > 0, plus (optionally) the bilinear residual that
 *      LocalImplicitSRNet.forward adds at ciaosr_net.py:107-108.
 *  ciaosr_plan_bytes / ciaosr_plan_init
 *      the weight tensors the reference reads from nn.Parameters named
 *      imnet_{q,k,v}.layers.{0,2,..}.{weight,bias} (MLPRefiner,
 *      mmedited/models/components/refiners/mlp_refiner.py:65-102) and
 *      cs_attn.{conv_match_1,conv_match_2,conv_assembly}.{0.weight,0.bias,
 *      1.weight}, cs_attn.down.{weight,bias}, cs_attn.escape_NaN
 *      (arch_csnln.py:407-428), re-laid-out once for the kernels.
 *  ciaosr_workspace_bytes
 *      replaces the intermediates eager PyTorch allocates implicitly.
 *  ciaosr_tile_blend_accumulate / ciaosr_tile_blend_finish
 *      the E/W overlap-average of CiaoSR.clip_test and the de-normalise +
 *      clamp of CiaoSR.forward_test
 *      (mmedited/models/restorers/ciaosr.py:253-256, 160-163).
 *
 * Error convention: every function returns 0 on success or a negative
 * CIAOSR_E_* code; ciaosr_last_error() returns a thread-local description.
 * Nothing falls back to a CPU path: with no usable device the calls fail.
 *
 * Threading / streams: functions enqueue work on `stream` and return without
 * synchronising (ciaosr_plan_init included).  A plan is immutable after
 * ciaosr_plan_init; concurrent forwards may share a plan but not a workspace.
 */
#ifndef CIAOSR_B200_H_
#define CIAOSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIAOSR_ABI_VERSION 1
#define CIAOSR_MAX_LAYERS 8
#define CIAOSR_MAX_SCALES 4

enum {
  CIAOSR_OK = 0,
  CIAOSR_E_INVALID = -1,     /* bad argument / unsupported configuration        */
  CIAOSR_E_WORKSPACE = -2,   /* workspace or plan buffer too small / misaligned */
  CIAOSR_E_CUDA = -3,        /* a CUDA runtime call or kernel launch failed     */
  CIAOSR_E_NO_DEVICE = -4    /* no sm_100 device available                      */
};

/* Which implementation runs the MLP stacks of the head. */
enum {
  CIAOSR_ENGINE_AUTO = 0,    /* tcgen05 when the shapes admit it, else SIMT      */
  CIAOSR_ENGINE_SIMT = 1,    /* fp32 CUDA-core path (any shape)                  */
  CIAOSR_ENGINE_TCGEN05 = 2  /* sm_100a tensor-core path (hidden = 256 x 4)      */
};

/* One MLPRefiner (mlp_refiner.py:65-102): Linear+ReLU ... Linear.
 * weight[i] is row-major [dims[i+1], dims[i]] exactly as in the state_dict. */
typedef struct ciaosr_mlp_desc {
  int32_t n_layers;                         /* number of Linear layers            */
  int32_t dims[CIAOSR_MAX_LAYERS + 1];      /* dims[0]=in_dim ... dims[n]=out_dim */
  const float* weight[CIAOSR_MAX_LAYERS];
  const float* bias[CIAOSR_MAX_LAYERS];
} ciaosr_mlp_desc;

/* CrossScaleAttention parameters (arch_csnln.py:407-428). conv weights are
 * [out,in,1,1] / [out,in,3,3] row-major as in the state_dict; *_slope is the
 * single-element PReLU weight. */
typedef struct ciaosr_cs_attn_desc {
  int32_t channels;                         /* C                                  */
  int32_t n_scales;                         /* len(multi_scale)                   */
  int32_t scales[CIAOSR_MAX_SCALES];        /* only {2} is implemented on device  */
  float softmax_scale;                      /* 10 in the reference                */
  const float* match1_w; const float* match1_b; const float* match1_slope;   /* C -> C/2 */
  const float* match2_w; const float* match2_b; const float* match2_slope;   /* C -> C/2 */
  const float* assembly_w; const float* assembly_b; const float* assembly_slope; /* C -> C */
  const float* down_w; const float* down_b;                                  /* 3x3 stride 2 */
  const float* escape_nan;                  /* 1-element buffer (1e-4)            */
} ciaosr_cs_attn_desc;

/* The whole head (ciaosr_net.py:31-85). */
typedef struct ciaosr_head_desc {
  int32_t abi_version;                      /* CIAOSR_ABI_VERSION                 */
  int32_t channels;                         /* encoder feature channels C         */
  int32_t feat_unfold;                      /* must be 1                          */
  int32_t local_size;                       /* 1, 2 or 3 -> 1, 4 or 9 neighbours  */
  int32_t non_local_attn;                   /* 0/1                                */
  float softmax_scale;                      /* inner attention: softmax(a / s)    */
  ciaosr_mlp_desc imnet_q, imnet_k, imnet_v;
  ciaosr_cs_attn_desc cs_attn;              /* ignored when non_local_attn == 0   */
} ciaosr_head_desc;

const char* ciaosr_last_error(void);
int ciaosr_abi_version(void);

/* Number of CUDA kernels this library has launched in the calling process
 * since load (bench.py's `gpu_launches`). */
long long ciaosr_launch_count(void);

/* 1 if `engine` (CIAOSR_ENGINE_SIMT / _TCGEN05) can run this head, else 0;
 * negative on an invalid descriptor. */
int ciaosr_engine_supported(const ciaosr_head_desc* desc, int engine);

/* ---- stage timing (bench.py's roofline) ----------------------------------
 * When enabled, the forward calls bracket each stage with CUDA events on the
 * caller's stream.  ciaosr_profile_read synchronises those events, adds the
 * elapsed milliseconds per stage into ms[0..n) / launches[0..n) and clears the
 * list.  Stages: 0 layout, 1 cross-scale attention, 2 LR precompute,
 * 3 (query,neighbour) MLP stacks + inner attention, 4 query MLP + residual,
 * 5 native RDN encoder, 6 encoder Linear layers (ciaosr_linear_forward). */
#define CIAOSR_N_STAGES 7
int ciaosr_profile_enable(int on);
int ciaosr_profile_read(float* ms, int* launches, int n);

/* ---- plan: weights re-laid-out for the kernels ------------------------- */
int ciaosr_plan_bytes(const ciaosr_head_desc* desc, size_t* bytes);
int ciaosr_plan_init(const ciaosr_head_desc* desc, void* plan, size_t plan_bytes,
                     void* stream);

/* ---- workspace --------------------------------------------------------- */
/* Upper bound for one call with these shapes (q = queries per image).      */
int ciaosr_workspace_bytes(const ciaosr_head_desc* desc, int B, int H, int W,
                           int q, int engine, size_t* bytes);

/* ---- cross-scale attention --------------------------------------------- */
/* feature [B,C,H,W] NCHW -> out [B,C*n_scales,H,W] NCHW.  engine: AUTO uses the
 * tensor-core path when C % 8 == 0, SIMT forces the fp32 CUDA-core path.     */
int ciaosr_cross_scale_attn_workspace_bytes(const ciaosr_head_desc* desc, int B, int H, int W,
                                            int engine, size_t* bytes);
int ciaosr_cross_scale_attn_forward(const ciaosr_head_desc* desc, const void* plan,
                                    const float* feature, int B, int H, int W, int engine,
                                    float* out, void* workspace, size_t workspace_bytes,
                                    void* stream);

/* ---- the head ----------------------------------------------------------- */
/* feature   [B,C,H,W]   encoder output (NCHW, as gen_feature returns it)
 * nonlocal  [B,Cn,H,W]  cross-scale attention of `feature`, or NULL: computed
 *                       here (once per call, not once per eval_bsize chunk;
 *                       the result is identical) when desc->non_local_attn
 * coord     [B,q,2]     (y,x) in [-1,1]
 * cell      [B,q,2]
 * lr_image  [B,3,H,W]   normalised LR input for the bilinear residual of
 *                       LocalImplicitSRNet.forward, or NULL for bare query_rgb
 * eval_bsize            > 0: the reference's chunk length; only effect on the
 *                       result is that tx,ty (ciaosr_net.py:162-163) are read
 *                       from the first cell of each chunk.  <= 0: one chunk.
 * out       [B,q,3]
 */
int ciaosr_query_rgb_forward(const ciaosr_head_desc* desc, const void* plan,
                             const float* feature, const float* nonlocal,
                             const float* coord, const float* cell,
                             const float* lr_image,
                             int B, int H, int W, int q, int eval_bsize,
                             int engine, float* out,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- RDN encoder fast path (SURVEY.md 8f "next" #2) ------------------------------
 * mmedit's RDN as the generator hoists it (ciaosr_net.py:314-318) and runs it in
 * gen_feature (ciaosr_net.py:321-342): sfe1, sfe2, num_blocks x [num_layers x
 * (conv3x3 + ReLU, dense concat), 1x1 local fusion, + input], 1x1 + 3x3 global fusion,
 * + sfe1 output.  Weights are [out, in, kh, kw] row-major as in the state_dict;
 * dense_w/dense_b are HOST arrays of num_blocks*num_layers device pointers
 * (block-major), lff_w/lff_b host arrays of num_blocks device pointers.
 * Only mid_channels == channel_growth == 64 is implemented. */
typedef struct ciaosr_rdn_desc {
  int32_t abi_version;
  int32_t mid_channels, channel_growth, num_blocks, num_layers;
  const float* sfe1_w; const float* sfe1_b;     /* [64,3,3,3]  */
  const float* sfe2_w; const float* sfe2_b;     /* [64,64,3,3] */
  const float* const* dense_w; const float* const* dense_b;
  const float* const* lff_w; const float* const* lff_b;
  const float* gff0_w; const float* gff0_b;     /* [64, 64*num_blocks, 1, 1] */
  const float* gff1_w; const float* gff1_b;     /* [64,64,3,3] */
} ciaosr_rdn_desc;

int ciaosr_rdn_plan_bytes(const ciaosr_rdn_desc* desc, size_t* bytes);
int ciaosr_rdn_plan_init(const ciaosr_rdn_desc* desc, void* plan, size_t plan_bytes, void* stream);
int ciaosr_rdn_workspace_bytes(const ciaosr_rdn_desc* desc, int B, int H, int W, size_t* bytes);
/* x [B,3,H,W] (normalised LR image, NCHW fp32) -> feature [B,64,H,W] NCHW fp32 */
int ciaosr_rdn_forward(const ciaosr_rdn_desc* desc, const void* plan, const float* x, int B, int H,
                       int W, float* feature, void* workspace, size_t workspace_bytes, void* stream);

/* ---- fp32-grade Linear on the tensor cores (SURVEY.md 8f "next" #2) ---------------
 * out[rows, out_features] = act(x[rows, in_features] . weight^T + bias): the nn.Linear
 * layers of the SwinIR trunk (qkv / proj / fc1 / fc2, swinir_net.py:15-31, 66-146) that
 * LocalImplicitSRSWINIR.gen_feature runs (ciaosr_net.py:475-525).  weight [out, in]
 * row-major fp32 as in the state_dict, bias [out] or NULL; in/out multiples of 4.
 * activation: 0 none, 1 exact GELU (nn.GELU()), 2 ReLU (ciaosr_linear_forward_res only).  Same fp16 hi/lo split arithmetic as the
 * head (fp32-grade; TF32 would put the features ~1e-3 off). */
typedef struct ciaosr_linear_desc {
  int32_t abi_version;
  int32_t in_features, out_features;
  const float* weight;
  const float* bias;
} ciaosr_linear_desc;

int ciaosr_linear_plan_bytes(const ciaosr_linear_desc* desc, size_t* bytes);
int ciaosr_linear_plan_init(const ciaosr_linear_desc* desc, void* plan, size_t plan_bytes, void* stream);
int ciaosr_linear_forward(const ciaosr_linear_desc* desc, const void* plan, const float* x, long long rows,
                          int activation, float* out, void* stream);

/* as above, plus `residual` [rows, out_features] (or NULL) added after the activation: x + proj(o), x + fc2(h)
 * of SwinTransformerBlock.forward (swinir_net.py:276-278) without a separate elementwise pass. */
int ciaosr_linear_forward_res(const ciaosr_linear_desc* desc, const void* plan, const float* x, long long rows,
                              int activation, const float* residual, float* out, void* stream);

/* The same Linear with its operand (and optionally its result) in the SPLIT representation the kernels compute in:
 * a row-major matrix of IEEE fp16 "hi" halves and one of "lo" halves (x = hi + lo, hi = fp16(x), lo = fp16(x - hi);
 * 22 mantissa bits), [rows, ld] with ld a multiple of 8 and columns [features, ld) zero.  Producers in this library
 * write that form directly (ciaosr_layernorm_split_forward, ciaosr_window_attention_split_forward, this function
 * with out == NULL), so a chain LayerNorm -> Linear -> ... never pays for re-splitting fp32 activations in the GEMM's
 * row threads: the operand tiles are TMA-loaded straight into the UMMA layout.
 * out != NULL: fp32 result [rows, out_features]; out == NULL: split result in out_hi / out_lo [rows, ldo]. */
int ciaosr_linear_forward_split(const ciaosr_linear_desc* desc, const void* plan, const uint16_t* a_hi,
                                const uint16_t* a_lo, int lda, long long rows, int activation,
                                const float* residual, float* out, uint16_t* out_hi, uint16_t* out_lo, int ldo,
                                void* stream);

/* ---- fp32-grade 3x3 convolution on NHWC maps (SURVEY.md 8f "next" #2) -------------------------------
 * nn.Conv2d(Cin, Cout, 3, 1, 1) as the SwinIR trunk uses it after every residual Swin block group and after the body
 * (`RSTB.conv`, `conv_after_body`; swinir_net.py:446-483, 706-713 -> ciaosr_net.py:503-525), evaluated directly on
 * the token tensor [B, H*W, C] = NHWC map (no patch_unembed / patch_embed transposes) as an implicit GEMM on the
 * tensor cores with the same fp16 hi/lo split arithmetic; `residual` [B,H,W,Cout] (or NULL) is added in the epilogue
 * (the RSTB's `+ x`).  weight [Cout, Cin, 3, 3] row-major as in the state_dict; Cin, Cout multiples of 4. */
typedef struct ciaosr_conv3x3_desc {
  int32_t abi_version;
  int32_t in_channels, out_channels;
  const float* weight;
  const float* bias;                       /* [Cout] or NULL */
} ciaosr_conv3x3_desc;

int ciaosr_conv3x3_plan_bytes(const ciaosr_conv3x3_desc* desc, size_t* bytes);
int ciaosr_conv3x3_plan_init(const ciaosr_conv3x3_desc* desc, void* plan, size_t plan_bytes, void* stream);
/* activation: 0 none, 2 ReLU (EDSR's residual blocks); out = act(conv(x) + bias) + residual */
int ciaosr_conv3x3_nhwc_forward(const ciaosr_conv3x3_desc* desc, const void* plan, const float* x, int B, int H,
                                int W, int activation, const float* residual, float* out, void* stream);

/* ---- window attention of the SwinIR trunk (SURVEY.md 8f "next" #2) ---------------------------------
 * W-MSA / SW-MSA as SwinTransformerBlock.forward runs it (swinir_net.py:240-280 around WindowAttention.forward
 * :112-146): cyclic shift by -shift, ws x ws window partition, per window and head
 * softmax(q * scale . k^T + relative_position_bias [+ the -100 region mask of SW-MSA when shift > 0]) . v,
 * window reverse, reverse shift, heads concatenated.
 * qkv        [B, H*W, 3C]  output of the block's qkv Linear on the LayerNorm'd tokens in natural (y, x) order
 *                          (q | k | v, each [heads, C/heads])
 * bias_table [(2ws-1)^2, heads]  relative_position_bias_table as in the state_dict
 * out        [B, H*W, C]   natural token order: what the block feeds to `proj`
 * H, W multiples of ws; C/heads <= 32; ws*ws*heads <= 384; 0 <= shift < ws.  fp32 CUDA cores. */
int ciaosr_window_attention_forward(const float* qkv, const float* bias_table, int B, int H, int W, int C,
                                    int heads, int ws, int shift, float scale, float* out, void* stream);

/* nn.LayerNorm(C) of the trunk (swinir_net.py:195, 207, 702) over [rows, C] fp32, C <= 512, affine. */
int ciaosr_layernorm_forward(const float* x, const float* gamma, const float* beta, float eps, long long rows,
                             int C, float* out, void* stream);

/* split-output variants (see ciaosr_linear_forward_split): results as fp16 hi / lo halves [rows, ld], ld % 8 == 0 */
int ciaosr_layernorm_split_forward(const float* x, const float* gamma, const float* beta, float eps, long long rows,
                                   int C, uint16_t* out_hi, uint16_t* out_lo, int ld, void* stream);
int ciaosr_window_attention_split_forward(const float* qkv, const float* bias_table, int B, int H, int W, int C,
                                          int heads, int ws, int shift, float scale, uint16_t* out_hi,
                                          uint16_t* out_lo, int ld, void* stream);

/* ---- tiled inference epilogue (ciaosr.py:218-258, 160-163) -------------- */
/* acc/cnt [B,3,Ho,Wo] += tile prediction [B, th*tw, 3] placed at (y0,x0).   */
int ciaosr_tile_blend_accumulate(const float* tile_pred, int B, int th, int tw,
                                 float* acc, float* cnt, int Ho, int Wo,
                                 int y0, int x0, void* stream);
/* out [B,Ho*Wo,3] = clamp01?(acc/cnt * std + mean).                         */
int ciaosr_tile_blend_finish(const float* acc, const float* cnt, int B, int Ho, int Wo,
                             const float* mean3, const float* std3, int clamp01,
                             float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CIAOSR_B200_H_ */
