/*
 * parq_b200 -- C ABI of the B200-native PARQ decoder hot path (libparq_b200.so).
 *
 * Drop-in boundary: everything the reference computes inside
 *   PARQDecoder.forward                      (model/parq_decoder.py:134-163)
 *   -> Transformer.forward                   (model/transformer_parq.py:95-126)
 *   -> TransformerDecoder.forward            (model/transformer_parq.py:283-337)
 * is behind parq_decoder_forward().  The reference is pure Python/PyTorch and has no FFI of its
 * own; the binding a maintainer would add is the ctypes stub shown in INTEGRATION.md (it is what
 * parq_b200/_lib.py implements).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  All data pointers are DEVICE pointers owned by the
 *     caller (PyTorch's allocator); the library never allocates or frees device memory and keeps no
 *     pointer across calls.  Alignment: 16 bytes for every pointer; C % 256 == 0; Nq % 128 == 0.
 *   - every call is asynchronous on the caller's stream (cudaStream_t passed as void*), performs no
 *     host synchronisation (except parq_pack_weights, a one-off), and is graph-capturable.
 *   - return 0 on success, negative on error (PARQ_ERR_*); parq_last_error() gives the message
 *     (thread local).  No C++ exception crosses the ABI.  sm_100 devices only: there is no fallback.
 */
#ifndef PARQ_B200_H_
#define PARQ_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PARQ_ABI_VERSION 2

#define PARQ_OK 0
#define PARQ_ERR_SHAPE (-1)       /* unsupported shape / alignment / null pointer */
#define PARQ_ERR_ARCH (-2)        /* device is not sm_100 */
#define PARQ_ERR_CUDA (-3)        /* CUDA runtime / driver error */
#define PARQ_ERR_WORKSPACE (-4)   /* workspace or packed-weight buffer too small */

/* Problem shape.  Nk = T*H*W image tokens per clip; head_dim = C / heads must be 256. */
typedef struct ParqShape {
  int32_t B, T, H, W;     /* clips, views per clip, feature-map height / width */
  int32_t C;              /* token / model width (1024) */
  int32_t Nq;             /* queries per clip (256) */
  int32_t heads;          /* 4 */
  int32_t ffn;            /* 768 */
  int32_t iters;          /* recurrent iterations == DEC_LAYERS (8), weights shared */
  int32_t num_cls;        /* NUM_SEMCLS + 1 (10) */
  float scale[6];         /* [x0,x1,y0,y1,z0,z1] of TRANSFORMER.SCALE (config/eval.yaml:55) */
} ParqShape;

/* fp32 parameters exactly as stored in the reference's state dict, device pointers.
 * Conv1d(k=1) weights (n, C, 1) are passed as (n, C). */
typedef struct ParqWeightsF32 {
  const float *pe0_w, *pe0_b, *pe2_w, *pe2_b;               /* decoder.position_encoder.{0,2}      */
  const float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b;     /* layers.0.self_attn                  */
  const float *ca_in_w, *ca_in_b, *ca_out_w, *ca_out_b;     /* layers.0.multihead_attn             */
  const float *lin1_w, *lin1_b, *lin2_w, *lin2_b;           /* layers.0.linear1 / linear2          */
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
  const float *cls_w, *cls_b;                               /* mlp_heads.sem_cls_head.layers.0     */
  const float *ctr0_w, *ctr1_g, *ctr1_b, *ctr4_w, *ctr5_g, *ctr5_b, *ctr8_w, *ctr8_b;   /* center_head   */
  const float *size_w, *size_b;                             /* mlp_heads.size_head.layers.0        */
  const float *rot0_w, *rot1_g, *rot1_b, *rot4_w, *rot5_g, *rot5_b, *rot8_w, *rot8_b;   /* rotation_head */
  const float *mean_size;                                   /* (num_cls, 3) BoxProcessor table     */
  const float *dim_t;                                       /* (128) pos2posemb3d denominators     */
} ParqWeightsF32;

/* Per-iteration outputs, each (iters, B, Nq, n) fp32 contiguous; the six required tensors are the
 * keys of the reference's output dicts (transformer_parq.py:271-279).  Optional ones may be NULL. */
typedef struct ParqOutputs {
  float *pred_logits;          /* n = num_cls */
  float *center_unnormalized;  /* n = 3 */
  float *size_unnormalized;    /* n = 3 */
  float *ortho6d;              /* n = 6 */
  float *sem_cls_prob;         /* n = num_cls */
  float *coord_pos;            /* n = 3 */
  float *rotation;             /* optional (iters,B,Nq,9): compute_rotation_matrix_from_ortho6d (utils/ortho6d_transforms.py:53-66) */
  float *center_im;            /* optional (iters,B,T,Nq,2): projected pixel coordinates (transformer_parq.project) */
  uint8_t *center_valid;       /* optional (iters,B,T,Nq) */
  float *features;             /* optional (iters,B,Nq,C): pixel-aligned sampled features */
  float *decoder_out;          /* optional (iters,B,Nq,C): decoder-layer output fed to the heads */
} ParqOutputs;

#define PARQ_FLAG_SKIP_KV 1u     /* workspace already holds K / V^T of these tokens (parq_kv_project) */
#define PARQ_FLAG_WEIGHT_LO 2u   /* parq_pack_weights returned 1: weights are not bf16-exact, use the 3-term GEMMs */
#define PARQ_FLAG_KV_HI_ONLY 16u /* with WEIGHT_LO: project K / V^T with the high weight part only (they are stored in bf16 anyway) */
#define PARQ_FLAG_NO_CHAIN 64u   /* per-iteration linears as separate GEMM + LayerNorm launches instead of the chained cluster kernel (A/B timing, tests) */
#define PARQ_FLAG_FORCE_CHAIN 128u /* use the chained kernel also below its break-even batch (B*Nq < 2048 rows): tests */
/* PARQ_FLAG_HI_ONLY_SET + bits 16..26: explicit per-GEMM mask "use only the high-order activation term" for the chained path with
 * bf16-exact weights (precision / energy trade-off, measured in profiles/): bit 16 sa_qk, 17 sa_v, 18 ca_q, 19 pe0, 20 pe2,
 * 21 sa_out, 22 ca_out, 23 lin1, 24 lin2, 25 hd1, 26 hd2.  Without PARQ_FLAG_HI_ONLY_SET the library default applies. */
#define PARQ_FLAG_HI_ONLY_SHIFT 16
#define PARQ_FLAG_HI_ONLY_MASK 0x07FF0000u
#define PARQ_FLAG_HI_ONLY_SET 0x08000000u
#define PARQ_FLAG_FUSED_MERGE 512u /* merge the stream-K pieces of the cross-attention inside the attention kernel (flags + spin
                                     * wait on an earlier-scheduled CTA pair) instead of the attn3_combine_kernel launch; off by default */
#define PARQ_FLAG_NO_FORK 1024u  /* un-chained launch path: keep every launch on the caller's stream (no side-stream branches; A/B timing) */
#define PARQ_FLAG_NO_PDL 4u      /* launch without programmatic dependent launch (plain stream order; for A/B timing) */

int parq_version(void);
const char *parq_last_error(void);

/* Sizes of the caller-allocated buffers. */
size_t parq_packed_bytes(const ParqShape *shape);
size_t parq_workspace_bytes(const ParqShape *shape);

/* One-off: split/convert the fp32 parameters into the packed bf16 [hi|lo] + fp32 layout the kernels
 * read.  Synchronises the stream once.  Returns 0 when every GEMM weight is exactly representable in
 * bf16, 1 when a low-order weight term exists (pass PARQ_FLAG_WEIGHT_LO to the calls below), < 0 on error. */
int parq_pack_weights(const ParqShape *shape, const ParqWeightsF32 *w, void *packed, size_t packed_bytes, void *stream);

/* T_camera_local = T_camera_pseudoCam o (T_world_pseudoCam^-1 o T_world_local)   (transformer_parq.py:298-300)
 * T_cp, T_wp (B,T,12), T_wl (B,1,12) -> T_cl (B,T,12). */
int parq_pose_chain(const float *T_cp, const float *T_wp, const float *T_wl, float *T_cl, int B, int T, void *stream);

/* fp32 tokens (parq_lightning.py:78-85 hands the decoder fp32) as an exact bf16 pair: hi = bf16(x), lo = bf16(x - hi),
 * n elements (multiple of 4), 16-byte aligned buffers.  `hi` is the plane the K / V^T projection consumes; passing `lo`
 * as tokens_lo_bf16 below makes the gather see x to 16 mantissa bits instead of 8. */
int parq_split_tokens(const float *tokens_f32, void *hi_bf16, void *lo_bf16, long long n, void *stream);

/* transformer_parq.project (:129-161) for normalised reference points `ref` (B,Nq,3):
 * tokens bf16 (B,T*H*W,C) [+ optional low-order plane tokens_lo_bf16, may be NULL] -> features (B,Nq,C),
 * center_im (B,T,Nq,2), center_valid (B,T,Nq), coord_pos (B,Nq,3).  Output pointers other than `features` may be NULL. */
int parq_project_sample(const ParqShape *shape, const void *tokens_bf16, const void *tokens_lo_bf16, const float *ref,
                        const float *T_cl, const float *camera, float *features, float *center_im, uint8_t *center_valid,
                        float *coord_pos, void *stream);

/* Hoisted cross-attention K / V^T projection of all image tokens into the workspace (once per clip batch). */
int parq_kv_project(const ParqShape *shape, const void *tokens_bf16, const void *packed, void *workspace,
                    size_t workspace_bytes, uint32_t flags, void *stream);

/* "Next" row f-4, streaming window: re-project K / V^T of `n_views` consecutive view slots [slot0, slot0 + n_views) of every
 * clip from view_tokens_bf16 (B, n_views*H*W, C) into the workspace caches; the other views' K / V^T stay where they are.
 * A sliding 8-view window then costs one view's projection per step instead of eight (datasets/transforms.py:191-208 slides
 * the window by re-running everything).  Needs H*W % 32 == 0.  Follow with parq_decoder_forward(..., PARQ_FLAG_SKIP_KV). */
int parq_kv_project_views(const ParqShape *shape, const void *view_tokens_bf16, int slot0, int n_views, const void *packed,
                          void *workspace, size_t workspace_bytes, uint32_t flags, void *stream);

/* Instrumentation of the chained GEMM kernel: clock64 stamps of its CTA 0 for the next <= 64 launches go to `buf`
 * (device memory, 64 x 64 int64); NULL switches it off.  Read by tools/chain_timeline.py. */
int parq_chain_debug(void *buf);

/* Instrumentation of the dependent launch chain: while a buffer is set, thread 0 of block 0 of every kernel of this library
 * appends the global timer (ns) at the moment its stream dependency resolved (griddepcontrol.wait returned).  buf = device memory
 * of `capacity` uint64, slot 0 counts the stamps; NULL switches it off.  Read by tools/launch_trace.py. */
int parq_trace(void *buf, int capacity);

/* The whole recurrent decoder.  tokens_lo_bf16: optional low-order token plane (parq_split_tokens) or NULL.
 * ref0 (B,Nq,3): normalised initial reference points (sigmoid(refpoint.weight) repeated per clip).
 * forced_refs (iters,B,Nq,3) or NULL: teacher-forced reference points per iteration. */
int parq_decoder_forward(const ParqShape *shape, const void *tokens_bf16, const void *tokens_lo_bf16, const float *camera,
                         const float *T_cp, const float *T_wp, const float *T_wl, const float *ref0,
                         const float *forced_refs, const void *packed, void *workspace, size_t workspace_bytes,
                         const ParqOutputs *out, uint32_t flags, void *stream);

/* "Next" row f-3: the FPN upsample + concat of ResnetFPN.forward (model/resnet_fpn.py:73-80): the four pyramid levels
 * l0..l3, each (BT, channels_per_level, h_l, w_l) fp32 channels-first with level_hw = HOST array {h0,w0,h1,w1,h2,w2,h3,w3},
 * are bilinearly resized (F.interpolate(mode="bilinear"), align_corners=False) to the size of `target_level` and
 * concatenated along channels into out_nchw (BT, 4*channels_per_level, H, W) -- the `all_features` tensor. */
int parq_fpn_concat(const float *l0, const float *l1, const float *l2, const float *l3, const int32_t *level_hw, int BT,
                    int channels_per_level, int target_level, float *out_nchw, void *stream);
/* Same with bf16 pyramid levels (what an evaluation host ships over PCIe: 3.3 MB per view instead of 9.8 MB of tokens). */
int parq_fpn_concat_bf16(const void *l0, const void *l1, const void *l2, const void *l3, const int32_t *level_hw, int BT,
                         int channels_per_level, int target_level, float *out_nchw, void *stream);
/* General form: fp32 or bf16 levels, fp32 or bf16 output.  A bf16 `all_features` (half the bytes of this memory-bound pass
 * and of the read-back in the AddRayPE producer, which takes it with PARQ_RAYPE_FEAT_BF16) is rounded once more when the
 * encoding is added: the tokens then carry up to 1.5 bf16 ulp instead of 0.5. */
int parq_fpn_concat_ex(const void *l0, const void *l1, const void *l2, const void *l3, int levels_bf16, const int32_t *level_hw,
                       int BT, int channels_per_level, int target_level, void *out_nchw, int out_bf16, void *stream);

/* "Next" row f-1: AddRayPE.forward (model/ray_positional_encoding.py:61-139; utils/encoding_utils.py:15-100) fused with
 * the tokeniser of PARQ.forward (model/parq_lightning.py:75-85).  feat_nchw (B,T,C,H,W) fp32 backbone features (may be
 * NULL when only the encoding is wanted), depth_planes (num_samples) device floats (exp(log(min)+log(max/min)*linspace)),
 * ray_points_scale HOST array of 6 floats.  Outputs (either may be NULL): tokens_bf16 (B, T*H*W, C) = features + encoding,
 * channels-last, the decoder's input; encoding_nchw (B,T,C,H,W) fp32, AddRayPE's own return value.
 * parq_raype_pack_weights takes encoder.0.weight (C, 3n), encoder.0.bias, encoder.2.weight (C, C), encoder.2.bias and
 * returns 1 when a weight is not bf16-exact (pass PARQ_FLAG_WEIGHT_LO).  PARQ_RAYPE_SPLIT_HIDDEN keeps the hidden layer as an
 * exact [hi|lo] bf16 split (fp32-grade encoding, 2x the second GEMM); without it the hidden layer is plain bf16, which is
 * below the bf16 rounding of the tokens themselves. */
#define PARQ_RAYPE_SPLIT_HIDDEN 8u
#define PARQ_RAYPE_FEAT_BF16 256u /* feat_nchw points at bf16 features (parq_fpn_concat_ex with out_bf16) */
size_t parq_raype_packed_bytes(int C, int num_samples);
size_t parq_raype_workspace_bytes(int B, int T, int H, int W, int C, int num_samples);
int parq_raype_pack_weights(int C, int num_samples, const float *w0, const float *b0, const float *w2, const float *b2,
                            void *packed, size_t packed_bytes, void *stream);
int parq_raype_forward(int B, int T, int H, int W, int C, int num_samples, const float *feat_nchw, const float *camera,
                       const float *T_cp, const float *T_wp, const float *T_wl, const float *depth_planes,
                       const float *ray_points_scale, const void *packed, void *workspace, size_t workspace_bytes,
                       void *tokens_bf16, float *encoding_nchw, uint32_t flags, void *stream);

/* "Next" row f-2: PARQDecoder.parse_pred (model/parq_decoder.py:372-424) + nms / nms_3d_faster[_samecls]
 * (utils/nms.py:20-70, 141-224) on the device, one CTA per clip, K <= 1024 boxes per clip.  Inputs are the LAST
 * iteration's center_unnormalized, size_unnormalized (B,K,3), ortho6d (B,K,6) and sem_cls_prob (B,K,num_cls); the
 * background label is num_cls-1.  track_scale is a HOST array of 6 floats (cfg TRACK_SCALE); mode 0 is the eval
 * configuration (class-agnostic NMS, IoU threshold 0.1, track-scale filter), PARQ_NMS_SAME_CLASS |
 * PARQ_NMS_NO_TRACK_SCALE with threshold 0.2 is the FOR_VIS branch.  Outputs: pred_mask (B,K) = nms & in-scope;
 * optional nms_mask (B,K), scores (B,K), labels (B,K) int32, obbs (B,K,19) in the reference's Obb3D layout
 * [xmin,xmax,ymin,ymax,zmin,zmax | T_local_object: R row-major, t | sem_id]  (utils/wrappers.py:297-321). */
#define PARQ_NMS_SAME_CLASS 1u
#define PARQ_NMS_NO_TRACK_SCALE 2u
int parq_parse_pred(const float *center, const float *size, const float *ortho6d, const float *prob, int B, int K,
                    int num_cls, const float *track_scale, double overlap_threshold, uint32_t mode, uint8_t *pred_mask,
                    uint8_t *nms_mask, float *scores, int32_t *labels, float *obbs, void *stream);

/* ---- building blocks exported for unit tests and micro-benchmarks -------------------------------------- */

/* D[M,N] = sum_t A[:, a_koff[t] : +K] * Bw[:, b_koff[t] : +K]^T  (bf16 K-major operands, fp32 accumulate),
 * epilogue: + bias (per column, or per row when bias_per_row), optional ReLU, outputs fp32 and/or 16-bit. */
int parq_gemm_bf16(const void *A, int64_t a_rows, int64_t a_cols, const void *Bw, int64_t b_rows, int64_t b_cols,
                   int M, int N, int K, int nterms, const int32_t *a_koff, const int32_t *b_koff, const float *bias,
                   int bias_per_row, int relu, float *out_f32, int64_t ld_f32, void *out_lp, int64_t ld_lp, int lp_fp16,
                   int64_t lp_lo_off, void *stream);

/* The chained cluster kernel (csrc/chain_tc.cuh) on a two-stage chain, for unit tests:
 *   y = LayerNorm_1024(A W1^T + b1 + resid) * gamma + beta     (eps 1e-5; resid_cm is COLUMN-major fp32 [1024][M])
 *   z = relu(y W2^T + b2)
 * a_split (M, 2*K1), w1_split (1024, 2*K1), w2_split (N2, 2048): bf16 [hi|lo] along K; w_lo != 0 adds the A_hi x W_lo term.
 * Outputs: y_f32 (M, 1024) row-major, y_split (M, 2048) and z_split (M, 2*N2) bf16 [hi|lo].  M % 128 == 0, N2 in {768, 1024, 2048}. */
int parq_chain_ln_linear(const void *a_split, const void *w1_split, const float *b1, const float *resid_cm, const float *gamma,
                         const float *beta, const void *w2_split, const float *b2, int M, int K1, int N2, int w_lo, float *y_f32,
                         void *y_split, void *z_split, void *stream);

/* softmax(Q K^T) V for head_dim 256: Q (B*Nq, H*256) pre-scaled, K (B*Nk, H*256), Vt (H*256, ldv) 16-bit
 * (bf16, or fp16 when fp16 != 0); out_split (B*Nq, 2*H*256) bf16 [hi|lo]; scratch >= parq_attention_scratch_bytes.
 * force_nsplit: > 0 fixes the number of key splits, 0 lets the library choose (stream-K schedule for long key
 * sequences, else a split-KV grid), < 0 forces the stream-K schedule (needs Nq % 256 == 0). */
size_t parq_attention_scratch_bytes(int B, int H, int Nq, int Nk);
int parq_attention(const void *Q, int64_t ldq, const void *K, int64_t ldk, const void *Vt, int64_t ldv, int B, int H,
                   int Nq, int Nk, int fp16, void *scratch, size_t scratch_bytes, void *out_split, int force_nsplit,
                   void *stream);

/* Byte offset of a named intermediate inside the workspace (tests / debugging): "pe","x0","x1","x2","x3","y","h1","h2",
 * "q_c","qk_s","vt_s","a_attn","Kc","Vt","T_cl"; "ldv","ldvs","cross_nsplit","self_nsplit" return those values. -1 if unknown. */
long long parq_workspace_offset(const ParqShape *shape, const char *name);

/* ---- instrumentation ------------------------------------------------------------------------------------ */

/* Number of kernels this library has launched from the calling thread since load. */
unsigned long long parq_kernel_launches(void);

/* Event profiling of the launches inside the calls above.  tag bits: 0 K/V projection GEMMs, 1 project_sample,
 * 2 per-iteration GEMMs, 3 self-attention, 4 cross-attention, 5 split combine, 6 row-wise kernels.
 * parq_profile_enable(mask, max_records) arms it (mask 0 disarms); parq_profile_collect waits for the recorded
 * events and returns, per tag (arrays of 8), the summed device milliseconds and the number of launches; its
 * return value is 1 if records were dropped because max_records was reached. */
int parq_profile_enable(uint32_t tag_mask, int max_records);
int parq_profile_collect(float *ms_per_tag, int *launches_per_tag);

#ifdef __cplusplus
}
#endif
#endif /* PARQ_B200_H_ */
