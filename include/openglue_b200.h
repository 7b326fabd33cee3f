/*
 * openglue_b200.h  --  C ABI of libopenglue_b200.so
 *
 * A B200 (sm_100a) implementation of ONE path of ucuapps/OpenGlue: the SuperGlue-style
 * matching core (reference models/superglue/*, models/matching_module.py:149-187).
 * The reference is pure Python and has no FFI; these entry points are what a binding for
 * this path binds (ctypes today).  Each function names the reference
 * code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to caller-owned memory unless the name ends in
 *     `_host`; nothing is allocated, freed or synchronised behind the caller's back;
 *   - every call enqueues work on `stream` and returns immediately;
 *   - return value: OG_OK (0) or a negative og_status; og_last_error() gives a
 *     thread-local message for the last failure on the calling thread;
 *   - activations are keypoint-major: row = keypoint, `d` channels contiguous.  The
 *     reference's channel-first [B, d, n] tensors appear only at the Python-visible
 *     outputs (context descriptors);
 *   - all floating-point data is IEEE fp32; match indices are int64 (torch.int64).
 */
#ifndef OPENGLUE_B200_H_
#define OPENGLUE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OG_VERSION 100          /* major*10000 + minor*100 + patch */

typedef enum og_status {
  OG_OK = 0,
  OG_EINVAL = -1,               /* bad argument (null pointer, non-positive size, misalignment) */
  OG_EUNSUPPORTED = -2,         /* shape / option outside what the kernels cover               */
  OG_ECUDA = -3,                /* a CUDA runtime call failed (message has the CUDA error)      */
  OG_EWORKSPACE = -4            /* workspace too small                                          */
} og_status;

/* arithmetic of the GEMM-shaped contractions */
typedef enum og_precision {
  OG_PREC_FP32 = 0,             /* CUDA-core FFMA, fp32 accumulate: the exact mode                 */
  OG_PREC_TF32X3 = 1,           /* tcgen05 kind::tf32, hi/lo operand split, 3 products, fp32 accum  */
  OG_PREC_FP16X3 = 2            /* tcgen05 kind::f16: fp16 hi/lo operands with power-of-two tensor scales (same 10-bit
                                   mantissas, twice the MMA rate, half the operand bytes); GNN layers with head_dim 64,
                                   everything else as OG_PREC_TF32X3.  Entry point: og_superglue_forward_f16          */
} og_precision;

#define OG_MAX_HIDDEN 8

/* Mirrors the reference's nested config dict (models/superglue/superglue.py:12-27). */
typedef struct og_config {
  int32_t descriptor_dim;       /* d;  config['descriptor_dim']                                  */
  int32_t num_heads;            /* H;  attention_gnn.num_heads, heads = contiguous channel blocks */
  int32_t num_layers;           /* 2 * attention_gnn.num_stages (even = self, odd = cross)        */
  int32_t side_info_size;       /* S;  positional_encoding.side_info_size                          */
  int32_t num_hidden;           /* len(positional_encoding.hidden_layers_sizes)                    */
  int32_t hidden[OG_MAX_HIDDEN];
  int32_t sinkhorn_iters;       /* otp.num_iters                                                   */
  float   sinkhorn_reg;         /* otp.reg                                                         */
  float   match_threshold;      /* inference.match_threshold (config/config.yaml:40)               */
  int32_t precision;            /* og_precision                                                    */
  int32_t no_descriptors;       /* config.get('no_descriptors'): the GNN starts from the positional encoding alone
                                   (superglue.py:45-49); the residual mix still uses the raw descriptors (:59-62) */
} og_config;

int         og_version(void);
const char* og_last_error(void);
/* Number of SMs / compute capability of the current device (capability cache, read-only). */
int         og_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Packed weights.  The host side folds eval-mode BatchNorm forward into the following conv
 * (reference models/utils.py:48-58: Conv -> ReLU -> BN), folds `out_proj` into `fc.0`
 * (attention_gnn.py:32,52-55) and sigmoid(mix_coefs) into `linear_proj` (superglue.py:58-62),
 * in float64, and lays the result out as one flat fp32 buffer:
 *
 *   kenc:   for i in 0..num_hidden:  W_i [out_i, in_i],  b_i [out_i]      (in_0 = 2 + S)
 *   layer l (0..num_layers-1):  Wqkv [3d, d], bqkv [3d], W1 [2d, 2d], b1 [2d], W2 [d, 2d], b2 [d]
 *   final:  Wp [d, d], bp [d], rmix [d]   (rmix = 1 - sigmoid(mix_coefs), zeros without residual)
 *   dustbin_score [1]
 *
 * og_packed_offset() returns the float offset of a tensor inside that buffer.
 * ------------------------------------------------------------------------------------------- */
typedef enum og_tensor_id {
  OG_T_KENC_W = 0, OG_T_KENC_B = 1,               /* index = kenc linear layer 0..num_hidden     */
  OG_T_QKV_W = 2, OG_T_QKV_B = 3, OG_T_FC1_W = 4, OG_T_FC1_B = 5, OG_T_FC2_W = 6, OG_T_FC2_B = 7,
                                                  /* index = GNN layer                           */
  OG_T_PROJ_W = 8, OG_T_PROJ_B = 9, OG_T_PROJ_RMIX = 10, OG_T_DUSTBIN = 11   /* index ignored    */
} og_tensor_id;

int64_t og_packed_weight_floats(const og_config* cfg);
/* hi = round-to-nearest tf32(src), lo = round-to-nearest tf32(src - hi): the operand split of the
 * OG_PREC_TF32X3 kernels (x ~= hi + lo to 2^-23 |x|).  Device pointers, n floats each.            */
int og_split_tf32(const float* src, float* hi, float* lo, int64_t n, void* stream);
int64_t og_packed_offset(const og_config* cfg, int tensor_id, int index);

/* ---------------------------------------------------------------------------------------------
 * The whole path.  Replaces SuperGlue.forward (models/superglue/superglue.py:29-72) plus the
 * match extraction of MatchingTrainingModule.forward (models/matching_module.py:174-187) and
 * its reverse direction (inference.py:176-190).
 *
 *   kpts{0,1}  [B, n|m, 2] pixel (x, y);  side{0,1} [B, n|m, S];  desc{0,1} [B, n|m, d]
 *   img_wh     host array {W0, H0, W1, H1}  (superglue.py:35-41, 74-78)
 *   ctx{0,1}   [B, d, n|m]  channel-first context descriptors   (may be NULL)
 *   scores     [B, n+1, m+1] log assignment incl. dustbins
 *   matches0   [B, n] int64 (-1 = no match), mscores0 [B, n];  matches1/mscores1 [B, m] (may be NULL)
 *   packed_hi/lo  tf32 split of packed_weights (og_split_tf32), same offsets; required when
 *              cfg->precision == OG_PREC_TF32X3, ignored (may be NULL) for OG_PREC_FP32
 *   workspace  >= og_workspace_bytes(cfg, B, n, m), 256-byte aligned
 * ------------------------------------------------------------------------------------------- */
int64_t og_workspace_bytes(const og_config* cfg, int batch, int n, int m);

int og_superglue_forward(const og_config* cfg, const float* packed_weights,
                         const float* packed_hi, const float* packed_lo,
                         int batch, int n, int m,
                         const float* kpts0, const float* kpts1,
                         const float* side0, const float* side1,
                         const float* desc0, const float* desc1,
                         const float* img_wh_host,
                         float* ctx0, float* ctx1, float* scores,
                         int64_t* matches0, float* mscores0,
                         int64_t* matches1, float* mscores1,
                         void* workspace, int64_t workspace_bytes, void* stream);

/* OG_PREC_FP16X3 form of the whole path: additionally takes the fp16 hi/lo split of the GNN weights and its per-tensor
 * meta data (og_pack_f16; same element offsets as packed_weights).  packed_hi / packed_lo (tf32) are still required: the
 * final projection and the score GEMM run the tf32 form.                                                            */
int64_t og_f16_meta_floats(const og_config* cfg);
int og_pack_f16(const og_config* cfg, const float* packed_weights, void* hi16, void* lo16, float* meta, void* stream);
int og_superglue_forward_f16(const og_config* cfg, const float* packed_weights,
                             const float* packed_hi, const float* packed_lo,
                             const void* packed_hi16, const void* packed_lo16, const float* meta16,
                             int batch, int n, int m,
                             const float* kpts0, const float* kpts1,
                             const float* side0, const float* side1,
                             const float* desc0, const float* desc1,
                             const float* img_wh_host,
                             float* ctx0, float* ctx1, float* scores,
                             int64_t* matches0, float* mscores0,
                             int64_t* matches1, float* mscores1,
                             void* workspace, int64_t workspace_bytes, void* stream);

/* Number of kernel launches the last og_superglue_forward on this thread enqueued. */
int og_last_forward_launches(void);

/* Kernel-variant switches of the tcgen05 path (process-wide; -1 leaves a switch unchanged).
 *   gemm_pair / attention_pair = 1 (default): cta_group::2 form (one MMA spans a CTA pair, M = 256); 0: one CTA per tile.
 * Both forms are parity-tested; on B200 the paired forms are ~7 % faster end to end (profiles/README.md).  Env
 * defaults: OG_GEMM_PAIR, OG_ATTN_PAIR; a negative argument leaves that setting unchanged.                                                                           */
int og_set_tuning(int gemm_pair, int attention_pair);
/* fuse_projections = 1 (default; env OG_FUSE_QKV): the Q / K / V projections that share their input run as one launch over the
 * stacked weights (fp16x3 path, descriptor_dim % 128 == 0); 0: one launch per projection.  Bit-identical results (tested).
 * Returns the previous setting; a negative argument only queries.                                                   */
int og_set_fusion(int fuse_projections);

/* ---------------------------------------------------------------------------------------------
 * Operator-level entry points (what the whole-path call is built from; tested one by one
 * against the matching oracle function).
 * ------------------------------------------------------------------------------------------- */

/* Y = epilogue( alpha * [A | A2] . W^T + bias ).   Replaces every Conv1d(k=1) of the path
 * (attention_gnn.py:16-20,24-26,32; models/utils.py:53-57; superglue.py:58) and, batched,
 * calculate_matching_score (superglue.py:80-86).
 *   A  [batch][rows, k1] (lda, strideA)   A2 [batch][rows, k2] or NULL (concat along K)
 *   W  [nout, k1+k2] (ldw; strideW != 0 => one W per batch item)     bias [nout] or NULL
 *   relu: clamp at 0 after bias.   R/rscale: Y += rscale[o] * R[r, o]  (rscale NULL => 1)
 *   Y  [rows, nout] (ldy) and/or Yt [nout, rows] (ldyt) - either may be NULL              */
typedef struct og_linear_args {
  const float* A;  int64_t lda;  int64_t strideA;
  const float* A2; int64_t lda2; int64_t strideA2;
  int32_t k1, k2;
  const float* W;  int64_t ldw;  int64_t strideW;
  const float* bias;
  int32_t rows, nout, batch;
  float   alpha;
  int32_t relu;
  const float* R;  int64_t ldr;  int64_t strideR;
  const float* rscale;
  float* Y;  int64_t ldy;  int64_t strideY;
  float* Yt; int64_t ldyt; int64_t strideYt;
} og_linear_args;
int og_linear_fwd(const og_linear_args* args, int precision, void* stream);
/* Tensor-core (tcgen05, 3xTF32) form of og_linear_fwd: W is given pre-split (Whi/Wlo, same layout as
 * args->W, which is ignored).  mode 2: production kernel (persistent, TMA-fed, chunked accumulation
 * drained with round-to-nearest adds); mode 0 / 1: first-generation kernel with the A operand through
 * TMEM / through shared memory (kept as cross-checks of the descriptor and TMEM-operand paths).  Yhi/Ylo (Ythi/Ytlo): optional split copies of Y (Yt) for use as the next
 * kernel's B operand; same ld/stride as Y (Yt).                                                    */
int og_linear_tc_fwd(const og_linear_args* args, const float* Whi, const float* Wlo,
                     float* Yhi, float* Ylo, float* Ythi, float* Ytlo, int mode, void* stream);

/* out[b, i, h*Dh + c] = sum_j softmax_j(q_i . k_j * Dh^-0.5) v_j[c]  per head h.
 * Replaces softmax_attention (models/superglue/attention.py:8-19) inside
 * MultiheadAttention.forward (attention_gnn.py:22-32); the N x M probabilities are never
 * materialised.  q [batch][nq, *] row stride ldq; k, v [batch][nk, *]; heads are contiguous
 * channel blocks of width Dh = d / H.  Raw fp32 operands: this entry point always runs the exact
 * fp32 kernel whatever `precision` says; the tensor-core forms take operands that the projection
 * GEMMs have already split (og_attention_tc_fwd, og_attention_f16_fwd).                        */
int og_attention_fwd(const float* q, int64_t ldq, int64_t strideq,
                     const float* k, int64_t ldk, int64_t stridek,
                     const float* v, int64_t ldv, int64_t stridev,
                     float* out, int64_t ldo, int64_t strideo,
                     int batch, int nq, int nk, int num_heads, int head_dim,
                     int precision, void* stream);

/* Tensor-core (tcgen05, 3xTF32) form of og_attention_fwd.  Operands as the projection GEMM leaves them:
 *   q fp32 [batch][nq, ldq];  khi/klo tf32-split [batch*nk, ldk] keypoint-major;
 *   vthi/vtlo tf32-split [batch*d, ldvt] channel-major (d = num_heads*head_dim; ldvt >= nk, multiple of 4).
 * head_dim must be 32 or 64.                                                                        */
int og_attention_tc_fwd(const float* q, int64_t ldq, int64_t strideq,
                        const float* khi, const float* klo, int64_t ldk,
                        const float* vthi, const float* vtlo, int64_t ldvt,
                        float* out, int64_t ldo, int64_t strideo,
                        int batch, int nq, int nk, int num_heads, int head_dim, void* stream);

/* fp16 hi/lo ("3xFP16") forms of the two tensor-core operators (OG_PREC_FP16X3).  Operand scales never pass through the
 * host: every tensor has a device scalar - its tracked max |x| (fp32 tensors) or the power-of-two scale it was written
 * with (fp16 tensors).
 *   og_weight_split_f16: w [rows, cols] fp32 (+ bias [rows] or NULL) -> hi16 / lo16 (fp16, same layout) and
 *       meta[4] = {scale, max_n ||w_n||_1, max |bias|, 0}; also the way to split an activation tensor for a test.
 *   og_amax: slot = max |x|.
 *   og_linear_f16_fwd: args as og_linear_fwd (W, rscale, Yt ignored / must be NULL); exactly ONE output kind:
 *       args->Y (fp32, + optional R residual, amax_out) | Yh,Yl (fp16 [rows, nout], ldy) | Yth,Ytl (fp16 [nout, rows], ldyt);
 *       fp16 outputs are written with *scale_out (derived from a bound: a_amax * meta[1] + meta[2]).
 *   og_attention_f16_fwd: q fp32 with its amax; khi/klo fp16 [batch*nk, ldk] with k_scale; vthi/vtlo fp16 [batch*d, ldvt]
 *       with v_scale; head_dim 64.  swap_halves is a layout probe and must be 0.                                      */
int og_weight_split_f16(const float* w, const float* bias, int rows, int cols, void* hi16, void* lo16, float* meta, void* stream);
int og_amax(const float* x, int64_t n, float* slot, void* stream);
int og_linear_f16_fwd(const og_linear_args* args, const void* Wh16, const void* Wl16, const float* w_meta, const float* a_amax,
                      float* amax_out, float* scale_out, void* Yh, void* Yl, void* Yth, void* Ytl, int swap_halves, void* stream);
int og_attention_f16_fwd(const float* q, int64_t ldq, int64_t strideq, const float* q_amax,
                         const void* khi, const void* klo, int64_t ldk, const float* k_scale,
                         const void* vthi, const void* vtlo, int64_t ldvt, const float* v_scale,
                         float* out, int64_t ldo, int64_t strideo, float* out_amax,
                         int batch, int nq, int nk, int num_heads, int head_dim, int swap_halves, void* stream);

/* Dustbin-augmented log-domain Sinkhorn.  Replaces SuperGlue.get_matching_probs
 * (superglue.py:88-111) + log_otp_solver (optimal_transport.py:4-28).
 *   S       [B][n, lds]  inner score block (lds >= m, multiple of 4, 16-byte aligned rows)
 *   dustbin device scalar;   scores [B, n+1, m+1]
 *   workspace >= og_sinkhorn_workspace_bytes(B, n, m)                                          */
int64_t og_sinkhorn_workspace_bytes(int batch, int n, int m);
int og_sinkhorn_fwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin,
                    int batch, int n, int m, int iters, float reg,
                    float* scores, void* workspace, int64_t workspace_bytes, void* stream);

/* Training form of the Sinkhorn operator.  og_sinkhorn_train_fwd = og_sinkhorn_fwd that also
 * records the scaling vectors of every iteration (hist: og_sinkhorn_hist_floats floats: u [B][T][n+1], v [B][T+1][m+1]);
 * og_sinkhorn_bwd = the gradient through all T unrolled iterations, as torch autograd computes it for the reference's
 * get_matching_probs / log_otp_solver in training_step (matching_module.py:99-105):
 *   dscores [B,n+1,m+1] = d loss / d scores (dense)  ->  dS_aug [B,n+1,m+1] = d loss / d S_aug (its [:n,:m] block is
 *   d loss / d S of the score GEMM, reg included) and ddustbin[0] = d loss / d dustbin_score.  Deterministic.        */
int64_t og_sinkhorn_hist_floats(int batch, int n, int m, int iters);
int og_sinkhorn_train_fwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin,
                          int batch, int n, int m, int iters, float reg, float* scores, float* hist,
                          void* workspace, int64_t workspace_bytes, void* stream);
int64_t og_sinkhorn_bwd_workspace_bytes(int batch, int n, int m, int iters);
int og_sinkhorn_bwd(const float* S, int64_t lds, int64_t strideS, const float* dustbin,
                    int batch, int n, int m, int iters, float reg, const float* hist,
                    const float* dscores, float* dS_aug, float* ddustbin,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Mutual-argmax match extraction on scores[:, :n, :m].  Replaces
 * models/matching_module.py:174-187 and inference.py:176-190 (ties -> lowest index).
 *   workspace >= og_match_workspace_bytes(B, n, m)                                             */
int64_t og_match_workspace_bytes(int batch, int n, int m);
int og_match_fwd(const float* scores, int batch, int n, int m, float threshold,
                 int64_t* matches0, float* mscores0, int64_t* matches1, float* mscores1,
                 void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Ground-truth match generation: the step immediately BEFORE the matching core in the reference's
 * training / validation step (models/matching_module.py:84-93).  Replaces
 * generate_gt_matches (models/gt_matches_generation.py:17-93) with reproject_keypoints /
 * get_inverse_transformation (utils/misc.py:21-103): reprojection of both keypoint sets, the two
 * N x M torch.cdist + min, the mutual check and the UNMATCHED (-1) / IGNORE (-2) marks.  The
 * reference's threshold refinements (:56-67, :76-78) assign through boolean-mask copies and have
 * no effect; they are not reproduced, so the thresholds are not parameters here.
 * ------------------------------------------------------------------------------------------- */
enum { OG_GT_PERSPECTIVE = 0, OG_GT_3D_REPROJECTION = 1 };   /* transformation['type'] (utils/misc.py:23-34) */
typedef struct og_gt_transform {
  int32_t type;
  const float* H;                    /* perspective: [B,3,3]                                            */
  const float* K0; const float* K1;  /* 3d: intrinsics [B,3,3]                                          */
  const float* R;  const float* T;   /* 3d: relative pose [B,3,3], [B,3]  (x1 = R x0 + T)               */
  const float* depth0;               /* 3d: per-keypoint depth [B,n] / [B,m], or depth images           */
  const float* depth1;               /*     [B,depth{0,1}_h,depth{0,1}_w] when depth_is_image           */
  int32_t depth_is_image;
  int32_t depth0_h, depth0_w, depth1_h, depth1_w;
} og_gt_transform;

int64_t og_gt_matches_workspace_bytes(int batch, int n, int m);
/* kpts0 [B,n,2], kpts1 [B,m,2] pixel coordinates; gt_matches0 [B,n], gt_matches1 [B,m] int64
 * (index of the match, -1 unmatched, -2 ignore).  All pointers are device memory.                  */
int og_gt_matches_fwd(const float* kpts0, const float* kpts1, int batch, int n, int m, const og_gt_transform* tf,
                      int64_t* gt_matches0, int64_t* gt_matches1, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batch collation of cached local features: the step that FEEDS the path from the cached-feature dataset.  Replaces
 * MegaDepthPairsDataModuleFeatures.stack_keypoints_batch (data/megadepth_datamodule.py:105-168): per image, the
 * `target_keypoints` most confident keypoints in descending score order (select == NULL; torch.topk, ties -> lower
 * index) or the caller's selection (select [2*batch, target_keypoints] int32: torch.randperm indices, training), all of
 * them + zero padding when an image has fewer; depth{0,1} [batch, H, W] images are sampled at (int(y), int(x)) of every
 * kept keypoint (NULL: no depth).  Raw inputs are the images' features concatenated in the order (pair 0, image 0),
 * (pair 0, image 1), (pair 1, image 0) ...: lafs [total,2,3], scores [total], desc [total,D], offsets [2*batch+1] int32;
 * max_count = the largest image (<= 16384).  Outputs [batch, target, ...] per image side.  Index work: bit-exact.      */
int og_collate_fwd(const float* lafs, const float* scores, const float* desc, const int* offsets, const int* select, int max_count,
                   const float* depth0, int depth0_h, int depth0_w, const float* depth1, int depth1_h, int depth1_w,
                   int batch, int target_keypoints, int descriptor_dim,
                   float* out_lafs0, float* out_lafs1, float* out_scores0, float* out_scores1, float* out_desc0, float* out_desc1,
                   float* out_depth0, float* out_depth1, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Matching loss: the step immediately AFTER the matching core in the reference's training step
 * (models/matching_module.py:101).  Replaces criterion (utils/losses.py:7-53) for margin = None (every
 * shipped config): loss[0] = 'loss' (negative log-likelihood of the ground-truth assignment, mean per set and
 * per pair), loss[1] = 'metric_loss' = 0.  gt_matches0 [B,n] / gt_matches1 [B,m] int64 as og_gt_matches_fwd
 * writes them (-1 unmatched, -2 ignore).  dscores (optional, [B,n+1,m+1], ZERO-FILLED by the caller) receives
 * grad_scale * d loss / d scores (a sparse scatter: the backward pass of the gather).  Deterministic.      */
int64_t og_criterion_workspace_bytes(int batch);
int og_criterion_fwd(const float* scores, const int64_t* gt_matches0, const int64_t* gt_matches1, int batch, int n, int m,
                     float* loss, float* dscores, float grad_scale, void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step operators: what nn.BatchNorm1d in training mode and torch autograd do for the
 * reference around the contractions of SuperGlue.forward in MatchingTrainingModule.training_step
 * (models/matching_module.py:71-105).  The contractions themselves (dX = dY W, dW = dY^T X, the attention gradients
 * dV = P^T dO, dP = dO V^T, dQ = dS K, dK = dS^T Q) run on og_linear_fwd / og_linear_tc_fwd with transposed operands.
 * Activations are row-major [rows, channels], rows = batch x keypoints.  All reductions are deterministic.
 *   og_transpose        transpose != 0: out[b][c, r] = in[b][r, c];  == 0: pitch-changing copy.  Padding is left untouched.
 *   og_colsum           out[c] = sum_r x[r,c] * (y ? y[r,c] - (z ? z[r,c] : 0) : 1)           (bias / mix gradients)
 *   og_bn_train_fwd     y = gamma (r - mean) invstd + beta, r = relu ? max(a, 0) : a, batch statistics over the rows
 *                       (biased variance), save_mean / save_invstd [cols] kept for the backward pass, running_mean /
 *                       running_var (optional) updated with `momentum` (unbiased variance): models/utils.py:48-58,
 *                       torch.nn.BatchNorm1d(training=True)
 *   og_bn_train_bwd     da (through the fused ReLU), dgamma, dbeta from dy
 *   og_softmax_rows     in-place row softmax of the materialised attention scores (models/superglue/attention.py:12-13)
 *   og_softmax_bwd_rows dP <- scale * P * (dP - sum_j P dP)
 *   og_axpby            out = a x + b y (y NULL: a x)
 *   og_mix_fwd / _bwd / _param_grad   residual mix alpha = sigmoid(mix_coefs) (superglue.py:59-62) and its gradients
 *   og_kenc_input       [2 x / (W - 1) - 1, 2 y / (H - 1) - 1, side info]  (superglue.py:74-78, positional_encoding.py:16-18)
 * workspace: >= og_train_workspace_floats(cols) floats.                                                           */
int64_t og_train_workspace_floats(int cols);
/* One GEMM of the training step, args as og_linear_fwd: the tcgen05 3xTF32 kernel when the shape is tileable (K >= 32,
 * K % 4 == 0, 16-byte aligned rows, dense batches; W is split into split_scratch, og_linear_auto_scratch_floats(args)
 * floats, on the fly), the exact fp32 CUDA-core kernel otherwise or when precision == OG_PREC_FP32.               */
int64_t og_linear_auto_scratch_floats(const og_linear_args* args);
int og_linear_auto_fwd(const og_linear_args* args, int precision, float* split_scratch, void* stream);
int og_transpose(const float* in, int64_t ld_in, int64_t stride_in, float* out, int64_t ld_out, int64_t stride_out,
                 int batch, int rows, int cols, int transpose, void* stream);
int og_colsum(const float* x, int64_t ldx, const float* y, int64_t ldy, const float* z, int64_t ldz, int rows, int cols,
              float* out, float* workspace, void* stream);
int og_bn_train_fwd(const float* a, int64_t lda, int rows, int cols, int relu, const float* gamma, const float* beta,
                    float eps, float momentum, float* y, int64_t ldy, float* save_mean, float* save_invstd,
                    float* running_mean, float* running_var, float* workspace, void* stream);
int og_bn_train_bwd(const float* dy, int64_t lddy, const float* a, int64_t lda, int rows, int cols, int relu,
                    const float* gamma, const float* save_mean, const float* save_invstd,
                    float* da, int64_t ldda, float* dgamma, float* dbeta, float* workspace, void* stream);
int og_softmax_rows(float* S, int64_t ld, int64_t rows, int cols, void* stream);
int og_softmax_bwd_rows(const float* P, float* dP, int64_t ld, int64_t rows, int cols, float scale, void* stream);
int og_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream);
/* out[r, c] (+)= sum_s part[s][r, c] (s ascending; out row stride ld_out): reduction of the split-K weight gradients      */
int og_sum_batches(const float* part, int S, int rows, int cols, float* out, int64_t ld_out, int accumulate, void* stream);
int og_mix_fwd(const float* g, const float* l, const float* mix, float* out, int64_t rows, int d, void* stream);
int og_mix_bwd(const float* dm, const float* mix, float* dg, float* dl, int64_t rows, int d, void* stream);
int og_mix_param_grad(const float* colsum, const float* mix, float* dmix, int d, void* stream);
int og_kenc_input(const float* kpts, const float* side, int rows, int side_info_size, float width, float height,
                  float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SuperPoint front-end operators: SuperPointNet.forward (models/features/superpoint/model.py:61-129)
 * on NHWC activations.  Convolutions = og_sp_im2col3x3 (3x3, pad 1) + og_linear_auto_fwd (bias / ReLU fused; 1x1: the GEMM alone).
 *   og_sp_im2col3x3     out[p, (3 ky + kx) C + c] = x[b, y + ky - 1, x + kx - 1, c] (zero padding);  x [B,H,W,C], out [B H W, 9 C]
 *   og_sp_maxpool2x2    nn.MaxPool2d(2, 2) on NHWC
 *   og_row_normalize    mode 0: x /= ||x||_2 per row (model.py:70-71);  mode 1: F.normalize (x /= max(||x||_2, eps))
 *   og_sp_heat_nms      probs [B,Hc,Wc,65] (channel softmax done) -> heat [B, 8 Hc, 8 Wc]: pixel shuffle (model.py:84-86), nms2d
 *                       (kornia >= 0.6.1, restated: x > max(0, the other k*k - 1 values of the replicate-padded window)),
 *                       F.threshold + nonzero (model.py:89-92) and remove_borders (utils.py:4-11): the score where kept, else 0
 *   og_sp_compact       per image: surviving pixels in row-major order (torch.nonzero) -> cand_idx / cand_score [B, cap], count [B]
 *   og_sp_select        per image: n_out[b] keypoints, mode[b] = 0 in candidate order | 1 = the largest scores, descending (torch.topk,
 *                       equal scores: lower index first; top_k_keypoints utils.py:34-39, min_stack models/features/utils.py:28-56);
 *                       kpts [B,out_cap,2] as (x, y) floats, scores [B,out_cap];  max_count = the largest count (<= 16384)
 *   og_sp_sample_desc   sample_desc_from_points (utils.py:14-31): bilinear grid_sample (align_corners False) of the coarse descriptors
 *                       [B,Hc,Wc,D] at the keypoints + F.normalize -> desc [B,out_cap,D]                                        */
int og_sp_im2col3x3(const float* x, int B, int H, int W, int C, float* out, void* stream);
int og_sp_maxpool2x2(const float* x, int B, int H, int W, int C, float* out, void* stream);
int og_row_normalize(float* x, int64_t rows, int C, int mode, float eps, void* stream);
int og_sp_heat_nms(const float* probs, int B, int Hc, int Wc, int nms_kernel, float threshold, int border, float* heat, void* stream);
int og_sp_compact(const float* heat, int B, int HW, int cap, int* cand_idx, float* cand_score, int* count, void* stream);
int og_sp_select(const int* cand_idx, const float* cand_score, const int* count, const int* n_out, const int* mode, int B, int cap, int W,
                 int out_cap, int max_count, float* kpts, float* scores, void* stream);
int og_sp_sample_desc(const float* coarse, int B, int Hc, int Wc, int D, const float* kpts, const int* n_out, int out_cap, int max_n, int cell,
                      float* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* OPENGLUE_B200_H_ */
