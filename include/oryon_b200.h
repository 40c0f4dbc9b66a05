/*
 * oryon_b200.h -- C ABI of liboryon_b200.so: the B200 (sm_100a) implementation of the Oryon
 * inference hot path.
 *
 * The reference (jcorsetti/oryon) has no FFI: its boundary for this path is a set of Python
 * functions.  Each entry point below names the reference function (file:line under the reference
 * tree) whose arithmetic it replaces; the oryon_b200 Python modules are the mirrors that keep the
 * reference signatures and call these through ctypes (INTEGRATION.md shows the binding a reference
 * maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - pointers marked DEVICE are CUDA device pointers on the handle's device, HOST are host pointers.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation
 *     unless the function is documented as returning a host-visible count.
 *   - the caller owns every input and output buffer; the library owns only its handle-scoped
 *     workspace and packed weights.  Buffers are borrowed for the duration of the call's stream work.
 *   - every function returns 0 on success or a negative oryon_status; oryon_last_error() returns a
 *     thread-local human readable message for the last failure on this thread.
 *   - a handle is bound to one (process, device) and is not thread-safe; use one per thread.
 */
#ifndef ORYON_B200_H_
#define ORYON_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORYON_ABI_VERSION 3

typedef struct oryon_handle oryon_handle;

typedef enum {
  ORYON_OK = 0,
  ORYON_ERR_INVALID_ARGUMENT = -1,
  ORYON_ERR_CUDA = -2,
  ORYON_ERR_UNSUPPORTED_DEVICE = -3, /* not an sm_100 device: there is no fallback path */
  ORYON_ERR_OUT_OF_MEMORY = -4,
  ORYON_ERR_NOT_LOADED = -5          /* weights for the requested stage were not loaded */
} oryon_status;

/* Matching arithmetic (reference utils/pcd.py:195-200 picks fp32 on CPU or fp16 on CUDA). */
typedef enum {
  /* fp16-operand tcgen05 pass (fp32 accumulate) that keeps every candidate within the proven
   * rounding bound of the row maximum, then exact fp32 re-scoring of the candidates: the result
   * equals a full fp32 evaluation (corrs_device='cpu' semantics).  Default. */
  ORYON_MATCH_TC_REFINED = 0,
  /* straight fp32 CUDA-core evaluation of every (anchor,query) pair; slow, used as the device-side
   * cross-check and as the fallback for rows whose candidate list overflowed. */
  ORYON_MATCH_EXACT_FP32 = 1
} oryon_match_mode;

typedef enum { ORYON_DEPTH_I32 = 0, ORYON_DEPTH_F32 = 1, ORYON_DEPTH_I16 = 2, ORYON_DEPTH_U16 = 3 } oryon_depth_dtype;

/* ---- lifecycle --------------------------------------------------------------------------- */

int oryon_abi_version(void);
const char* oryon_last_error(void);

/* Creates the per-device context (workspace, TMA descriptor entry point).  Fails with
 * ORYON_ERR_UNSUPPORTED_DEVICE unless the device is compute capability 10.x. */
int oryon_create(int device, oryon_handle** out);
int oryon_destroy(oryon_handle* h);

/* Bytes of device workspace currently held by the handle (diagnostics). */
int64_t oryon_workspace_bytes(const oryon_handle* h);

/* Per-kernel device timing for benchmarks.  When enabled, every kernel the library launches is bracketed
 * by CUDA events recorded on the caller's stream.  oryon_profile_read waits for the recorded events,
 * returns per kernel id the summed duration (ms) and launch count since the last read, and clears them.
 * Kernel ids: 0 prep_rows, 1 match_tc (tcgen05 similarity + argmax epilogue), 2 refine_rows,
 * 3 exact_rows, 4 mask_to_roi, 5 lift/corrs_to_pcd, 6 pointdsc SC matrix, 7 pointdsc NonLocalNet,
 * 8 pointdsc seeds/kNN/power/Kabsch/fitness, 9 pointdsc refinement, 10 backbone GEMM (tcgen05), 11 attention,
 * 12 normalisations, 13 element-wise / small fused heads, 14 im2col, 15 fused tcgen05 attention, 16 V transpose. */
int oryon_profile_enable(oryon_handle* h, int enable);
int oryon_profile_read(oryon_handle* h, double* total_ms, int64_t* launches, int n_ids);

/* ---- a8: masked nearest-neighbour feature matching ----------------------------------------
 * Replaces, for B independent (anchor, query) pairs at once, the arithmetic of
 *   utils/pcd.py:192-205   (ROI feature gather, pdist 'inv_norm_cosine' :28-29, amin/argmin)
 * of the reference's nn_correspondences.  ROI selection order and the two random draws
 * (utils/pcd.py:184-190, :211; utils/misc.py:242-254) stay with the caller so that torch's
 * nonzero ordering and CPU generator are preserved bit-for-bit.
 *
 *   feat_a, feat_q   DEVICE  float32 planar feature maps [B][D][HW_a], [B][D][HW_q]
 *   roi_a, roi_q     DEVICE  int32 flattened pixel ids (y*W+x) in match order, [B][cap_a] / [B][cap_q];
 *                            NULL means the dense identity list 0..HW-1 (n_* must then be NULL too)
 *   n_a, n_q         HOST    int32 [B] list lengths (<= cap); NULL with dense lists
 *   out_idx          DEVICE  int32 [B][cap_a]: position in the pair's query list of the nearest
 *                            neighbour of anchor row i (lowest position on exact ties, as
 *                            torch.argmin); -1 for rows >= n_a[b] or when n_q[b] == 0
 *   out_dist         DEVICE  float32 [B][cap_a]: 0.5*(1-cos) of that neighbour (fp32)
 */
int oryon_match_nn(oryon_handle* h, const float* feat_a, const float* feat_q, int B, int D, int HW_a, int HW_q,
                   const int32_t* roi_a, const int32_t* roi_q, const int32_t* n_a, const int32_t* n_q,
                   int cap_a, int cap_q, int mode, int32_t* out_idx, float* out_dist, void* stream);

/* Statistics of the last oryon_match_nn call on this handle, valid after the stream work finished
 * (reads a small device counter block: synchronises the stream).
 *   stats[0] rows re-scored from the candidate list   stats[1] candidate chunks re-scored
 *   stats[2] rows that overflowed to the exact fallback   stats[3] kernels launched by the call */
int oryon_match_last_stats(oryon_handle* h, int64_t stats[4], void* stream);

/* Diagnostic: with `enable` != 0 the re-scoring pass of every following oryon_match_nn call on this handle also records, per
 * anchor row, how long its candidate list was (one atomic per row: off by default).  oryon_match_list_hist reads the
 * histogram of the LAST call (synchronises the stream):
 *   hist[0..16]  rows whose lists hold that many 8-column candidate chunks     hist[17] rows that overflowed (exact fallback)
 *   hist[18..25] rows by candidate COLUMNS re-scored in float32: 1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65+
 * What it answers: on smooth feature maps (the reference's decoder outputs) how far the tensor-core
 * candidate filter is from its best case of one column per row. */
int oryon_match_set_hist(oryon_handle* h, int enable);
int oryon_match_list_hist(oryon_handle* h, int64_t hist[26], void* stream);

/* HOST-ONLY diagnostic (no device work, no handle): the work decomposition oryon_match_nn uses for its tensor-core pass on
 * a device with `sm_count` SMs.  The units (pair, 256-row anchor block, 128-column query tile) of the batch are distributed
 * over the CTAs of the persistent kernel as lists of segments: whole row blocks round robin for the full waves, the rest cut
 * into per-CTA tile quotas that level the finishing times (kind 0, the default; 1 = one contiguous range per CTA, 2 = whole
 * row blocks only; the environment variable ORYON_MATCH_PLAN=contiguous|whole selects 1 / 2 inside oryon_match_nn).
 * kind | 0x100: the decomposition of the CTA-PAIR kernel (tcgen05 cta_group::2, selected by ORYON_MATCH_PAIR=1): units are
 * (pair, 256-row anchor block, 256-column query tile) distributed over sm_count / 2 CTA pairs; begin_out then has
 * sm_count / 2 + 1 entries.  Without the flag: the default single-CTA kernel, 128-column tiles over sm_count CTAs.
 *   n_a, n_q   HOST int32 [B] list lengths          seg_cap   capacity of segs_out in segments (0: only count)
 *   segs_out   HOST int32 [seg_cap][5]: pair, row block, first tile, end tile, candidate-list slot of the row block
 *   begin_out  HOST int32 [sm_count + 1]: first segment of CTA c in [c], total in [grid]
 *   info_out   HOST int32 [3]: grid, candidate lists per row (max slot + 1), number of segments
 * Returns 0, or ORYON_ERR_INVALID_ARGUMENT when seg_cap is too small (info_out is still filled). */
int oryon_match_plan(const int32_t* n_a, const int32_t* n_q, int B, int sm_count, int kind, int32_t* segs_out, int seg_cap,
                     int32_t* begin_out, int32_t info_out[3]);

/* Row-major compaction of a [B][HW] mask into pixel-id lists: ids of elements equal to `value`
 * in ascending order == torch.nonzero(mask == value) (reference utils/pcd.py:184-185).
 *   mask    DEVICE int32 [B][HW]      roi_out DEVICE int32 [B][HW]      n_out DEVICE int32 [B] */
int oryon_mask_to_roi(oryon_handle* h, const int32_t* mask, int B, int HW, int value, int32_t* roi_out,
                      int32_t* n_out, void* stream);

/* ---- a9 + a10: coordinate scaling, bounds test, truncation, 2D->3D lifting -----------------
 * Replaces pipeline.py:447-460: utils/coordinates.py:5-13 scale_coords (float32 multiply by the
 * Python-float ratio), :36-47 get_valid_coords, the .to(long) truncation, utils/pcd.py:35-81
 * lift_pcd on the selected pixels and the division by 1000.
 *
 *   corrs            DEVICE  int64 [n][4] (y1,x1,y2,x2) in feature-map coordinates
 *   depth_a/q        DEVICE  [Ha][Wa] / [Hq][Wq] of `depth_dtype` (millimetres)
 *   cam_a, cam_q     HOST    float64 [9] row-major intrinsics
 *   pcd_a, pcd_q     DEVICE  float32 [n][3] metres; only the first *n_valid rows are written, rows
 *                            keep their order (stable compaction of the bounds mask)
 *   n_valid          DEVICE  int32 [1]
 */
int oryon_corrs_to_pcd(oryon_handle* h, const int64_t* corrs, int n, int feat_h, int feat_w, const void* depth_a,
                       const void* depth_q, int depth_dtype, int Ha, int Wa, int Hq, int Wq, const double* cam_a,
                       const double* cam_q, float* pcd_a, float* pcd_q, int32_t* n_valid, void* stream);

/* Batched form of the tail of nn_correspondences plus the function above, for B pairs in one launch: the
 * row selection utils/pcd.py:207-212 (roi1[valid], roi2[argmin][valid], final_corrs[idxs]) followed by
 * pipeline.py:447-460.  The caller performs the random draws (they stay on torch's generator) and passes, per pair,
 * the positions in the anchor ROI list of the max_corrs selected rows.
 *   rows            DEVICE  int32 [B][n]  positions in pair b's anchor list; rows[b][0] < 0: pair has no correspondences
 *   roi_a, roi_q    DEVICE  int32 [B][cap_a] / [B][cap_q]  pixel-id lists given to oryon_match_nn
 *   nn_idx          DEVICE  int32 [B][cap_a]               out_idx of oryon_match_nn
 *   depth_a/q       DEVICE  [B][Ha][Wa] / [B][Hq][Wq] of `depth_dtype` (one frame size per call)
 *   cams_a, cams_q  HOST    float64 [B][9]
 *   corrs           DEVICE  int64 [B][n][4]  (y1,x1,y2,x2) feature-map coordinates == nn_correspondences' result
 *   pcd_a, pcd_q    DEVICE  float32 [B][n][3] metres, first n_valid[b] rows of pair b written
 *   n_valid         DEVICE  int32 [B]  (-1 for pairs without correspondences) */
int oryon_select_lift(oryon_handle* h, const int32_t* rows, int B, int n, const int32_t* roi_a, const int32_t* roi_q,
                      const int32_t* nn_idx, int cap_a, int cap_q, int feat_h, int feat_w, const void* depth_a, const void* depth_q,
                      int depth_dtype, int Ha, int Wa, int Hq, int Wq, const double* cams_a, const double* cams_q, int64_t* corrs,
                      float* pcd_a, float* pcd_q, int32_t* n_valid, void* stream);

/* utils/pcd.py:35-81 lift_pcd with xy_idxs: out[i] = ((x-cx)*z/fx, (y-cy)*z/fy, z), z = depth[y][x].
 *   xs, ys DEVICE int64 [n]; out DEVICE float32 [n][3] in depth units (the caller divides by 1000). */
int oryon_lift_pcd(oryon_handle* h, const void* depth, int depth_dtype, int H, int W, const double* cam,
                   const int64_t* xs, const int64_t* ys, int n, float* out, void* stream);

/* ---- a11: PointDSC registration ------------------------------------------------------------
 * Replaces utils/pointdsc/init.py:10-29 get_pointdsc_pose + models/pointdsc/PointDSC.py:128-197
 * PointDSC.forward in test mode (spatial-consistency matrix :150-153, NonLocalNet :48-77, confidence MLP
 * :171, pick_seeds :199-217, cal_seed_trans :234-336 with knn common.py:48-69, power iteration :338-358,
 * weighted Kabsch common.py:7-45 -- the 3x3 SVD runs on the device, not through H.cpu() --, hypothesis
 * scoring, post_refinement :403-438), for P independent correspondence sets in one call.
 *
 * oryon_pointdsc_config mirrors the constructor arguments the reference reads from the snapshot's
 * config.json (utils/pointdsc/init.py:37-50).  `nms_radius` is what init.py:49 feeds from
 * config.inlier_threshold; `inlier_threshold` is the model attribute, which init.py leaves at the
 * constructor default 0.10 (PointDSC.py:87); sigma is the learnt nn.Parameter `sigma` (state_dict),
 * sigma_d the buffer `sigma_spat`. */
typedef struct {
  int32_t in_dim;          /* 6 */
  int32_t num_layers;      /* 12 in the 3DMatch release */
  int32_t num_channels;    /* 128 (the kernels are specialised for it) */
  int32_t num_iterations;  /* power iterations, 10 */
  int32_t k;               /* neighbourhood size, 40 (<= 64) */
  int32_t reserved;
  double ratio;            /* seeds = int(n * ratio) */
  double sigma_d;          /* sigma_spat */
  double sigma;            /* state_dict['sigma'] */
  double nms_radius;
  double inlier_threshold;
} oryon_pointdsc_config;

/* Loads (replaces) the PointDSC weights of this handle.
 *   weights  HOST float32, the state_dict tensors flattened and concatenated in this order
 *            (C = num_channels, H = C/2; BN = weight, bias, running_mean, running_var):
 *              encoder.layer0.{weight [C][in_dim], bias [C]}
 *              for i in 0..num_layers-1:
 *                encoder.blocks.PointCN_layer_i.0.{weight [C][C], bias}, .1 BN [C]
 *                encoder.blocks.NonLocal_layer_i.fc_message.0.{weight [H][C], bias}, .1 BN [H],
 *                  .3.{weight [H][H], bias}, .4 BN [H], .6.{weight [C][H], bias}
 *                encoder.blocks.NonLocal_layer_i.projection_q/.projection_k/.projection_v {weight [C][C], bias}
 *              classification.0.{weight [32][C], bias}, .2.{weight [32][32], bias}, .4.{weight [1][32], bias}
 *            Eval-mode BatchNorm is folded into the preceding convolution at load time. */
int oryon_pointdsc_load(oryon_handle* h, const oryon_pointdsc_config* cfg, const float* weights, int64_t n_floats, void* stream);

/* Optional intermediate outputs for parity tests; every pointer may be NULL.  All DEVICE. */
typedef struct {
  float* conf;            /* [P][cap]          confidence logits (PointDSC.py:171) */
  float* features;        /* [P][cap][C]       L2-normalised correspondence features (:156) */
  int32_t* seeds;         /* [P][seeds_cap]    seed indices in rank order (:174) */
  int32_t* fitness;       /* [P][seeds_cap]    inlier count of each seed hypothesis (:331) */
  int32_t seeds_cap;
  int32_t reserved;
  float* initial_trans;   /* [P][16]           best hypothesis before post-refinement (:335) */
  int32_t* best_seed;     /* [P]               argmax position among the seeds */
} oryon_pointdsc_debug;

/*   src, tgt   DEVICE float32 [P][cap][3] corresponding 3-D points in metres (pcd1, pcd2 of init.py:10)
 *   n          HOST   int32 [P] correspondences per pair, 2 <= n[p] <= cap <= 2048, int(n*ratio) >= 1
 *   out_T      DEVICE float32 [P][16] row-major 4x4 final_trans (PointDSC.py:186) */
int oryon_pointdsc_pose(oryon_handle* h, const float* src, const float* tgt, const int32_t* n, int P, int cap, float* out_T,
                        const oryon_pointdsc_debug* debug, void* stream);

/* ---- a1-a6 building block: the tensor-core GEMM every linear / convolution / attention-score op of the
 * backbone runs on (oryon_b200/csrc/gemm.cu), exposed for parity tests and kernel benchmarks.
 *   out[b][m][n] = residual[b][m][n] + act(alpha * sum_k A[b][m][k] * W[b][n][k] + bias[n])
 * i.e. torch.nn.functional.linear(A, W, bias) per batch matrix (the form of nn.Linear / nn.Conv1d(k=1) /
 * nn.MultiheadAttention projections used by models/vlm.py:46-56, models/fusion.py:83-85, ...).
 *   A DEVICE float32 [batch][M][K], W DEVICE float32 [batch][N][K], bias DEVICE float32 [N] or NULL,
 *   residual DEVICE float32 [batch][M][N] or NULL, out DEVICE float32 [batch][M][N]
 *   act: 0 none, 1 QuickGELU x*sigmoid(1.702x) (CLIP), 2 GELU(erf) (timm Mlp), 3 ReLU
 *   precision: 3 = fp16 split pairs, three tcgen05 products, float32-equivalent;
 *              2 = one fp16 product + both cross terms in one 8-bit (e5m2 x e4m3) product of the same depth, K % 64 == 0: two tensor-pipe
 *                  units instead of three, ~1.7e-5 relative RMS error per product against 2.9e-4 for one product and 7e-8 for three
 *                  (oryon_b200/csrc/gemm.cuh);
 *              1 = single fp16 product (run_test.py:14 'medium' precision class) */
int oryon_gemm_f32(oryon_handle* h, const float* A, const float* W, const float* bias, const float* residual, float* out, int M, int N,
                   int K, int batch, int act, float alpha, int precision, void* stream);

/* ---- a1-a6: the network, Oryon.forward (net.py:142-167) ------------------------------------------
 * Weights are handed over by their reference state_dict names (the keys of FPM_Pipeline's checkpoint with the
 * leading "model." stripped, i.e. Oryon's own: vlm.clip_model.*, guidance_backbone.features.*, fusion.*,
 * decoder.*; net.py:99-139 shows how the reference fills them), float32, in the reference's shapes.  Integer
 * buffers (relative_position_index, attn_mask) are derived by the library and need not be passed.
 * oryon_backbone_finalize packs them (fp16 split pairs, K-major, conv weights permuted to the im2col order)
 * and frees the host copies. */
typedef struct {
  int32_t vis_layers;          /* CLIP vision transformer depth: 24 for ViT-L/14@336 (vlm.py:19) */
  int32_t txt_layers;          /* CLIP text transformer depth: 12; 0 = text tower not loaded */
  int32_t precision;           /* GEMM precision, see oryon_gemm_f32: 3 = three products in every GEMM; 2 (default of the Python mirror) = three
                                * products everywhere except the four linear layers of every CLIP vision block, which run with fp8 cross terms;
                                * 1 = one product everywhere (NOT a parity mode) */
  int32_t max_pairs_per_pass;  /* pairs processed per pass over the network (bounds activation memory); 0 = 16 */
} oryon_backbone_config;

int oryon_backbone_set_weight(oryon_handle* h, const char* name, const float* data /*HOST*/, int64_t numel);
int oryon_backbone_finalize(oryon_handle* h, const oryon_backbone_config* cfg, void* stream);

/* CLIPEncoder.encode_prompt after tokenisation (models/vlm.py:74-83):
 *   tokens DEVICE int32 [n][77] (SimpleTokenizer ids; EOT is the arg-max id) -> out DEVICE float32 [n][768].
 * A pure function of the prompt strings: the caller caches it per object. */
int oryon_text_forward(oryon_handle* h, const int32_t* tokens, int n, float* out, void* stream);

/* Optional intermediate outputs (all DEVICE float32, NCHW, 2B images: anchors [0,B), queries [B,2B)). */
typedef struct {
  int32_t B;
  int32_t reserved;
  float* clip_tokens;   /* [2B][1024][24][24]  encode_image output (vlm.py:58-61) */
  float* guid1;         /* [2B][512][24][24]   get_guidance_embeds (net.py:60-75) */
  float* guid2;         /* [2B][256][48][48] */
  float* guid3;         /* [2B][128][96][96] */
  float* fusion;        /* [2B][128][24][24]   ImageTextFusion.forward output (T = 1 squeezed) */
} oryon_backbone_debug;

/* Oryon.forward for B pairs:
 *   rgb_a, rgb_q   DEVICE float32 [B][3][224][224] in [0,1]        (xs['anchor'|'query']['rgb'])
 *   text_emb       DEVICE float32 [B][80][768]                      (encode_prompt output, oryon_text_forward)
 *   featmap_a/q    DEVICE float32 [B][32][192][192]                 ('featmap_a', 'featmap_q')
 *   mask_a/q       DEVICE float32 [B][1][192][192] logits           ('mask_a', 'mask_q') */
int oryon_backbone_forward(oryon_handle* h, const float* rgb_a, const float* rgb_q, int B, const float* text_emb, float* featmap_a,
                           float* featmap_q, float* mask_a, float* mask_q, const oryon_backbone_debug* debug, void* stream);

/* GEMM accounting since the last call: launches, algorithmic FLOPs 2*M*N*K, and (may be NULL) the tensor-pipe work actually issued in
 * fp16-equivalent FLOPs -- 3x / 2x / 1x the algorithmic ones for a GEMM at precision 3 / 2 / 1.  Resets the counters. */
int oryon_gemm_counters(oryon_handle* h, int64_t* launches, double* flops, double* tensor_flops);

/* ---- a7: mask post-processing ----------------------------------------------------------------------
 * Replaces the no-grad part of FeatureLoss.mask_loss (losses.py:56-60: torch.where(sigmoid(logits) > mask_th, 1, 0)
 * and mask_iou, utils/metrics.py:18-40, against the nearest-resized ground truth, losses.py:52-53) and the
 * counting / external-mask resize of is_detection_valid and get_featmap_corrs (pipeline.py:372-395, :409-412).
 *   logits      DEVICE float32 [B][H][W] or NULL (external mask only)
 *   gt          DEVICE uint8   [B][Hg][Wg] or NULL (no IoU)
 *   pred_mask   DEVICE int32   [B][H][W]  (with logits)
 *   gt_resized  DEVICE int32   [B][H][W]  or NULL  nearest resize of gt (F.interpolate mode='nearest')
 *   n_pred      DEVICE int32   [B] or NULL  pixels with pred_mask == 1
 *   n_gt        DEVICE int32   [B] or NULL  pixels with gt_resized == 1
 *   iou         DEVICE float32 [B] or NULL  |pred & gt| / |pred | gt|  (NaN for 0/0, as the reference) */
int oryon_mask_postproc(oryon_handle* h, const float* logits, int B, int H, int W, float mask_th, const uint8_t* gt, int Hg, int Wg,
                        int32_t* pred_mask, int32_t* gt_resized, int32_t* n_pred, int32_t* n_gt, float* iou, void* stream);

/* ---- N1: batch staging (decoded frames -> network inputs) -------------------------------------------------
 * Replaces, for B frames in one launch, preprocess_item's rgb / 255. and mask == mask_id (utils/data/common.py:48-49,
 * :62-64), the test-time resize (utils/augmentations.py:129-141 under the reference's pinned torchvision 0.13: bilinear
 * without antialiasing for rgb, nearest for the mask) and CollateWrapper's casts (datasets.py:205-207).
 *   rgb       DEVICE uint8 [B][H][W][3] decoded frames            mask  DEVICE uint8 or int32 [B][H][W] label image, or NULL
 *   mask_ids  DEVICE int32 [B] label of the object in each frame, or NULL (label 1)
 *   out_rgb   DEVICE float32 [B][3][out_h][out_w] in [0,1]        out_mask DEVICE uint8 [B][out_h][out_w] in {0,1} */
int oryon_stage_inputs(oryon_handle* h, const uint8_t* rgb, const void* mask, int mask_is_i32, const int32_t* mask_ids, int B, int H,
                       int W, int out_h, int out_w, float* out_rgb, uint8_t* out_mask, void* stream);

/* ---- N2: pose-error metrics of the evaluator ---------------------------------------------------------
 * Replaces the arithmetic of Evaluator.register_eval (utils/evaluator.py:206-288) for P pose pairs per call:
 * compute_RT_distances (utils/metrics.py:236-259), compute_add / compute_adds (:205-234, float16 model transform of
 * utils/pcd.py:127-133), my_mssd / my_mspd (bop_toolkit_lib/pose_error.py:370-426, float16-rounded poses, first three
 * model points -- see csrc/eval.cu for the restated quirks).  Thresholding and bookkeeping stay with the caller
 * (oryon_b200/utils/evaluator.py).
 *
 * oryon_eval_set_object registers (or replaces) one object model on the handle's device (synchronous copies):
 *   pts   HOST float64 [n][3] model points in mm        syms  HOST float64 [n_sym][3][4] symmetry transforms [R | t]
 * oryon_eval_pose_errors:
 *   obj_ids HOST int32 [P]; pred, gt DEVICE float64 [P][16] row-major 4x4 (metres; `pred` after the zero-pose
 *   substitution of evaluator.py:222-223); cams DEVICE float64 [P][9]
 *   out DEVICE float64 [P][6]: R error (deg), T error (cm), ADD or ADD-S (m), 1.0 if ADD-S was used (n_sym > 1),
 *   MSSD (mm), MSPD (px) */
int oryon_eval_set_object(oryon_handle* h, int obj_id, const double* pts, int n, const double* syms, int n_sym);
int oryon_eval_pose_errors(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams,
                           double* out, void* stream);

/* VSD of the evaluator (bop_toolkit_lib/pose_error.py:17-96 as called from utils/evaluator.py:279-286: delta, taus,
 * normalisation by the object diameter, 'step' cost, visibility mode 'bop19').  The depth images of the model in the two
 * poses come from a z-buffer rasteriser that follows the conventions of bop_toolkit_lib/renderer_vispy.py (sample of
 * pixel (r, c) at (c + 0.5, r + 0.5), nearest surface, no culling); the reference renders with OpenGL, so this one step
 * has no reference output to be compared with.
 *   oryon_eval_set_object_mesh: faces HOST int32 [n_faces][3] vertex indices into the points given to oryon_eval_set_object
 *   oryon_eval_vsd: pred / gt / cams as for oryon_eval_pose_errors; depth_test DEVICE [P][H][W] int32 (depth_is_f32 = 0) or
 *   float32 (1), millimetres, 0 = missing; taus HOST float64 [n_tau <= 16]; diameters HOST float64 [P] (mm);
 *   out DEVICE float64 [P][n_tau] */
int oryon_eval_set_object_mesh(oryon_handle* h, int obj_id, const int32_t* faces, int n_faces);
int oryon_eval_vsd(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams,
                   const void* depth_test, int depth_is_f32, int H, int W, double delta, const double* taus, int n_tau,
                   const double* diameters, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ORYON_B200_H_ */
