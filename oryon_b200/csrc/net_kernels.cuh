// Non-GEMM kernels of the backbone (a1-a6): resize + patch extraction, LayerNorm (with window gather / concat),
// window / sequence attention, embeddings, im2col, GroupNorm, small fused heads.  Activations are fp32,
// token-major ([rows][C], i.e. NHWC for images); every producer of a GEMM operand writes fp16 split pairs
// (hi, lo) with the K extent zero-padded to 64 (gemm.cuh).  Declarations only; see net_kernels.cu.
#pragma once

#include "common.cuh"
#include "gemm.cuh"

namespace oryon {
namespace net {

// ---- K1 + K2a: bicubic resize (a = -0.75) + normalise + non-overlapping patch extraction ------------------
// rgb [n][3][in][in] in [0,1] -> split A [n*grid*grid][ld], column (c, ky, kx) = conv weight flatten order.
// align_corners = 0: CLIP path (torchvision Resize, models/vlm.py:45); 1: Swin path (net.py:67).
int resize_patch(oryon_handle* h, const float* rgb, int n, int in_size, int out_size, int patch, int align_corners, const float mean[3],
                 const float stdv[3], __half* hi, __half* lo, int ld, cudaStream_t st);

// ---- LayerNorm -------------------------------------------------------------------------------------------
struct LnArgs {
  const float* x = nullptr;       // [src_rows][ldx]
  int64_t ldx = 0;
  int C = 0;                      // normalised width (multiple of 32, <= 1024)
  const float* gamma = nullptr;   // null: no affine / no normalisation at all when `raw` is set
  const float* beta = nullptr;
  float eps = 1e-5f;
  int rows = 0;                   // output rows
  const int32_t* row_map = nullptr;  // output row r reads source row row_map[r]; -1 -> a zero row (padding AFTER the norm)
  const float* cat = nullptr;     // optional second source [src_rows][cat_C], appended un-normalised after the C columns
  int cat_C = 0;
  float* out32 = nullptr;         // fp32 output [rows][ld32] (may alias x when there is no row_map)
  int64_t ld32 = 0;
  __half* out_hi = nullptr;       // split output [rows][ldh], zero padded up to ldh
  __half* out_lo = nullptr;
  int64_t ldh = 0;
  int lo_format = 0;              // gemm::LO_F8X: out_lo receives the 8-bit cross-term blocks (C % 64 == 0, no concat, vector path)
};
int layernorm(oryon_handle* h, const LnArgs& a, cudaStream_t st);

// ---- attention over contiguous sequences -------------------------------------------------------------------
struct AttnArgs {
  const float* qkv = nullptr;     // [n_seq*S][ld]; q at column h*d, k at off_k + h*d, v at off_v + h*d
  int64_t ld = 0;
  int off_k = 0, off_v = 0;
  int n_seq = 0, S = 0, heads = 0, d = 0;   // d in {32, 64}
  float scale = 1.f;              // q * scale before the dot product
  int causal = 0;                 // CLIP text tower (upper-triangular -inf mask)
  const float* bias = nullptr;    // [heads][S][S] relative-position bias (torchvision swin)
  const float* bias_t = nullptr;  // optional: bias with the token indices swapped, [heads][key][query] (coalesced for a query per lane)
  const float* mask = nullptr;    // [n_win][S][S] shifted-window mask, window = seq % n_win; symmetric in the two token indices
  int n_win = 1;
  __half* out_hi = nullptr;       // [n_seq*S][ldh] heads concatenated
  __half* out_lo = nullptr;
  int64_t ldh = 0;
};
int attention(oryon_handle* h, const AttnArgs& a, cudaStream_t st);

// ---- pieces of the tensor-core attention path (long sequences: CLIP vision, S = 577) -------------------------------
// row-wise softmax of fp32 scores [rows][ld_in] (S valid columns) -> split probabilities [rows][ld_out], zero padded
int softmax_split(oryon_handle* h, const float* scores, int64_t rows, int S, int ld_in, __half* hi, __half* lo, int ld_out, cudaStream_t st);
// V^T per (sequence, head): split qkv [n_seq*S][ld] (V at column off_v + head*d) -> [n_seq*heads][d][ld_out] with zero padding s >= S
int transpose_v(oryon_handle* h, const __half* qkv_hi, const __half* qkv_lo, int64_t ld, int off_v, int n_seq, int S, int heads, int d,
                __half* vt_hi, __half* vt_lo, int ld_out, cudaStream_t st);

// ---- CLIP embeddings ---------------------------------------------------------------------------------------
// x[n][0] = cls + pos[0]; x[n][1+i] = patch[n*T+i] + pos[1+i]; then ln_pre.  (vlm.py:49-51)
int clip_embed_ln(oryon_handle* h, const float* patch, const float* cls, const float* pos, const float* gamma, const float* beta, int n,
                  int T, int C, float* x, cudaStream_t st);
// x[s][l] = tok_emb[tokens[s][l]] + pos[l]   (vlm.py:74-75)
int text_embed(oryon_handle* h, const int32_t* tokens, const float* tok_emb, const float* pos, int n_seq, int L, int C, int vocab, float* x,
               cudaStream_t st);
// row of the EOT token (first arg-max token id, vlm.py:81) per sequence -> row_map for the final LayerNorm
int eot_rows(oryon_handle* h, const int32_t* tokens, int n_seq, int L, int32_t* rows, cudaStream_t st);

// ---- convolution plumbing ----------------------------------------------------------------------------------
// im2col of an NHWC fp32 image (optionally the channel concatenation of two sources; the first may be the output
// of a 2x2 stride-2 transposed convolution stored as [n][H/2][W/2][4][C0] = "shuffle") into split A
// [n*H*W][ld], column (ky*k + kx)*(C0+C1) + c; zero padding of k/2 around the image.
struct Im2colArgs {
  const float* src0 = nullptr;
  int C0 = 0;
  int shuffle0 = 0;
  const float* src1 = nullptr;
  int C1 = 0;
  int n = 0, H = 0, W = 0, k = 3;
  __half* hi = nullptr;
  __half* lo = nullptr;
  int ld = 0;
};
int im2col(oryon_handle* h, const Im2colArgs& a, cudaStream_t st);

// GroupNorm(groups of 16 channels) + ReLU over NHWC fp32 [n][HW][C], in place (decoder.py:19-25)
int groupnorm_relu(oryon_handle* h, float* x, int n, int HW, int C, const float* gamma, const float* beta, double* stats_scratch,
                   cudaStream_t st);

// PatchMerging gather + LayerNorm(4C) (torchvision swin_transformer._patch_merging_pad + norm): x [n][H][W][C] -> split [n*(H/2)*(W/2)][4C]
int patch_merge_ln(oryon_handle* h, const float* x, int n, int H, int W, int C, const float* gamma, const float* beta, __half* hi, __half* lo,
                   cudaStream_t st);

// rows / max(|row|, 1e-12) -> split   (F.normalize, fusion.py:590-591)
int l2norm_split(oryon_handle* h, const float* x, int rows, int C, __half* hi, __half* lo, int ld, cudaStream_t st);

// fusion.py:617-620: mean over the prompts, renormalise, Linear + ReLU -> [B][out_c]
int text_guidance(oryon_handle* h, const float* text, int B, int P, int C, const float* w, const float* b, int out_c, float* out,
                  cudaStream_t st);

// ClassTransformerLayer with T = 1 (fusion.py:409-434): x [n][24*24][128] updated in place
struct ClassTfW {
  const float *n1_g, *n1_b, *n2_g, *n2_b;
  const float *q_w, *q_b, *k_w, *k_b;   // [128][256]
  const float *v_w, *v_b;               // [128][128]
  const float *m0_w, *m0_b;             // [512][128]
  const float *m2_w, *m2_b;             // [128][512]
};
int class_transformer(oryon_handle* h, float* x, const float* text_guid, int n, int B, const ClassTfW& w, cudaStream_t st);

// head conv3x3 32 -> 1 (+bias) and NHWC -> NCHW transpose of the feature map (decoder.py:97-98)
int decoder_head(oryon_handle* h, const float* x, int n, int H, int W, const float* w, const float* b, float* logits, float* featmap,
                 cudaStream_t st);

// out[r][c] = act(x[r][c])  helper: NHWC fp32 -> NCHW fp32 (guidance outputs for the debug interface)
int nhwc_to_nchw(oryon_handle* h, const float* x, int n, int HW, int C, float* out, cudaStream_t st);

}  // namespace net
}  // namespace oryon
