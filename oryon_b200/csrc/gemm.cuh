// Tensor-core GEMM used by every linear / convolution / attention-score op of the backbone (a1-a6):
//
//     D[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * W[b][n][k] )
//
// Operands are fp16 "split pairs": x = hi + lo with hi = fp16(x), lo = fp16(x - hi), both K-major with the
// K extent padded to 64 elements (one 128-byte swizzle atom).  precision = 3 evaluates hi*hi + lo*hi + hi*lo
// with fp32 accumulation in TMEM (about 2^-21 relative operand error: float32-equivalent for the parity
// gate); precision = 1 evaluates hi*hi only (what the reference's float32_matmul_precision('medium')
// permits, run_test.py:14).  See gemm.cu for the kernel.
//
// precision = 2 ("fp8 cross terms") evaluates hi*hi on fp16 and BOTH cross terms in one 8-bit product of the same depth: the `lo`
// matrix of an operand then holds, per 64-deep K block, 128 bytes of 8-bit values instead of 64 fp16 residuals --
//     activation row:  [ e5m2(x * 2^-4)  x 64 | e5m2((x - hi) * 2^7) x 64 ]
//     weight row:      [ e4m3((w - hi) * 2^4) x 64 | e4m3(w * 2^-7)  x 64 ]        (w pre-scaled, see below)
// so that A8 . W8 over the 128 bytes is x*(w - hi_w) + (x - hi_x)*w.  Same bytes per row, same tensor maps, same shared-memory
// stage as the fp16 residuals; the tensor pipe spends 2 units per product instead of 3 (kind::f8f6f4 runs 32 K elements per
// instruction where kind::f16 runs 16).  The cross terms are 2^-11 of the product and need only a few bits: their error is
// ~2^-15 relative per element (2.9e-4 relative RMS for one fp16 product, 1.7e-5 for this form, 7e-8 for three fp16 products;
// tools/f8_cross_sim.py).  Activations take e5m2 (two mantissa bits, range 2^-16 .. 57344: no scale to choose for tensors of any
// magnitude), weights e4m3 behind a per-tensor power-of-two scale fixed at load: W' = W * 2^k with max|W'| in (2^14, 2^15], the
// GEMM runs on W' (hi16 = fp16(W')) and its epilogue multiplies by 2^-k (exact).  Requires K % 64 == 0.
#pragma once

#include <cuda_fp8.h>

#include "common.cuh"

namespace oryon {
namespace gemm {

enum Act { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU = 2, ACT_RELU = 3 };
// What the `lo` matrix of a split pair holds: fp16 residuals / the 8-bit cross-term blocks of an activation (A operand) /
// LO_QKV: the output of a fused Q | K | V projection for the attention kernel, which multiplies two ACTIVATIONS: columns [0, w) (Q) get
// the A-operand blocks, columns [w, 2w) (K) the B-operand blocks [e5m2((k - hi) * 2^4) x 64 | e5m2(k * 2^-7) x 64], columns [2w, 3w)
// (V) stay fp16 residuals (w = Epilogue::qkv_width, a multiple of 64).
enum LoFormat { LO_F16 = 0, LO_F8X = 1, LO_QKV = 2 };

constexpr float kF8ActHi = 0.0625f, kF8ActLo = 128.f;      // activation block scales (2^-4, 2^7)
constexpr float kF8WLo = 16.f, kF8WHi = 0.0078125f;         // weight block scales (2^4, 2^-7); products: 2^-4 * 2^4 = 2^7 * 2^-7 = 1

#ifdef __CUDACC__
// byte offset, inside a row of the `lo` matrix, of the FIRST 8-bit value of column c; the second one lies 64 bytes further
__device__ __forceinline__ int64_t f8x_off(int c) { return (int64_t)(c >> 6) * 128 + (c & 63); }
__device__ __forceinline__ uint32_t pack4_f8(float a, float b, float c, float d, __nv_fp8_interpretation_t fmt) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, fmt);   // .x -> low byte
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, fmt);
  return lo | (hi << 16);
}
// The 8-bit cross-term words of four consecutive ACTIVATION values (already clamped to the fp16 range).
__device__ __forceinline__ void f8x_act4(float x0, float x1, float x2, float x3, uint32_t& first, uint32_t& second) {
  const float r0 = x0 - __half2float(__float2half_rn(x0)), r1 = x1 - __half2float(__float2half_rn(x1));
  const float r2 = x2 - __half2float(__float2half_rn(x2)), r3 = x3 - __half2float(__float2half_rn(x3));
  first = pack4_f8(x0 * kF8ActHi, x1 * kF8ActHi, x2 * kF8ActHi, x3 * kF8ActHi, __NV_E5M2);
  second = pack4_f8(r0 * kF8ActLo, r1 * kF8ActLo, r2 * kF8ActLo, r3 * kF8ActLo, __NV_E5M2);
}
// The B-operand form of four consecutive activation values (the K side of Q K^T): same scales as a weight row, e5m2.
__device__ __forceinline__ void f8x_actb4(float x0, float x1, float x2, float x3, uint32_t& first, uint32_t& second) {
  const float r0 = x0 - __half2float(__float2half_rn(x0)), r1 = x1 - __half2float(__float2half_rn(x1));
  const float r2 = x2 - __half2float(__float2half_rn(x2)), r3 = x3 - __half2float(__float2half_rn(x3));
  first = pack4_f8(r0 * kF8WLo, r1 * kF8WLo, r2 * kF8WLo, r3 * kF8WLo, __NV_E5M2);
  second = pack4_f8(x0 * kF8WHi, x1 * kF8WHi, x2 * kF8WHi, x3 * kF8WHi, __NV_E5M2);
}
// ... stored into row `lo_row` (pointer to the row's first element) at columns c .. c+3, c % 4 == 0
__device__ __forceinline__ void store_f8x_act4(__half* lo_row, int c, float x0, float x1, float x2, float x3) {
  uint32_t a, b;
  f8x_act4(x0, x1, x2, x3, a, b);
  uint8_t* p = reinterpret_cast<uint8_t*>(lo_row) + f8x_off(c);
  *reinterpret_cast<uint32_t*>(p) = a;
  *reinterpret_cast<uint32_t*>(p + 64) = b;
}
#endif

// One operand: a (possibly batched, possibly strided) K-major matrix of fp16 split pairs.
struct Operand {
  const __half* hi = nullptr;
  const __half* lo = nullptr;   // may be null when precision == 1; precision 2: the 8-bit cross-term blocks (LO_F8X)
  int64_t ld = 0;               // elements between consecutive rows
  int64_t stride_b0 = 0;        // elements between consecutive inner-batch matrices
  int64_t stride_b1 = 0;        // elements between consecutive outer-batch matrices
};

struct Epilogue {
  float alpha = 1.f;
  const float* bias = nullptr;        // [N]
  int act = ACT_NONE;
  const float* residual = nullptr;    // fp32, same addressing as out32 (added after the activation)
  float* out32 = nullptr;             // fp32 output, row stride ld32
  int64_t ld32 = 0;
  __half* out_hi = nullptr;           // split-pair output (feeds the next GEMM), row stride ldh
  __half* out_lo = nullptr;
  int64_t ldh = 0;
  const int32_t* row_map = nullptr;   // optional [M]: destination row of GEMM row m (-1: drop); batch 1 only
  int64_t out_b0 = 0, out_b1 = 0;     // batch strides (elements) of out32 / residual
  int64_t outh_b0 = 0, outh_b1 = 0;   // batch strides of out_hi / out_lo
  int transpose_h = 0;                // write out_hi/out_lo transposed: element (m, n) at n * ldh + m
  int lo_format = LO_F16;             // LO_F8X: out_lo receives the activation cross-term blocks (the consumer runs at precision 2)
  int qkv_width = 0;                  // LO_QKV: width of each of the three column ranges
};

// Implicit-GEMM A operand of a k x k convolution (stride 1, zero padding k / 2) over NHWC fp32 activations: row m is output
// pixel (n, y, x), column (tap * (C0 + C1) + c) is input channel c of tap (ty, tx) -- the column order of the im2col matrix
// and of the packed conv weights.  Channels come from src0 ([n][H][W][C0], or through a pixel-shuffle view of a
// ConvTranspose(k=2, s=2) GEMM output [n][H/2][W/2][2][2][C0] when shuffle0) followed by src1 ([n][H][W][C1], may be null).
struct ConvGather {
  const float* src0 = nullptr;
  int C0 = 0, shuffle0 = 0;
  const float* src1 = nullptr;
  int C1 = 0;
  int n = 0, H = 0, W = 0, k = 3;
};

struct Problem {
  int M = 0, N = 0, K = 0;      // K = logical depth; operands are readable up to Kpad = round_up(K, 64)
  int nb0 = 1, nb1 = 1;         // batch extents
  int precision = 3;            // 1, 2 (both `lo` matrices in LO_F8X form) or 3
  Operand A, W;
  Epilogue ep;
  const ConvGather* gather = nullptr;   // when set, A is produced on the fly from the activations (A.hi / A.lo unused); unbatched
};

// Enqueues the GEMM on `st`.  Returns an oryon_status.
int launch(oryon_handle* h, const Problem& p, cudaStream_t st);

// fp32 [rows][cols] (row stride ld_in) -> split pair [rows][ld_out] with zero K padding up to ld_out.
int split_rows(oryon_handle* h, const float* in, int64_t ld_in, int rows, int cols, __half* hi, __half* lo, int64_t ld_out,
               cudaStream_t st);
// The same for a precision-2 operand (cols % 64 == 0, ld_out == cols): `lo` receives the 8-bit cross-term blocks.  is_weight: the rows
// are multiplied by `scale` (a power of two, weight_scale()) first and packed in the weight form; otherwise the activation form.
int split_rows_f8x(oryon_handle* h, const float* in, int64_t ld_in, int rows, int cols, __half* hi, __half* lo, int64_t ld_out, bool is_weight,
                   float scale, cudaStream_t st);
// 2^k with max|w| * 2^k in (2^14, 2^15] (1 for an all-zero tensor)
float weight_scale(float absmax);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace gemm
}  // namespace oryon
