// Tensor-core GEMM used by every linear / convolution / attention-score op of the backbone (a1-a6):
//
//     D[b][m][n] = epilogue( alpha * sum_k A[b][m][k] * W[b][n][k] )
//
// Operands are fp16 "split pairs": x = hi + lo with hi = fp16(x), lo = fp16(x - hi), both K-major with the
// K extent padded to 64 elements (one 128-byte swizzle atom).  precision = 3 evaluates hi*hi + lo*hi + hi*lo
// with fp32 accumulation in TMEM (about 2^-21 relative operand error: float32-equivalent for the parity
// gate); precision = 1 evaluates hi*hi only (what the reference's float32_matmul_precision('medium')
// permits, run_test.py:14).  See gemm.cu for the kernel.
#pragma once

#include "common.cuh"

namespace oryon {
namespace gemm {

enum Act { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_GELU = 2, ACT_RELU = 3 };

// One operand: a (possibly batched, possibly strided) K-major matrix of fp16 split pairs.
struct Operand {
  const __half* hi = nullptr;
  const __half* lo = nullptr;   // may be null when precision == 1
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
  int precision = 3;            // 1 or 3
  Operand A, W;
  Epilogue ep;
  const ConvGather* gather = nullptr;   // when set, A is produced on the fly from the activations (A.hi / A.lo unused); unbatched
};

// Enqueues the GEMM on `st`.  Returns an oryon_status.
int launch(oryon_handle* h, const Problem& p, cudaStream_t st);

// fp32 [rows][cols] (row stride ld_in) -> split pair [rows][ld_out] with zero K padding up to ld_out.
int split_rows(oryon_handle* h, const float* in, int64_t ld_in, int rows, int cols, __half* hi, __half* lo, int64_t ld_out,
               cudaStream_t st);

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace gemm
}  // namespace oryon
