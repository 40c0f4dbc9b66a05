// a8 -- masked nearest-neighbour feature matching (reference utils/pcd.py:192-205).
//
//   prep_rows_kernel      gather ROI pixels from the planar fp32 map, L2-normalise (fp32), write
//                         fp32 unit rows [N][D4] and fp16 unit rows [Npad][Dpad] (K-major)     HBM-bound
//   match_tc_kernel       S = A * Q^T on tcgen05 (fp16 operands, fp32 accumulate in TMEM), never
//                         materialised: the epilogue keeps, per anchor row, the running maximum and
//                         the list of 8-column chunks whose maximum is within the proven fp16
//                         rounding bound of it                                                  tensor-bound
//   refine_rows_kernel    exact fp32 re-scoring of the listed chunks -> (argmax, distance)
//   exact_rows_kernel     fp32 CUDA-core evaluation of full rows: ORYON_MATCH_EXACT_FP32 mode and the
//                         fallback for rows whose candidate list overflowed
//
// Layouts: rows16_x [B][Npad_x][Dpad] __half, rows32_x [B][Npad_x][D4] float (D4 = D rounded up to 4,
// zero padded).  Candidate lists: cand_m / cand_cnt [B][S][Npad_a], cand_chunk [B][S][Npad_a][CAND_CAP].
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace oryon {
namespace match {

constexpr int kTileM = 128;        // rows per UMMA
constexpr int kRowBlocks = 2;      // UMMA row blocks per CTA (A tile = 256 rows resident in smem)
constexpr int kCtaRows = kTileM * kRowBlocks;
constexpr int kTileN = 128;        // query columns per tile
constexpr int kChunk = 8;          // columns per candidate chunk
constexpr int kCandCap = 16;       // candidate chunks kept per (row, split)
constexpr int kMaxSplits = 8;
// Epilogue warp sets: with 2, set e scans only the tiles that land in TMEM accumulator buffer e (twice the time per scan, one
// candidate list per (row, set), merged by the refine pass like splits).  Measured on B200 at config 2: 2.52 ms with two sets vs
// 2.45 ms with one -- the scan is not what bounds the kernel, so one set is used.
constexpr int kEpiSets = 1;
constexpr int kTcThreads = 64 + 256 * kEpiSets;   // warp 0 TMA, warp 1 MMA, 8 epilogue warps per set
// |fp16-operand score - fp32 score| <= 2^-10 (two roundings of unit-norm operands, Cauchy-Schwarz)
// + accumulation slack; a column can be the fp32 argmax only if its fp16 score is within twice that
// of the fp16 row maximum.
constexpr float kAmbiguity = 2.1e-3f;

struct PairMeta {
  int n_a, n_q;
};

// ------------------------------------------------------------------------------------------------
// prep
// ------------------------------------------------------------------------------------------------
struct PrepArgs {
  const float* feat[2];
  const int32_t* roi[2];
  int hw[2], cap[2], npad[2];
  __half* rows16[2];
  float* rows32[2];
  const PairMeta* meta;
  int D, D4, Dpad;
};

// a.X[side] with a run-time `side` makes the compiler copy the whole argument struct to local memory (112 bytes of stack per
// thread; the dead stack lines are later written back to DRAM: ncu showed 999 MB written for 786 MB of output).  A select
// between the two parameter words keeps the arguments in the constant bank.
#define SIDE(arr) (side == 0 ? (arr)[0] : (arr)[1])

__global__ void __launch_bounds__(256) prep_rows_kernel(PrepArgs a) {
  extern __shared__ float tile[];  // [D][33]
  __shared__ float red[8][32];
  __shared__ float inv_norm[32];
  const int side = blockIdx.z, b = blockIdx.y, r0 = blockIdx.x * 32;
  const int n = side == 0 ? a.meta[b].n_a : a.meta[b].n_q;
  if (r0 >= n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = r0 + lane;
  const int hw = SIDE(a.hw);
  int pix = 0;
  if (r < n) pix = SIDE(a.roi) ? SIDE(a.roi)[(size_t)b * SIDE(a.cap) + r] : r;
  const float* src = SIDE(a.feat) + (size_t)b * a.D * hw + pix;
  float ss = 0.f;
  for (int d = warp; d < a.D; d += 8) {
    const float v = (r < n) ? __ldg(src + (size_t)d * hw) : 0.f;
    tile[d * 33 + lane] = v;
    ss = fmaf(v, v, ss);
  }
  red[warp][lane] = ss;
  __syncthreads();
  if (warp == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][lane];
    // reference: x / norm.clamp_min(eps), eps = 1e-8 (torch cosine_similarity default)
    inv_norm[lane] = fmaxf(sqrtf(s), 1e-8f);
  }
  __syncthreads();
  float* d32 = SIDE(a.rows32) + ((size_t)b * SIDE(a.npad) + r0) * a.D4;
  for (int i = threadIdx.x; i < 32 * a.D4; i += 256) {
    const int row = i / a.D4, d = i - row * a.D4;
    d32[i] = d < a.D ? __fdiv_rn(tile[d * 33 + row], inv_norm[row]) : 0.f;
  }
  __half* d16 = SIDE(a.rows16) + ((size_t)b * SIDE(a.npad) + r0) * a.Dpad;
  for (int i = threadIdx.x; i < 32 * a.Dpad; i += 256) {
    const int row = i / a.Dpad, d = i - row * a.Dpad;
    d16[i] = __float2half_rn(d < a.D ? __fdiv_rn(tile[d * 33 + row], inv_norm[row]) : 0.f);
  }
}

// Dense fast path (no ROI list, HW % 2 == 0): 64 consecutive pixels per CTA, float2 loads (256 contiguous bytes per channel
// row instead of 128), the tile is normalised in place once and then written twice (fp32 rows, fp16 rows) with
// lane <-> channel so that every store instruction covers a contiguous 128 / 64 bytes of one output row.
constexpr int kPrepPix = 64;

__global__ void __launch_bounds__(256) prep_dense_kernel(PrepArgs a) {
  extern __shared__ float tile[];  // [D][kPrepPix + 1]
  __shared__ float red[4][kPrepPix];
  __shared__ float nrm[kPrepPix];
  constexpr int LD = kPrepPix + 1;
  const int side = blockIdx.z, b = blockIdx.y, r0 = blockIdx.x * kPrepPix;
  const int hw = SIDE(a.hw);
  if (r0 >= hw) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = SIDE(a.feat) + (size_t)b * a.D * hw + r0;
  const int npix = min(kPrepPix, hw - r0);
  for (int d = warp; d < a.D; d += 8) {
    float2 v = make_float2(0.f, 0.f);
    if (2 * lane + 1 < npix) v = __ldg(reinterpret_cast<const float2*>(src + (size_t)d * hw) + lane);
    else if (2 * lane < npix) v.x = __ldg(src + (size_t)d * hw + 2 * lane);
    tile[d * LD + 2 * lane] = v.x, tile[d * LD + 2 * lane + 1] = v.y;
  }
  __syncthreads();
  {
    const int p = threadIdx.x & (kPrepPix - 1), part = threadIdx.x / kPrepPix;
    float ss = 0.f;
    for (int d = part; d < a.D; d += 4) ss = fmaf(tile[d * LD + p], tile[d * LD + p], ss);
    red[part][p] = ss;
  }
  __syncthreads();
  if (threadIdx.x < kPrepPix) {
    const int p = threadIdx.x;
    // x / max(|x|, eps) evaluated as x * (1 / max(|x|, eps)): one division per pixel instead of one per element; the unit
    // rows differ from a true division by <= 1 ulp, inside the tolerance of the fp32 re-scoring they feed
    nrm[p] = __fdiv_rn(1.f, fmaxf(sqrtf(red[0][p] + red[1][p] + red[2][p] + red[3][p]), 1e-8f));
  }
  __syncthreads();
  float* d32 = SIDE(a.rows32) + ((size_t)b * SIDE(a.npad) + r0) * a.D4;
  __half* d16 = SIDE(a.rows16) + ((size_t)b * SIDE(a.npad) + r0) * a.Dpad;
  for (int row = warp; row < npix; row += 8) {
    const float inv = nrm[row];
    for (int d = lane; d < a.Dpad || d < a.D4; d += 32) {
      const float v = d < a.D ? tile[d * LD + row] * inv : 0.f;
      if (d < a.D4) d32[(size_t)row * a.D4 + d] = v;
      if (d < a.Dpad) d16[(size_t)row * a.Dpad + d] = __float2half_rn(v);
    }
  }
}

// Dense path, second version (D in {32, 64, 128, 256}): the transposition happens in registers.  A thread loads the same pixel
// pair of four consecutive channels (float2 each, 256 contiguous bytes per warp instruction, every load of the thread issued
// before the first use) and stores the two pixels' channel quads as float4 into a PIXEL-major tile.  After the only barrier
// a warp owns a pixel: one float4 per lane (conflict free), squared norm by warp shuffles, and the unit row leaves as one
// 16-byte fp32 store and one 8-byte fp16 store per lane (512 / 256 contiguous bytes per instruction) instead of 4- and
// 2-byte stores after two more barriers.  Same arithmetic as prep_dense_kernel up to the summation order of the norm.
template <int NIT>   // D = 32 * NIT channels: NIT channel quads per warp
__global__ void __launch_bounds__(256) prep_dense2_kernel(PrepArgs a) {
  extern __shared__ __align__(16) float tile4[];  // [kPrepPix][LD]
  float* tile = tile4;
  constexpr int D = 32 * NIT, LD = D + 4;         // LD / 4 odd: float4 stores of consecutive pixels fall into distinct bank quads
  const int side = blockIdx.z, b = blockIdx.y, r0 = blockIdx.x * kPrepPix;
  const int hw = SIDE(a.hw);
  if (r0 >= hw) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* src = SIDE(a.feat) + (size_t)b * D * hw + r0;
  const int npix = min(kPrepPix, hw - r0);
  const bool in0 = 2 * lane < npix, in1 = 2 * lane + 1 < npix;
  float2 v[NIT][4];
  if (npix == kPrepPix) {   // whole tile (CTA-uniform): straight-line loads
#pragma unroll
    for (int i = 0; i < NIT; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c)   // hw even, r0 even: 8-byte aligned
        v[i][c] = __ldg(reinterpret_cast<const float2*>(src + (size_t)(4 * (warp + 8 * i) + c) * hw) + lane);
  } else {
#pragma unroll
    for (int i = 0; i < NIT; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float* ch = src + (size_t)(4 * (warp + 8 * i) + c) * hw;
        v[i][c] = make_float2(0.f, 0.f);
        if (in1) v[i][c] = __ldg(reinterpret_cast<const float2*>(ch) + lane);
        else if (in0) v[i][c].x = __ldg(ch + 2 * lane);
      }
  }
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    const int d = 4 * (warp + 8 * i);
    *reinterpret_cast<float4*>(&tile[(2 * lane) * LD + d]) = make_float4(v[i][0].x, v[i][1].x, v[i][2].x, v[i][3].x);
    *reinterpret_cast<float4*>(&tile[(2 * lane + 1) * LD + d]) = make_float4(v[i][0].y, v[i][1].y, v[i][2].y, v[i][3].y);
  }
  __syncthreads();
  float* d32 = SIDE(a.rows32) + ((size_t)b * SIDE(a.npad) + r0) * a.D4;   // D4 == D here
  __half* d16 = SIDE(a.rows16) + ((size_t)b * SIDE(a.npad) + r0) * a.Dpad;
  constexpr int Q = (D + 127) / 128;   // float4 quads per lane (1 up to D = 128, 2 at D = 256)
#pragma unroll 4
  for (int row = warp; row < npix; row += 8) {
    float4 x[Q];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int d = 4 * (lane + 32 * q);
      x[q] = d < D ? *reinterpret_cast<const float4*>(&tile[row * LD + d]) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss = fmaf(x[q].x, x[q].x, ss), ss = fmaf(x[q].y, x[q].y, ss), ss = fmaf(x[q].z, x[q].z, ss), ss = fmaf(x[q].w, x[q].w, ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    // x / max(|x|, eps) as x * (1 / max(|x|, eps)), as in prep_dense_kernel
    const float inv = __fdiv_rn(1.f, fmaxf(sqrtf(ss), 1e-8f));
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int d = 4 * (lane + 32 * q);
      const float4 u = make_float4(x[q].x * inv, x[q].y * inv, x[q].z * inv, x[q].w * inv);
      if (d < D) *reinterpret_cast<float4*>(d32 + (size_t)row * D + d) = u;
      if (d < a.Dpad) {   // columns D .. Dpad - 1 (D = 32 with a 64-wide swizzle atom never occurs: Dpad == D for these D)
        const __half2 lo = __floats2half2_rn(u.x, u.y), hi = __floats2half2_rn(u.z, u.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo), pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(d16 + (size_t)row * a.Dpad + d) = pk;
      }
    }
  }
}

#undef SIDE

// ------------------------------------------------------------------------------------------------
// exact fp32 rows (mode ORYON_MATCH_EXACT_FP32 and overflow fallback)
// ------------------------------------------------------------------------------------------------
struct ExactArgs {
  const float* rows32_a;
  const float* rows32_q;
  const PairMeta* meta;
  const int32_t* row_list;   // packed b * npad_a + r, or nullptr = all rows
  const int32_t* n_items_dev;  // number of entries in row_list (device) or nullptr
  int n_items_host;          // used when row_list == nullptr (B * npad_a)
  int npad_a, npad_q, D4, cap_a;
  int32_t* out_idx;
  float* out_dist;
};

// 64 anchor rows per CTA, 256 threads as 16x16, each thread a 4x4 micro-tile, K chunks of 32.
__global__ void __launch_bounds__(256) exact_rows_kernel(ExactArgs a) {
  __shared__ float As[32][64 + 4];
  __shared__ float Qs[32][64 + 4];
  __shared__ int item_row[64];
  __shared__ int item_pair[64];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n_items = a.row_list ? *a.n_items_dev : a.n_items_host;
  for (int t0 = blockIdx.x * 64; t0 < n_items; t0 += gridDim.x * 64) {
    __syncthreads();
    if (threadIdx.x < 64) {
      const int it = t0 + threadIdx.x;
      int packed = -1;
      if (it < n_items) packed = a.row_list ? a.row_list[it] : it;
      int b = -1, r = 0;
      if (packed >= 0) {
        b = packed / a.npad_a;
        r = packed - b * a.npad_a;
        if (r >= a.meta[b].n_a) b = -1;
      }
      item_pair[threadIdx.x] = b;
      item_row[threadIdx.x] = r;
    }
    __syncthreads();
    // all 64 items of a tile may belong to different pairs when row_list is used; the column loop
    // is per pair, so process the distinct pairs present one after the other (usually one).
    int done_pair = -1;
    while (true) {
      int b = 0x7fffffff;
      for (int i = 0; i < 64; ++i) {
        const int p = item_pair[i];
        if (p > done_pair && p < b) b = p;
      }
      if (b == 0x7fffffff) break;
      done_pair = b;
      const int n_q = a.meta[b].n_q;
      float best[4];
      int best_j[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) best[i] = -INFINITY, best_j[i] = -1;
      for (int c0 = 0; c0 < n_q; c0 += 64) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < a.D4; k0 += 32) {
          __syncthreads();
          for (int i = threadIdx.x; i < 64 * 32; i += 256) {
            const int row = i >> 5, k = i & 31;
            float va = 0.f, vq = 0.f;
            if (k0 + k < a.D4) {
              if (item_pair[row] == b) va = a.rows32_a[((size_t)b * a.npad_a + item_row[row]) * a.D4 + k0 + k];
              if (c0 + row < n_q) vq = a.rows32_q[((size_t)b * a.npad_q + c0 + row) * a.D4 + k0 + k];
            }
            As[k][row] = va;
            Qs[k][row] = vq;
          }
          __syncthreads();
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 qv = *reinterpret_cast<const float4*>(&Qs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const float qr[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], qr[j], acc[i][j]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = c0 + tx * 4 + j;
            if (col < n_q && acc[i][j] > best[i]) best[i] = acc[i][j], best_j[i] = col;
          }
      }
      // reduce over the 16 threads (tx) that share rows ty*4..ty*4+3: max value, lowest column on ties
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = best[i];
        int j = best_j[i];
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, v, off, 16);
          const int oj = __shfl_xor_sync(0xffffffffu, j, off, 16);
          if (ov > v || (ov == v && oj >= 0 && (j < 0 || oj < j))) v = ov, j = oj;
        }
        const int row = ty * 4 + i;
        if (tx == 0 && item_pair[row] == b) {
          const size_t o = (size_t)b * a.cap_a + item_row[row];
          a.out_idx[o] = j;
          a.out_dist[o] = j >= 0 ? 0.5f * (-1.f * v + 1.f) : INFINITY;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 similarity + running-max / candidate-chunk epilogue
// ------------------------------------------------------------------------------------------------
// Work decomposition ("stream-K" over query tiles).  The unit of work is one 128-column query tile of one 256-row anchor
// block; all units of the batch, in (pair, row block, tile) order, are cut into gridDim.x contiguous ranges of equal length
// (+-1), so every CTA of the persistent kernel finishes at the same time whatever the task count (2 400 row blocks on 148 SMs
// are 16.2 waves: whole-task round robin idles 116 SMs during a 17th).  A range is a list of segments = (row block, tile
// range); a row block cut by a range boundary is scanned by two (at most `splits`) CTAs, each with its own candidate list
// (slot), merged by the refine pass.  The table is built on the host (list lengths are host-side arguments of the C ABI).
struct Seg {
  int32_t b_slot;   // pair | slot << 24
  int32_t rb;       // 256-row anchor block within the pair
  int32_t j0, j1;   // query tiles [j0, j1)
};
static_assert(sizeof(Seg) == 16, "Seg is loaded as one 16-byte word");

struct TcArgs {
  const PairMeta* meta;
  const Seg* segs;
  const int32_t* seg_begin;   // [gridDim.x + 1]
  int B, npad_a, npad_q, splits;
  float* cand_m;
  int32_t* cand_cnt;
  uint32_t* cand_chunk;
  float ambiguity;
  int dbg_mode;   // timing experiments only (ORYON_MATCH_DEBUG_MODE; results are garbage with bits 1 / 2): 1 = the epilogue releases accumulators
                  // unread, 2 = the producer loads each query stage once per CTA and re-announces it afterwards; 4 = unit-boundary commits
                  // issued at the boundary instead of after the next unit's first MMA (valid results, the A/B switch of that placement)
};

template <int KB_ELEMS, int NUM_KB, int STAGES>
struct TcSmem {
  static constexpr int kRowBytes = KB_ELEMS * 2;
  static constexpr int kABlock = kTileM * kRowBytes;            // one (kb, row block) A tile
  static constexpr int kABytes = kABlock * NUM_KB * kRowBlocks;
  static constexpr int kQStage = kTileN * kRowBytes;            // one k-block of a query tile
  static constexpr int kQBytes = kQStage * STAGES;
  static constexpr int kBarBytes = 8 * (2 * STAGES + 2 + 8) + 16;
  static constexpr int kTotal = 1024 /*align slack*/ + kABytes + kQBytes + kBarBytes;
};

__device__ __forceinline__ float max8(const uint32_t* v) {
  float m = fmaxf(fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), __uint_as_float(v[2]));
  m = fmaxf(fmaxf(m, __uint_as_float(v[3])), __uint_as_float(v[4]));
  m = fmaxf(fmaxf(m, __uint_as_float(v[5])), __uint_as_float(v[6]));
  return fmaxf(m, __uint_as_float(v[7]));
}

template <int KB_ELEMS, int NUM_KB, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_q, TcArgs args) {
  using L = TcSmem<KB_ELEMS, NUM_KB, STAGES>;
  constexpr int kRowBytes = L::kRowBytes;
  constexpr int kKSteps = KB_ELEMS / 16;  // UMMA_K = 16 for 16-bit operands
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kTileM, kTileN, /*fp16*/ 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_q = smem + L::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_q + L::kQBytes);
  uint64_t* full = bars;                 // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* a_full = bars + 2 * STAGES;  // A tile landed
  uint64_t* a_empty = a_full + 1;        // all MMAs of the task retired: A may be overwritten
  // Accumulators are handed over per (buffer, ROW BLOCK): 4 units of 128 lanes x 128 columns.  Round 1 handed over whole buffers (both
  // row blocks at once): at D = 128 the MMAs of a buffer take 1 024 cycles, its read-out by the 8 epilogue warps plus the two barrier
  // hops ~2 000, and the MMA warp waited for a drained buffer 3/4 of the time (profiles/r02_match_pair_vs_1cta.md).  With the
  // row blocks issued one after the other (all k-blocks of row block 0, then of row block 1) and their own full / empty barriers,
  // a unit is filled in 512 cycles and has the 1 536 cycles of the other three units' MMAs to be read out.
  uint64_t* t_full = a_empty + 1;        // [2 buffers][kRowBlocks] accumulator unit ready
  uint64_t* t_empty = t_full + 4;        // [2 buffers][kRowBlocks] accumulator unit drained (4 warps)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(t_empty + 4);
  static_assert(kRowBlocks == 2 && NUM_KB < STAGES, "unit hand-off assumes two row blocks and a ring deeper than one query tile");

  const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_q);
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    ptx::mbar_init(a_full, 1);
    ptx::mbar_init(a_empty, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&t_full[i], 1), ptx::mbar_init(&t_empty[i], 4 * kEpiSets);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_base_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int seg_lo = args.seg_begin[blockIdx.x], seg_hi = args.seg_begin[blockIdx.x + 1];

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, task_iter = 0;
      for (int si = seg_lo; si < seg_hi; ++si) {
        const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);
        const int b = sg.x & 0xFFFFFF, rb = sg.y, j0 = sg.z, j1 = sg.w;
        ptx::mbar_wait(a_empty, (task_iter & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(a_full, L::kABytes);
#pragma unroll
        for (int kb = 0; kb < NUM_KB; ++kb)
#pragma unroll
          for (int r = 0; r < kRowBlocks; ++r)
            ptx::tma_load_2d(smem_a + (kb * kRowBlocks + r) * L::kABlock, &tm_a, a_full, kb * KB_ELEMS,
                             b * args.npad_a + rb * kCtaRows + r * kTileM);
        for (int j = j0; j < j1; ++j) {
          for (int kb = 0; kb < NUM_KB; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            if ((args.dbg_mode & 2) && (phase | task_iter) != 0) {
              ptx::mbar_arrive(&full[stage]);   // experiment: the stage keeps its first contents, no L2 / shared-memory write traffic
            } else {
              ptx::mbar_arrive_expect_tx(&full[stage], L::kQStage);
              ptx::tma_load_2d(smem_q + stage * L::kQStage, &tm_q, &full[stage], kb * KB_ELEMS, b * args.npad_q + j * kTileN);
            }
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        ++task_iter;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // Every lane runs the (warp-uniform) control flow; one elected lane issues.  Descriptors are built once per stage /
    // row block and advanced by adding 2 (= 32 bytes >> 4) per 16-element K step.
    // Commit placement (measured, tools/tmem_ld_bw.cu modes 4-9, profiles/r02_match_commit_placement.md): a tcgen05.commit that is
    // IMMEDIATELY followed by the first MMA of another accumulator costs ~62 tensor cycles (the new accumulator's MMAs do not overlap
    // the tail of the old one across the commit) -- 0.94 -> 0.84 tensor duty for 8-MMA accumulations; the same commit issued after
    // that first MMA costs nothing (0.93).  The commits at a unit boundary (t_full of the finished unit, the release of the query
    // stage after its second use) are therefore deferred past the first MMA of the next unit; they then also cover that MMA, i.e.
    // fire 64 cycles later.  They are flushed at once whenever the warp is about to block, so nothing ever waits on a deferred signal.
    uint32_t stage = 0, phase = 0, task_iter = 0, tile_iter = 0;
    uint64_t* pend_stage = nullptr;   // warp-uniform
    uint64_t* pend_unit = nullptr;
    auto flush = [&](bool leader) {
      if (leader) {
        if (pend_stage) ptx::umma_commit(pend_stage);
        if (pend_unit) ptx::umma_commit(pend_unit);
      }
      pend_stage = pend_unit = nullptr;
    };
    auto wait_flushing = [&](uint64_t* bar, uint32_t parity) {
      if (ptx::mbar_try_wait(bar, parity)) return;
      if (pend_stage || pend_unit) {
        flush(ptx::elect_one());
        __syncwarp();
      }
      ptx::mbar_wait(bar, parity);
    };
    for (int si = seg_lo; si < seg_hi; ++si) {
      const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);   // same address in every lane
      const int j0 = __shfl_sync(0xffffffffu, sg.z, 0), j1 = __shfl_sync(0xffffffffu, sg.w, 0);
      wait_flushing(a_full, task_iter & 1);
      for (int j = j0; j < j1; ++j) {
        const uint32_t buf = tile_iter & 1;
        // row-block-major issue: the NUM_KB stages of this query tile are consumed twice (row block 0, then 1) and released after
        // the second use
#pragma unroll
        for (int r = 0; r < kRowBlocks; ++r) {
          wait_flushing(&t_empty[buf * kRowBlocks + r], ((tile_iter >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * (kRowBlocks * kTileN) + r * kTileN;
          uint32_t st = stage, ph = phase;
          for (int kb = 0; kb < NUM_KB; ++kb) {
            if (r == 0) {
              wait_flushing(&full[st], ph);
              ptx::tc_fence_after();
            }
            const uint64_t dq0 = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem_q + st * L::kQStage), kRowBytes);
            const uint64_t da0 = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem_a + (kb * kRowBlocks + r) * L::kABlock), kRowBytes);
            const bool leader = ptx::elect_one();
#pragma unroll
            for (int k = 0; k < kKSteps; ++k) {
              if (leader) ptx::umma_f16(d_tmem, da0 + 2 * k, dq0 + 2 * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
              if (k == 0 && kb == 0) flush(leader);   // the previous unit's deferred commits, now behind this unit's first MMA
            }
            if (kb == NUM_KB - 1) {
              // unit boundary: defer (an early flush happens if the warp has to block before the next unit's first MMA)
              pend_stage = r == kRowBlocks - 1 ? &empty[st] : nullptr;     // smem slot reusable when these MMAs retire
              pend_unit = &t_full[buf * kRowBlocks + r];                   // this row block's accumulator complete
              if (args.dbg_mode & 4) flush(leader);                        // A/B switch: commit at the boundary (the round-1 placement)
            } else if (r == kRowBlocks - 1 && leader) {
              ptx::umma_commit(&empty[st]);   // mid-unit: followed by MMAs of the same accumulator, free
            }
            __syncwarp();
            if (++st == STAGES) st = 0, ph ^= 1;
          }
          if (r == kRowBlocks - 1) stage = st, phase = ph;
        }
        ++tile_iter;
      }
      {
        const bool leader = ptx::elect_one();
        flush(leader);
        if (leader) ptx::umma_commit(a_empty);
      }
      __syncwarp();
      ++task_iter;
    }
  } else {
    // ============================ epilogue (kEpiSets x 8 warps) ============================
    // The scan of one accumulator (4 TMEM loads + ~240 dependent ALU instructions per warp) takes longer than the MMAs that
    // fill the other buffer.  Two warp sets therefore split the tiles by accumulator buffer: set e owns buffer e, i.e.
    // every second tile, and has two tile periods per scan.  Each (row, set) keeps its own candidate list; the refine
    // pass sees the sets as extra splits (slot = ((b * splits + s) * kEpiSets + set) * npad_a + row).
    const int ew = (warp - 2) & 7;         // 0..7 within the set
    const int set = (warp - 2) >> 3;
    const int rblk = ew >> 2;              // which 128-row block of the CTA tile
    const int quarter = warp & 3;          // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int row_in_cta = rblk * kTileM + quarter * 32 + lane;
    uint32_t tile_iter = 0;
    for (int si = seg_lo; si < seg_hi; ++si) {
      const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);
      const int b = sg.x & 0xFFFFFF, s = sg.x >> 24, rb = sg.y, j0 = sg.z, j1 = sg.w;
      const PairMeta pm = args.meta[b];
      const int row = rb * kCtaRows + row_in_cta;
      const bool row_ok = row < pm.n_a;
      const size_t slot = (((size_t)b * args.splits + s) * kEpiSets + set) * args.npad_a + row;
      uint32_t* my_list = args.cand_chunk + slot * kCandCap;
      float m_run = -INFINITY;
      float thr = row_ok ? -INFINITY : INFINITY;  // rows beyond n_a never record anything
      int cnt = 0;
      for (int j = j0; j < j1; ++j) {
        const uint32_t buf = tile_iter & 1;
        if (kEpiSets > 1 && (int)buf != set) {   // the other set's tile
          ++tile_iter;
          continue;
        }
        ptx::mbar_wait(&t_full[buf * kRowBlocks + rblk], (tile_iter >> 1) & 1);
        ptx::tc_fence_after();
        if (args.dbg_mode & 1) {   // experiment: hand the accumulator back unread
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&t_empty[buf * kRowBlocks + rblk]);
          ++tile_iter;
          continue;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * (kRowBlocks * kTileN) + rblk * kTileN;
        const int col_base = j * kTileN;
        const bool ragged = col_base + kTileN > pm.n_q;
        // two register buffers: the TMEM load of group g + 1 is in flight while group g is scanned
        uint32_t va[32], vb[32];
        auto scan = [&](uint32_t (&v)[32], int g) {
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col_base + g * 32 + i >= pm.n_q) v[i] = 0xff800000u;  // -inf
          }
          // all chunk maxima of the group first (four independent FMNMX chains), ONE test per 32 columns in the common
          // case; the per-chunk bookkeeping runs only when some chunk of the group can enter the band
          float cm[32 / kChunk];
#pragma unroll
          for (int c = 0; c < 32 / kChunk; ++c) cm[c] = max8(v + c * kChunk);
          float gm = cm[0];
#pragma unroll
          for (int c = 1; c < 32 / kChunk; ++c) gm = fmaxf(gm, cm[c]);
          if (gm >= thr) {
#pragma unroll
            for (int c = 0; c < 32 / kChunk; ++c) {
              if (cm[c] >= thr) {
                if (cm[c] > m_run + args.ambiguity) cnt = 0;  // every earlier candidate is now out of range
                if (cnt < kCandCap) {
                  // which columns of the chunk can still win: within the band of the CHUNK maximum (a superset of the
                  // columns within the band of the final row maximum, which is >= cm)
                  const float lim = cm[c] - args.ambiguity;
                  uint32_t bits = 0;
#pragma unroll
                  for (int i = 0; i < kChunk; ++i) bits |= (__uint_as_float(v[c * kChunk + i]) >= lim ? 1u : 0u) << i;
                  my_list[cnt] = static_cast<uint32_t>((col_base + g * 32) / kChunk + c) | (bits << 24);
                }
                ++cnt;
                m_run = fmaxf(m_run, cm[c]);
                thr = m_run - args.ambiguity;
              }
            }
          }
        };
        ptx::tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
        for (int g = 0; g < kTileN / 32; g += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(taddr + (g + 1) * 32, vb);
          scan(va, g);
          ptx::tmem_ld_wait();
          if (g + 2 < kTileN / 32) ptx::tmem_ld_32x32b_x32(taddr + (g + 2) * 32, va);
          scan(vb, g + 1);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[buf * kRowBlocks + rblk]);
        ++tile_iter;
      }
      if (row_ok) args.cand_m[slot] = m_run, args.cand_cnt[slot] = cnt;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair version (cluster of 2, tcgen05 cta_group::2): opt-in (ORYON_MATCH_PAIR=1) -- bit-exact and SLOWER than the single-CTA kernel
// (2.74 vs 2.53 ms at config 2, profiles/r02_match_pair_vs_1cta.md); kept as the record of the experiment and as a tested alternative.
// ------------------------------------------------------------------------------------------------
// The hypothesis it was written to test: ncu of match_tc_kernel (profiles/r01_match_tc_ncu_v4.md) showed the tensor pipe 68.5 % active,
// and the presumed reason was operand bandwidth -- a
// 128x128x16 UMMA reads 8 KB of shared memory in its 64 cycles (128 B/clk, all an SM delivers) while TMA refills 32 B/clk.  Here the
// two SMs of a TPC work as a pair on ONE 256-row anchor block: each CTA keeps ITS 128 anchor rows resident and stages HALF of every
// 256-column query tile; one tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16, issued by the leader CTA) multiplies the pair's
// rows with both halves, reading the peer's half over the inter-SM path.  Per SM and 128 cycles that is 4 KB of A + 4 KB of own
// B + 4 KB of refill instead of 16 KB + 4 KB: shared-memory traffic per flop halves.  Each CTA receives its 128 rows x 256 columns
// in its own tensor memory (two such accumulators = all 512 columns) and scans them with its 8 epilogue warps: warp (quarter,
// half) owns 32 rows x 128 columns of every tile, so a row has two candidate lists per segment (merged by the refine pass like
// the lists of a cut row block).  The anchor tile is double buffered, so the next row block's rows load during the current sweep.
// Synchronisation: both producers report their TMA bytes to the LEADER's `full` / `a_full` barriers (cp.async.bulk.tensor
// .cta_group::2 with a mapa'd barrier address; the leader alone expects the bytes of both halves); tcgen05.commit multicasts
// `empty`, `a_empty` and `t_full` to both CTAs; the 16 epilogue warps of the pair arrive on the leader's `t_empty`.
constexpr int kTileN2 = 256;       // query columns per tile of the pair kernel (128 staged by each CTA)

template <int KB_ELEMS, int NUM_KB, int STAGES>
struct Tc2Smem {
  static constexpr int kRowBytes = KB_ELEMS * 2;
  static constexpr int kABlock = kTileM * kRowBytes;            // one k-block of this CTA's 128 anchor rows
  static constexpr int kABytes = kABlock * NUM_KB;              // one anchor buffer (two are kept)
  static constexpr int kQStage = kTileM * kRowBytes;            // one k-block of this CTA's half (128 rows) of a query tile
  static constexpr int kQBytes = kQStage * STAGES;
  static constexpr int kBarBytes = 8 * (2 * STAGES + 4 + 4) + 16;
  static constexpr int kTotal = 1024 /*align slack*/ + 2 * kABytes + kQBytes + kBarBytes;
};

template <int KB_ELEMS, int NUM_KB, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
match_tc2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_q, TcArgs args) {
  using L = Tc2Smem<KB_ELEMS, NUM_KB, STAGES>;
  constexpr int kRowBytes = L::kRowBytes;
  constexpr int kKSteps = KB_ELEMS / 16;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(2 * kTileM, kTileN2, /*fp16*/ 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // [2][NUM_KB][128 rows]
  uint8_t* smem_q = smem + 2 * L::kABytes;                  // [STAGES][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_q + L::kQBytes);
  uint64_t* full = bars;                 // [STAGES]  both producers -> leader's MMA warp (leader's copy is the live one)
  uint64_t* empty = bars + STAGES;       // [STAGES]  MMA -> both producers (multicast commit)
  uint64_t* a_full = bars + 2 * STAGES;  // [2]       anchor buffer landed in both CTAs (leader's copy)
  uint64_t* a_empty = a_full + 2;        // [2]       all MMAs reading the buffer retired (multicast)
  uint64_t* t_full = a_empty + 2;        // [2]       accumulator ready (multicast)
  uint64_t* t_empty = t_full + 2;        // [2]       accumulator drained by all 16 epilogue warps of the pair (leader's copy)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_a);
    ptx::prefetch_tensormap(&tm_q);
    for (int s = 0; s < STAGES; ++s) ptx::mbar_init(&full[s], 1), ptx::mbar_init(&empty[s], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 1), ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1), ptx::mbar_init(&t_empty[i], 16);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_base_slot, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();      // the peer's barriers are initialised and its tensor memory allocated before anything is signalled
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int pair_id = blockIdx.x >> 1;
  const int seg_lo = args.seg_begin[pair_id], seg_hi = args.seg_begin[pair_id + 1];

  if (warp == 0) {
    // ============================ TMA producer (both CTAs: own anchor rows, own half of the query tiles) ============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_iter = 0;
      for (int si = seg_lo; si < seg_hi; ++si) {
        const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);
        const int b = sg.x & 0xFFFFFF, rb = sg.y, j0 = sg.z, j1 = sg.w;
        const uint32_t abuf = a_iter & 1;
        ptx::mbar_wait(&a_empty[abuf], ((a_iter >> 1) & 1) ^ 1);
        if (rank == 0) ptx::mbar_arrive_expect_tx(&a_full[abuf], 2 * L::kABytes);
        const uint32_t a_bar = ptx::mapa_shared(ptx::smem_u32(&a_full[abuf]), 0);
#pragma unroll
        for (int kb = 0; kb < NUM_KB; ++kb)
          ptx::tma_load_2d_pair(smem_a + abuf * L::kABytes + kb * L::kABlock, &tm_a, a_bar, kb * KB_ELEMS,
                                b * args.npad_a + rb * kCtaRows + (int)rank * kTileM);
        for (int j = j0; j < j1; ++j) {
          for (int kb = 0; kb < NUM_KB; ++kb) {
            ptx::mbar_wait(&empty[stage], phase ^ 1);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * L::kQStage);
            ptx::tma_load_2d_pair(smem_q + stage * L::kQStage, &tm_q, ptx::mapa_shared(ptx::smem_u32(&full[stage]), 0), kb * KB_ELEMS,
                                  b * args.npad_q + j * kTileN2 + (int)rank * kTileM);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
        ++a_iter;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ============================
    if (rank == 0) {
      uint32_t stage = 0, phase = 0, a_iter = 0, tile_iter = 0;
      for (int si = seg_lo; si < seg_hi; ++si) {
        const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);
        const int j0 = __shfl_sync(0xffffffffu, sg.z, 0), j1 = __shfl_sync(0xffffffffu, sg.w, 0);
        const uint32_t abuf = a_iter & 1;
        ptx::mbar_wait(&a_full[abuf], (a_iter >> 1) & 1);
        for (int j = j0; j < j1; ++j) {
          const uint32_t buf = tile_iter & 1;
          ptx::mbar_wait(&t_empty[buf], ((tile_iter >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kTileN2;
          for (int kb = 0; kb < NUM_KB; ++kb) {
            ptx::mbar_wait(&full[stage], phase);
            ptx::tc_fence_after();
            const uint64_t dq0 = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem_q + stage * L::kQStage), kRowBytes);
            const uint64_t da0 = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem_a + abuf * L::kABytes + kb * L::kABlock), kRowBytes);
            const bool leader = ptx::elect_one();
#pragma unroll
            for (int k = 0; k < kKSteps; ++k)
              if (leader) ptx::umma_f16_pair(d_tmem, da0 + 2 * k, dq0 + 2 * k, kIdesc, (kb | k) != 0 ? 1u : 0u);
            if (leader) {
              ptx::umma_commit_pair(&empty[stage], 3);                       // both CTAs' stage slots reusable
              if (kb == NUM_KB - 1) ptx::umma_commit_pair(&t_full[buf], 3);  // accumulator complete in both CTAs
            }
            __syncwarp();
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
          ++tile_iter;
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(&a_empty[abuf], 3);
        __syncwarp();
        ++a_iter;
      }
    }
  } else {
    // ============================ epilogue (8 warps per CTA: 4 lane quarters x 2 column halves) ============================
    const int ew = (warp - 2) & 7;
    const int half = ew >> 2;              // columns [128 * half, +128) of every 256-column tile
    const int quarter = warp & 3;          // TMEM lanes 32*quarter .. +31 are accessible to this warp
    const int row_in_block = (int)rank * kTileM + quarter * 32 + lane;
    const uint32_t t_empty_leader0 = ptx::mapa_shared(ptx::smem_u32(&t_empty[0]), 0), t_empty_leader1 = ptx::mapa_shared(ptx::smem_u32(&t_empty[1]), 0);
    uint32_t tile_iter = 0;
    for (int si = seg_lo; si < seg_hi; ++si) {
      const int4 sg = __ldg(reinterpret_cast<const int4*>(args.segs) + si);
      const int b = sg.x & 0xFFFFFF, s = sg.x >> 24, rb = sg.y, j0 = sg.z, j1 = sg.w;
      const PairMeta pm = args.meta[b];
      const int row = rb * kCtaRows + row_in_block;
      const bool row_ok = row < pm.n_a;
      const size_t slot = (((size_t)b * args.splits + s) * 2 + half) * args.npad_a + row;
      uint32_t* my_list = args.cand_chunk + slot * kCandCap;
      float m_run = -INFINITY;
      float thr = row_ok ? -INFINITY : INFINITY;  // rows beyond n_a never record anything
      int cnt = 0;
      for (int j = j0; j < j1; ++j) {
        const uint32_t buf = tile_iter & 1;
        ptx::mbar_wait(&t_full[buf], (tile_iter >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kTileN2 + half * (kTileN2 / 2);
        const int col_base = j * kTileN2 + half * (kTileN2 / 2);
        const bool ragged = col_base + kTileN2 / 2 > pm.n_q;
        uint32_t va[32], vb[32];
        auto scan = [&](uint32_t (&v)[32], int g) {
          if (ragged) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col_base + g * 32 + i >= pm.n_q) v[i] = 0xff800000u;  // -inf
          }
          float cm[32 / kChunk];
#pragma unroll
          for (int c = 0; c < 32 / kChunk; ++c) cm[c] = max8(v + c * kChunk);
          float gm = cm[0];
#pragma unroll
          for (int c = 1; c < 32 / kChunk; ++c) gm = fmaxf(gm, cm[c]);
          if (gm >= thr) {
#pragma unroll
            for (int c = 0; c < 32 / kChunk; ++c) {
              if (cm[c] >= thr) {
                if (cm[c] > m_run + args.ambiguity) cnt = 0;  // every earlier candidate is now out of range
                if (cnt < kCandCap) {
                  const float lim = cm[c] - args.ambiguity;
                  uint32_t bits = 0;
#pragma unroll
                  for (int i = 0; i < kChunk; ++i) bits |= (__uint_as_float(v[c * kChunk + i]) >= lim ? 1u : 0u) << i;
                  my_list[cnt] = static_cast<uint32_t>((col_base + g * 32) / kChunk + c) | (bits << 24);
                }
                ++cnt;
                m_run = fmaxf(m_run, cm[c]);
                thr = m_run - args.ambiguity;
              }
            }
          }
        };
        ptx::tmem_ld_32x32b_x32(taddr, va);
#pragma unroll
        for (int g = 0; g < kTileN2 / 64; g += 2) {
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(taddr + (g + 1) * 32, vb);
          scan(va, g);
          ptx::tmem_ld_wait();
          if (g + 2 < kTileN2 / 64) ptx::tmem_ld_32x32b_x32(taddr + (g + 2) * 32, va);
          scan(vb, g + 1);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(buf ? t_empty_leader1 : t_empty_leader0);
        ++tile_iter;
      }
      if (row_ok) args.cand_m[slot] = m_run, args.cand_cnt[slot] = cnt;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();      // neither CTA leaves (or frees tensor memory) while the pair's last MMAs / arrivals may touch it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// refine: exact fp32 score of every column of every listed chunk; one warp per anchor row
// ------------------------------------------------------------------------------------------------
struct RefineArgs {
  const float* rows32_a;
  const float* rows32_q;
  const PairMeta* meta;
  const float* cand_m;
  const int32_t* cand_cnt;
  const uint32_t* cand_chunk;
  const uint8_t* nseg;         // [B][npad_a / kCtaRows] candidate lists (segments) per anchor row block
  int B, npad_a, npad_q, D4, splits, cap_a;
  int halves;                  // candidate lists per (row, segment): 1 (single-CTA kernel) or 2 (pair kernel: one per column half)
  int item_base;               // first pair of this launch * npad_a (overflow rows are packed with batch-wide pair indices)
  float ambiguity;
  int32_t* out_idx;
  float* out_dist;
  int32_t* overflow_rows;      // packed b * npad_a + r
  int32_t* overflow_count;     // device counter
  unsigned long long* stats;   // [0] rows refined, [1] chunks scored, [2] overflow rows
  unsigned long long* hist;    // optional (oryon_match_set_hist): [0..17] rows by candidate chunks (17 = overflow), [18..25] rows by
                               // candidate columns re-scored (1, 2, 3-4, 5-8, 9-16, 17-32, 33-64, 65+)
};

__device__ __forceinline__ void record_hist(unsigned long long* hist, int chunks, int cols, bool overflow) {
  atomicAdd(hist + (overflow ? 17 : min(chunks, 16)), 1ull);
  if (!overflow) atomicAdd(hist + 18 + (cols <= 1 ? 0 : cols <= 2 ? 1 : min(32 - __clz(cols - 1), 7)), 1ull);
}

__global__ void __launch_bounds__(256) refine_rows_kernel(RefineArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = a.D4 >> 2;
  unsigned long long n_rows = 0, n_chunks = 0, n_over = 0;
  for (int item = wid; item < a.B * a.npad_a; item += nwarps) {
    const int b = item / a.npad_a, r = item - b * a.npad_a;
    const PairMeta pm = a.meta[b];
    if (r >= pm.n_a) {
      continue;
    }
    const size_t o = (size_t)b * a.cap_a + r;
    if (pm.n_q <= 0) {
      if (lane == 0) a.out_idx[o] = -1, a.out_dist[o] = INFINITY;
      continue;
    }
    const int lists = a.nseg[b * (a.npad_a / kCtaRows) + r / kCtaRows] * a.halves;
    float m_all = -INFINITY;
    bool overflow = false;
    for (int s = 0; s < lists; ++s) {
      const size_t slot = ((size_t)b * a.splits + s) * a.npad_a + r;
      m_all = fmaxf(m_all, a.cand_m[slot]);
      overflow |= a.cand_cnt[slot] > kCandCap;
    }
    if (overflow) {
      if (lane == 0) a.overflow_rows[atomicAdd(a.overflow_count, 1)] = a.item_base + item;
      if (lane == 0 && a.hist) record_hist(a.hist, 0, 0, true);
      ++n_over;
      continue;
    }
    ++n_rows;
    int h_chunks = 0, h_cols = 0;
    const float4* arow = reinterpret_cast<const float4*>(a.rows32_a + ((size_t)b * a.npad_a + r) * a.D4);
    float best = -INFINITY;
    int best_j = 0x7fffffff;
    for (int s = 0; s < lists; ++s) {
      const size_t slot = ((size_t)b * a.splits + s) * a.npad_a + r;
      if (a.cand_m[slot] < m_all - a.ambiguity) continue;  // nothing in this split can win
      const int cnt = a.cand_cnt[slot];
      for (int e = 0; e < cnt; ++e) {
        const uint32_t ent = a.cand_chunk[slot * kCandCap + e];
        const int col0 = (int)(ent & 0xFFFFFFu) * kChunk;
        uint32_t bits = ent >> 24;
        while (bits) {   // usually a single column: the whole warp scores it (one float4 per lane at D = 128)
          const int col = col0 + __ffs(bits) - 1;
          bits &= bits - 1;
          if (col >= pm.n_q) continue;
          const float4* qrow = reinterpret_cast<const float4*>(a.rows32_q + ((size_t)b * a.npad_q + col) * a.D4);
          float acc = 0.f;
          for (int i = lane; i < nvec; i += 32) {
            const float4 x = __ldg(arow + i), y = __ldg(qrow + i);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
            acc = fmaf(x.z, y.z, acc);
            acc = fmaf(x.w, y.w, acc);
          }
#pragma unroll
          for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          if (acc > best || (acc == best && col < best_j)) best = acc, best_j = col;
          ++h_cols;
        }
        ++n_chunks, ++h_chunks;
      }
    }
    if (lane == 0) {
      a.out_idx[o] = best_j;
      a.out_dist[o] = 0.5f * (-1.f * best + 1.f);
      if (a.hist) record_hist(a.hist, h_chunks, h_cols, false);
    }
  }
  if (lane == 0 && a.stats) {
    if (n_rows) atomicAdd(a.stats + 0, n_rows);
    if (n_chunks) atomicAdd(a.stats + 1, n_chunks);
    if (n_over) atomicAdd(a.stats + 2, n_over);
  }
}

// Second version: eight lanes per anchor row, four rows per warp.  The work per row is a chain of dependent loads (list
// length -> list entry -> query row) followed by a 128-element dot product; with a whole warp per row only one such chain
// per warp is in flight and the dot product is a single float4 per lane.  Eight lanes per row keep four chains in flight per
// warp and still read 128 contiguous bytes per row and instruction.  Groups diverge only when their list lengths differ.
// Same arithmetic as refine_rows_kernel up to the summation order of the dot product (8 partial sums of 16 instead of 32 of 4).
__global__ void __launch_bounds__(256) refine_rows4_kernel(RefineArgs a) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  const unsigned gmask = 0xffu << (grp * 8);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nvec = a.D4 >> 2;
  const int total = a.B * a.npad_a;
  unsigned long long n_rows = 0, n_chunks = 0, n_over = 0;
  for (int item0 = wid * 4; item0 < total; item0 += nwarps * 4) {
    const int item = item0 + grp;
    if (item >= total) continue;
    const int b = item / a.npad_a, r = item - b * a.npad_a;
    const PairMeta pm = a.meta[b];
    if (r >= pm.n_a) continue;
    const size_t o = (size_t)b * a.cap_a + r;
    if (pm.n_q <= 0) {
      if (sub == 0) a.out_idx[o] = -1, a.out_dist[o] = INFINITY;
      continue;
    }
    const int lists = a.nseg[b * (a.npad_a / kCtaRows) + r / kCtaRows] * a.halves;
    float m_all = -INFINITY;
    bool overflow = false;
    for (int s = 0; s < lists; ++s) {
      const size_t slot = ((size_t)b * a.splits + s) * a.npad_a + r;
      m_all = fmaxf(m_all, a.cand_m[slot]);
      overflow |= a.cand_cnt[slot] > kCandCap;
    }
    if (overflow) {
      if (sub == 0) a.overflow_rows[atomicAdd(a.overflow_count, 1)] = a.item_base + item, ++n_over;
      if (sub == 0 && a.hist) record_hist(a.hist, 0, 0, true);
      continue;
    }
    if (sub == 0) ++n_rows;
    int h_chunks = 0, h_cols = 0;
    const float4* arow = reinterpret_cast<const float4*>(a.rows32_a + ((size_t)b * a.npad_a + r) * a.D4);
    float best = -INFINITY;
    int best_j = 0x7fffffff;
    for (int s = 0; s < lists; ++s) {
      const size_t slot = ((size_t)b * a.splits + s) * a.npad_a + r;
      if (a.cand_m[slot] < m_all - a.ambiguity) continue;  // nothing in this list can win
      const int cnt = a.cand_cnt[slot];
      for (int e = 0; e < cnt; ++e) {
        const uint32_t ent = a.cand_chunk[slot * kCandCap + e];
        const int col0 = (int)(ent & 0xFFFFFFu) * kChunk;
        uint32_t bits = ent >> 24;
        while (bits) {
          const int col = col0 + __ffs(bits) - 1;
          bits &= bits - 1;
          if (col >= pm.n_q) continue;
          const float4* qrow = reinterpret_cast<const float4*>(a.rows32_q + ((size_t)b * a.npad_q + col) * a.D4);
          float acc = 0.f;
          for (int i = sub; i < nvec; i += 8) {
            const float4 x = __ldg(arow + i), y = __ldg(qrow + i);
            acc = fmaf(x.x, y.x, acc);
            acc = fmaf(x.y, y.y, acc);
            acc = fmaf(x.z, y.z, acc);
            acc = fmaf(x.w, y.w, acc);
          }
#pragma unroll
          for (int off = 4; off >= 1; off >>= 1) acc += __shfl_xor_sync(gmask, acc, off);
          if (acc > best || (acc == best && col < best_j)) best = acc, best_j = col;
          ++h_cols;
        }
        if (sub == 0) ++n_chunks;
        ++h_chunks;
      }
    }
    if (sub == 0) {
      a.out_idx[o] = best_j;
      a.out_dist[o] = 0.5f * (-1.f * best + 1.f);
      if (a.hist) record_hist(a.hist, h_chunks, h_cols, false);
    }
  }
  if (sub == 0 && a.stats) {
    if (n_rows) atomicAdd(a.stats + 0, n_rows);
    if (n_chunks) atomicAdd(a.stats + 1, n_chunks);
    if (n_over) atomicAdd(a.stats + 2, n_over);
  }
}

__global__ void fill_invalid_kernel(int32_t* out_idx, float* out_dist, const PairMeta* meta, int B, int cap_a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * cap_a) return;
  const int b = i / cap_a, r = i - b * cap_a;
  if (r >= meta[b].n_a) out_idx[i] = -1, out_dist[i] = INFINITY;
}

// ------------------------------------------------------------------------------------------------
// mask -> ordered pixel-id list (torch.nonzero order)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mask_to_roi_kernel(const int32_t* mask, int HW, int value, int32_t* roi, int32_t* n_out) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* m = mask + (size_t)b * HW;
  int32_t* out = roi + (size_t)b * HW;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < HW; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    const bool hit = i < HW && m[i] == value;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (hit) out[off + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_tot[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) n_out[b] = base;
}

// ------------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------------
static int make_rows_tensor_map(oryon_handle* h, CUtensorMap* tm, const void* base, int rows, int dpad, int kb_elems, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)dpad * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kb_elems, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = kb_elems == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const CUresult r = h->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d dpad=%d)", (int)r, rows, dpad);
    return ORYON_ERR_CUDA;
  }
  return ORYON_OK;
}

template <int KB_ELEMS, int NUM_KB, int STAGES>
static int launch_tc(oryon_handle* h, const CUtensorMap& tma, const CUtensorMap& tmq, const TcArgs& args, int grid, cudaStream_t st) {
  using L = TcSmem<KB_ELEMS, NUM_KB, STAGES>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  auto kern = match_tc_kernel<KB_ELEMS, NUM_KB, STAGES>;
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  h->span_begin(KID_MATCH_TC, st);
  kern<<<grid, kTcThreads, L::kTotal, st>>>(tma, tmq, args);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

template <int KB_ELEMS, int NUM_KB, int STAGES>
static int launch_tc2(oryon_handle* h, const CUtensorMap& tma, const CUtensorMap& tmq, const TcArgs& args, int pairs, cudaStream_t st) {
  using L = Tc2Smem<KB_ELEMS, NUM_KB, STAGES>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  auto kern = match_tc2_kernel<KB_ELEMS, NUM_KB, STAGES>;
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  h->span_begin(KID_MATCH_TC, st);
  kern<<<2 * pairs, kTcThreads, L::kTotal, st>>>(tma, tmq, args);   // __cluster_dims__(2,1,1): CTAs 2p, 2p+1 form pair p
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct TcPlan {
  std::vector<Seg> segs;
  std::vector<int32_t> begin;   // [grid + 1] first segment of every CTA
  std::vector<uint8_t> nseg;    // [B][rb_per_pair] segments (candidate lists) of a row block
  int grid = 0, splits = 0;
};

// Distributes the (pair, row block, query tile) units of a batch over the CTAs of the persistent kernel (see Seg).
//   kPlanHybrid      whole row blocks round robin for as many full waves as there are (CTAs that run at the same time then sweep
//                    neighbouring row blocks of the same pair, so the query tiles they stream are shared through L2), and the
//                    remaining row blocks cut into per-CTA quotas of query tiles that level the finishing times (water filling
//                    over the bulk loads, which differ when the pairs are ragged).
//   kPlanContiguous  every CTA gets one contiguous range of units.  Equally balanced, but at config 2 the 148 CTAs then sweep 148
//                    different row blocks spread over all 32 pairs: the query rows in flight (157 MB) no longer fit L2 and the
//                    kernel becomes HBM-bound (measured 2.69 ms against 2.44 ms for whole-task round robin).  Kept for A/B.
//   kPlanWholeTasks  round robin of whole row blocks whenever there are at least as many as CTAs (no cutting: the tail wave is
//                    partly idle); the first version's behaviour, kept for A/B.
// A quota is either 0 or >= tiles_max / (kMaxSplits - 1), so no row block is shared by more than kMaxSplits CTAs.
enum PlanKind { kPlanHybrid = 0, kPlanContiguous = 1, kPlanWholeTasks = 2 };

static void build_tc_plan(const std::vector<PairMeta>& meta, int rb_per_pair, int n_workers, int kind, TcPlan* plan, int tile_n = kTileN) {
  struct Task {
    int b, rb, tiles;
  };
  const int B = (int)meta.size();
  plan->nseg.assign((size_t)B * rb_per_pair, 0);
  std::vector<Task> tasks;
  long long units = 0;
  int tiles_max = 0;
  for (int b = 0; b < B; ++b) {
    const PairMeta& pm = meta[b];
    if (pm.n_a <= 0 || pm.n_q <= 0) continue;
    const int tiles = (pm.n_q + tile_n - 1) / tile_n, rbs = (pm.n_a + kCtaRows - 1) / kCtaRows;
    for (int rb = 0; rb < rbs; ++rb) tasks.push_back(Task{b, rb, tiles});
    units += (long long)rbs * tiles;
    tiles_max = std::max(tiles_max, tiles);
  }
  if (units == 0) {
    plan->begin.assign(1, 0);
    return;
  }
  const int G = n_workers;   // CTAs of the single-CTA kernel, CTA pairs of the pair kernel
  const long long min_share = (tiles_max + kMaxSplits - 2) / (kMaxSplits - 1);
  size_t n_bulk = tasks.size() / G * G;
  if (kind == kPlanContiguous) n_bulk = 0;
  if (kind == kPlanWholeTasks && tasks.size() >= (size_t)G) n_bulk = tasks.size();
  std::vector<std::vector<Seg>> lists(G);
  std::vector<long long> load(G, 0), quota(G, 0);
  for (size_t t = 0; t < n_bulk; ++t) {
    const Task& k = tasks[t];
    lists[t % G].push_back(Seg{k.b, k.rb, 0, k.tiles});
    load[t % G] += k.tiles;
    plan->nseg[(size_t)k.b * rb_per_pair + k.rb] = 1;
    plan->splits = std::max(plan->splits, 1);
  }
  long long rem = 0;
  for (size_t t = n_bulk; t < tasks.size(); ++t) rem += tasks[t].tiles;
  if (rem > 0) {
    // water filling over the k least loaded CTAs, with the largest k that leaves every participant a quota >= min_share
    std::vector<int> order(G);
    for (int c = 0; c < G; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return load[x] < load[y]; });
    std::vector<long long> pre(G + 1, 0);
    for (int i = 0; i < G; ++i) pre[i + 1] = pre[i] + load[order[i]];
    int k = G;
    long long level = 0;
    for (; k >= 1; --k) {
      level = (rem + pre[k]) / k;
      if (k == 1 || level - load[order[k - 1]] >= min_share) break;
    }
    long long left = rem - (level * k - pre[k]);   // < k
    std::vector<int> part(order.begin(), order.begin() + k);
    std::sort(part.begin(), part.end());
    for (int c : part) quota[c] = level - load[c] + (left > 0 ? 1 : 0), left -= left > 0 ? 1 : 0;
    int c = 0;
    long long room = quota[0];
    for (size_t t = n_bulk; t < tasks.size(); ++t) {
      const Task& tk = tasks[t];
      int j = 0, slot = 0;
      while (j < tk.tiles) {
        while (room == 0 && c + 1 < G) room = quota[++c];
        const int take = (int)std::min<long long>(tk.tiles - j, room > 0 ? room : tk.tiles - j);
        lists[c].push_back(Seg{tk.b | (slot << 24), tk.rb, j, j + take});
        j += take, room -= std::min<long long>(room, take), ++slot;
      }
      plan->nseg[(size_t)tk.b * rb_per_pair + tk.rb] = (uint8_t)slot;
      plan->splits = std::max(plan->splits, slot);
    }
  }
  plan->begin.assign(1, 0);
  for (int c = 0; c < G; ++c) {
    if (lists[c].empty()) continue;   // fewer units than CTAs x min_share: idle CTAs are not launched
    plan->segs.insert(plan->segs.end(), lists[c].begin(), lists[c].end());
    plan->begin.push_back((int32_t)plan->segs.size());
  }
  plan->grid = (int)plan->begin.size() - 1;
}

int run_match(oryon_handle* h, const float* feat_a, const float* feat_q, int B, int D, int HW_a, int HW_q, const int32_t* roi_a,
              const int32_t* roi_q, const int32_t* n_a, const int32_t* n_q, int cap_a, int cap_q, int mode, int32_t* out_idx,
              float* out_dist, cudaStream_t st) {
  ORYON_REQUIRE(h && feat_a && feat_q && out_idx && out_dist, "oryon_match_nn: null argument");
  ORYON_REQUIRE(B > 0 && D > 0 && HW_a > 0 && HW_q > 0, "oryon_match_nn: B, D, HW must be positive");
  ORYON_REQUIRE(D <= 256, "oryon_match_nn: D=%d not supported (max 256)", D);
  ORYON_REQUIRE(B < (1 << 24), "oryon_match_nn: B=%d pairs per call not supported", B);
  ORYON_REQUIRE((roi_a == nullptr) == (n_a == nullptr) && (roi_q == nullptr) == (n_q == nullptr),
                "oryon_match_nn: roi_x and n_x must both be given or both be NULL");
  ORYON_REQUIRE(mode == ORYON_MATCH_TC_REFINED || mode == ORYON_MATCH_EXACT_FP32, "oryon_match_nn: unknown mode %d", mode);
  if (!roi_a) ORYON_REQUIRE(cap_a >= HW_a, "oryon_match_nn: dense anchor list needs cap_a >= HW_a");
  if (!roi_q) cap_q = HW_q;

  std::vector<PairMeta> meta(B);
  int max_a = 0, max_q = 0;
  for (int b = 0; b < B; ++b) {
    meta[b].n_a = n_a ? n_a[b] : HW_a;
    meta[b].n_q = n_q ? n_q[b] : HW_q;
    ORYON_REQUIRE(meta[b].n_a >= 0 && meta[b].n_a <= cap_a && meta[b].n_q >= 0 && meta[b].n_q <= cap_q,
                  "oryon_match_nn: pair %d list lengths (%d, %d) outside [0, cap]", b, meta[b].n_a, meta[b].n_q);
    max_a = std::max(max_a, meta[b].n_a);
    max_q = std::max(max_q, meta[b].n_q);
  }
  h->last_launches = 0;
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  int rc;
  if ((rc = h->pair_meta.reserve(sizeof(PairMeta) * B, st))) return rc;
  if ((rc = h->counters.reserve(512, st))) return rc;
  ORYON_CUDA_CHECK(cudaMemcpyAsync(h->pair_meta.ptr, meta.data(), sizeof(PairMeta) * B, cudaMemcpyHostToDevice, st));
  ORYON_CUDA_CHECK(cudaMemsetAsync(h->counters.ptr, 0, 512, st));
  const PairMeta* d_meta = h->pair_meta.as<PairMeta>();
  // counters: [0] overflow count (int32), bytes 64.. : stats (3 x u64), bytes 128.. : optional list-length histogram (26 x u64)
  int32_t* d_overflow_count = h->counters.as<int32_t>();
  unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(h->counters.as<char>() + 64);

  {
    const int blocks = (B * cap_a + 255) / 256;
    fill_invalid_kernel<<<blocks, 256, 0, st>>>(out_idx, out_dist, d_meta, B, cap_a);
    ORYON_CUDA_CHECK(cudaGetLastError());
    ++h->last_launches;
  }
  if (max_a == 0) return ORYON_OK;

  const int D4 = round_up(D, 4);
  const bool sw64 = D <= 32;
  const int kb_elems = sw64 ? 32 : 64;
  const int Dpad = round_up(D, kb_elems);
  const int num_kb = Dpad / kb_elems;
  // tensor-core pass: the single-CTA kernel, or -- ORYON_MATCH_PAIR=1, read per call so that one process can run both -- CTA
  // pairs (tcgen05 cta_group::2, 256-column query tiles).  Measured on B200 at config 2 (profiles/r02_match_pair_vs_1cta.md):
  // pair 2.74 ms / 59 % tensor-pipe active, single CTA 2.53 ms / 68.5 %.  Warp-state sampling shows why neither is limited by
  // operand bandwidth: the MMA warp spends ~3/4 of its time waiting for a drained accumulator (t_empty) while the epilogue warps
  // wait a third of theirs for a filled one -- at D = 128 a 128 x 256 fp32 accumulator (128 KB of tensor memory) must be read
  // out every 1 024 tensor cycles, and that read-out plus the two barrier round trips takes ~2 000 cycles per buffer; with only
  // 512 tensor-memory columns there is no room for a third buffer.  The pair kernel's buffers are twice as wide per barrier
  // round trip, hence slower.  At D = 256 (config 5) the same single-CTA kernel is MMA-paced (0.98 of the sustained peak).
  const char* pair_env = std::getenv("ORYON_MATCH_PAIR");
  const bool use_pair = pair_env && pair_env[0] == '1' && h->sm_count >= 2;
  const int tile_n = use_pair ? kTileN2 : kTileN;
  const int npad_a = round_up(max_a, kCtaRows), npad_q = round_up(std::max(max_q, 1), tile_n);

  if ((rc = h->rows16_a.reserve((size_t)B * npad_a * Dpad * 2, st))) return rc;
  if ((rc = h->rows16_q.reserve((size_t)B * npad_q * Dpad * 2, st))) return rc;
  if ((rc = h->rows32_a.reserve((size_t)B * npad_a * D4 * 4, st))) return rc;
  if ((rc = h->rows32_q.reserve((size_t)B * npad_q * D4 * 4, st))) return rc;

  // ---- prep of pairs [b0, b0 + nb) on `stream` ----
  auto launch_prep = [&](int b0, int nb, cudaStream_t stream) -> int {
    PrepArgs pa;
    pa.feat[0] = feat_a + (size_t)b0 * D * HW_a, pa.feat[1] = feat_q + (size_t)b0 * D * HW_q;
    pa.roi[0] = roi_a ? roi_a + (size_t)b0 * cap_a : nullptr, pa.roi[1] = roi_q ? roi_q + (size_t)b0 * cap_q : nullptr;
    pa.hw[0] = HW_a, pa.hw[1] = HW_q;
    pa.cap[0] = cap_a, pa.cap[1] = cap_q;
    pa.npad[0] = npad_a, pa.npad[1] = npad_q;
    pa.rows16[0] = h->rows16_a.as<__half>() + (size_t)b0 * npad_a * Dpad, pa.rows16[1] = h->rows16_q.as<__half>() + (size_t)b0 * npad_q * Dpad;
    pa.rows32[0] = h->rows32_a.as<float>() + (size_t)b0 * npad_a * D4, pa.rows32[1] = h->rows32_q.as<float>() + (size_t)b0 * npad_q * D4;
    pa.meta = d_meta + b0;
    pa.D = D, pa.D4 = D4, pa.Dpad = Dpad;
    h->span_begin(KID_PREP, stream);
    static const bool prep_v1 = std::getenv("ORYON_PREP_V1") != nullptr;   // A/B switch: the first dense kernel
    const bool dense = !roi_a && !roi_q && (HW_a % 2) == 0 && (HW_q % 2) == 0;
    if (dense && !prep_v1 && (D == 32 || D == 64 || D == 128 || D == 256) && Dpad == D) {
      const dim3 grid((std::max(max_a, max_q) + kPrepPix - 1) / kPrepPix, nb, 2);
      const size_t smem = (size_t)kPrepPix * (D + 4) * sizeof(float);
      auto launch = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, 256, smem, stream>>>(pa);
        return e;
      };
      cudaError_t e = D == 32 ? launch(prep_dense2_kernel<1>) : D == 64 ? launch(prep_dense2_kernel<2>)
                    : D == 128 ? launch(prep_dense2_kernel<4>) : launch(prep_dense2_kernel<8>);
      ORYON_CUDA_CHECK(e);
    } else if (dense) {
      const dim3 grid((std::max(max_a, max_q) + kPrepPix - 1) / kPrepPix, nb, 2);
      const size_t smem = (size_t)D * (kPrepPix + 1) * sizeof(float);
      ORYON_CUDA_CHECK(cudaFuncSetAttribute(prep_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      prep_dense_kernel<<<grid, 256, smem, stream>>>(pa);
    } else {
      const dim3 grid((std::max(max_a, max_q) + 31) / 32, nb, 2);
      const size_t smem = (size_t)D * 33 * sizeof(float);
      ORYON_CUDA_CHECK(cudaFuncSetAttribute(prep_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      prep_rows_kernel<<<grid, 256, smem, stream>>>(pa);
    }
    h->span_end(stream);
    ORYON_CUDA_CHECK(cudaGetLastError());
    ++h->last_launches;
    return ORYON_OK;
  };

  ExactArgs ea;
  ea.rows32_a = h->rows32_a.as<float>(), ea.rows32_q = h->rows32_q.as<float>();
  ea.meta = d_meta;
  ea.npad_a = npad_a, ea.npad_q = npad_q, ea.D4 = D4, ea.cap_a = cap_a;
  ea.out_idx = out_idx, ea.out_dist = out_dist;

  if (mode == ORYON_MATCH_EXACT_FP32 || max_q == 0) {
    if ((rc = launch_prep(0, B, st))) return rc;
    ea.row_list = nullptr, ea.n_items_dev = nullptr, ea.n_items_host = B * npad_a;
    const int blocks = std::min((B * npad_a + 63) / 64, h->sm_count * 8);
    h->span_begin(KID_EXACT, st);
    exact_rows_kernel<<<blocks, 256, 0, st>>>(ea);
    h->span_end(st);
    ORYON_CUDA_CHECK(cudaGetLastError());
    ++h->last_launches;
    return ORYON_OK;
  }

  // ---- tensor-core pass ----
  // (Tried and measured, profiles/r01_match_pipelining_experiment_run35.json: cutting the batch into chunks of pairs and running
  // the prep of the next chunk / the refine of the previous one on side streams next to the persistent tensor-core CTAs does not
  // shorten the step -- match_tc slows down by what the co-runners hide.  The three launchers below still take a pair range.)
  static const int plan_kind = [] {
    const char* e = std::getenv("ORYON_MATCH_PLAN");   // A/B switch: "contiguous" | "whole" (default: hybrid)
    return !e ? kPlanHybrid : !strcmp(e, "contiguous") ? kPlanContiguous : !strcmp(e, "whole") ? kPlanWholeTasks : kPlanHybrid;
  }();
  const int rb_per_pair = npad_a / kCtaRows;
  const int halves = use_pair ? 2 : kEpiSets;   // candidate lists per (row, segment): one per column half (pair kernel) / warp set
  struct Chunk {   // a range of pairs with its work decomposition and its slice of the candidate / plan buffers
    int b0, nb, splits;
    TcPlan plan;
    size_t slot_base, slots, off_begin, off_segs, off_nseg;
  };
  Chunk all;
  all.b0 = 0, all.nb = B;
  build_tc_plan(meta, rb_per_pair, use_pair ? h->sm_count / 2 : h->sm_count, plan_kind, &all.plan, tile_n);
  all.splits = std::max(all.plan.splits, 1);
  all.slot_base = 0, all.slots = (size_t)B * all.splits * halves * npad_a;
  all.off_begin = 0;
  all.off_segs = (sizeof(int32_t) * all.plan.begin.size() + 15) & ~size_t(15);
  all.off_nseg = all.off_segs + sizeof(Seg) * all.plan.segs.size();
  const size_t slots_total = all.slots;
  if ((rc = h->cand.reserve(slots_total * (4 + 4 + 4 * kCandCap), st))) return rc;
  if ((rc = h->overflow_rows.reserve((size_t)B * npad_a * 4, st))) return rc;
  // one upload: [first segment of every CTA | segments | candidate lists per row block]
  std::vector<uint8_t> blob(all.off_nseg + all.plan.nseg.size());
  memcpy(blob.data() + all.off_begin, all.plan.begin.data(), sizeof(int32_t) * all.plan.begin.size());
  if (!all.plan.segs.empty()) memcpy(blob.data() + all.off_segs, all.plan.segs.data(), sizeof(Seg) * all.plan.segs.size());
  memcpy(blob.data() + all.off_nseg, all.plan.nseg.data(), all.plan.nseg.size());
  if ((rc = h->match_plan.reserve(blob.size(), st))) return rc;
  ORYON_CUDA_CHECK(cudaMemcpyAsync(h->match_plan.ptr, blob.data(), blob.size(), cudaMemcpyHostToDevice, st));
  float* cand_m = h->cand.as<float>();
  int32_t* cand_cnt = reinterpret_cast<int32_t*>(cand_m + slots_total);
  uint32_t* cand_chunk = reinterpret_cast<uint32_t*>(cand_cnt + slots_total);

  auto launch_match = [&](const Chunk& k, cudaStream_t stream) -> int {
    if (k.plan.grid == 0) return ORYON_OK;   // no (row block, query tile) unit: the refine pass writes (-1, inf)
    TcArgs ta;
    ta.meta = d_meta + k.b0;
    ta.seg_begin = reinterpret_cast<const int32_t*>(h->match_plan.as<char>() + k.off_begin);
    ta.segs = reinterpret_cast<const Seg*>(h->match_plan.as<char>() + k.off_segs);
    ta.B = k.nb, ta.npad_a = npad_a, ta.npad_q = npad_q, ta.splits = k.splits;
    ta.cand_m = cand_m + k.slot_base, ta.cand_cnt = cand_cnt + k.slot_base, ta.cand_chunk = cand_chunk + k.slot_base * kCandCap;
    ta.ambiguity = kAmbiguity;
    const char* dbg_env = std::getenv("ORYON_MATCH_DEBUG_MODE");
    ta.dbg_mode = dbg_env ? atoi(dbg_env) : 0;
    CUtensorMap tma, tmq;
    int r;
    if ((r = make_rows_tensor_map(h, &tma, h->rows16_a.as<__half>() + (size_t)k.b0 * npad_a * Dpad, k.nb * npad_a, Dpad, kb_elems, kTileM))) return r;
    if ((r = make_rows_tensor_map(h, &tmq, h->rows16_q.as<__half>() + (size_t)k.b0 * npad_q * Dpad, k.nb * npad_q, Dpad, kb_elems, kTileM))) return r;   // 128 query rows per TMA box: a whole tile (single CTA) / this CTA's half (pair)
    if (use_pair) {
      if (sw64) {
        r = launch_tc2<32, 1, 8>(h, tma, tmq, ta, k.plan.grid, stream);
      } else {
        switch (num_kb) {
          case 1: r = launch_tc2<64, 1, 8>(h, tma, tmq, ta, k.plan.grid, stream); break;
          case 2: r = launch_tc2<64, 2, 8>(h, tma, tmq, ta, k.plan.grid, stream); break;
          case 3: r = launch_tc2<64, 3, 6>(h, tma, tmq, ta, k.plan.grid, stream); break;
          default: r = launch_tc2<64, 4, 5>(h, tma, tmq, ta, k.plan.grid, stream); break;
        }
      }
    } else if (sw64) {
      r = launch_tc<32, 1, 8>(h, tma, tmq, ta, k.plan.grid, stream);
    } else {
      switch (num_kb) {
        case 1: r = launch_tc<64, 1, 8>(h, tma, tmq, ta, k.plan.grid, stream); break;
        case 2: r = launch_tc<64, 2, 8>(h, tma, tmq, ta, k.plan.grid, stream); break;
        case 3: r = launch_tc<64, 3, 6>(h, tma, tmq, ta, k.plan.grid, stream); break;
        default: r = launch_tc<64, 4, 5>(h, tma, tmq, ta, k.plan.grid, stream); break;
      }
    }
    if (r) return r;
    ++h->last_launches;
    return ORYON_OK;
  };

  auto launch_refine = [&](const Chunk& k, cudaStream_t stream) -> int {
    RefineArgs ra;
    ra.rows32_a = h->rows32_a.as<float>() + (size_t)k.b0 * npad_a * D4, ra.rows32_q = h->rows32_q.as<float>() + (size_t)k.b0 * npad_q * D4;
    ra.meta = d_meta + k.b0;
    ra.cand_m = cand_m + k.slot_base, ra.cand_cnt = cand_cnt + k.slot_base, ra.cand_chunk = cand_chunk + k.slot_base * kCandCap;
    ra.nseg = reinterpret_cast<const uint8_t*>(h->match_plan.as<char>() + k.off_nseg);
    ra.B = k.nb, ra.npad_a = npad_a, ra.npad_q = npad_q, ra.D4 = D4, ra.splits = k.splits * halves, ra.cap_a = cap_a;
    ra.halves = halves;
    ra.item_base = k.b0 * npad_a;
    ra.ambiguity = kAmbiguity;
    ra.out_idx = out_idx + (size_t)k.b0 * cap_a, ra.out_dist = out_dist + (size_t)k.b0 * cap_a;
    ra.overflow_rows = h->overflow_rows.as<int32_t>();
    ra.overflow_count = d_overflow_count;
    ra.stats = d_stats;
    ra.hist = h->match_hist ? reinterpret_cast<unsigned long long*>(h->counters.as<char>() + 128) : nullptr;
    static const bool refine_v1 = std::getenv("ORYON_REFINE_V1") != nullptr;   // A/B switch: one warp per row
    const int warps_needed = refine_v1 ? k.nb * npad_a : (k.nb * npad_a + 3) / 4;
    const int blocks = std::min((warps_needed + 7) / 8, h->sm_count * 16);
    h->span_begin(KID_REFINE, stream);
    if (refine_v1) refine_rows_kernel<<<blocks, 256, 0, stream>>>(ra);
    else refine_rows4_kernel<<<blocks, 256, 0, stream>>>(ra);
    h->span_end(stream);
    ORYON_CUDA_CHECK(cudaGetLastError());
    ++h->last_launches;
    return ORYON_OK;
  };

  if ((rc = launch_prep(0, B, st))) return rc;
  if ((rc = launch_match(all, st))) return rc;
  if ((rc = launch_refine(all, st))) return rc;
  {
    ea.row_list = h->overflow_rows.as<int32_t>(), ea.n_items_dev = d_overflow_count, ea.n_items_host = 0;
    h->span_begin(KID_EXACT, st);
    exact_rows_kernel<<<h->sm_count * 2, 256, 0, st>>>(ea);
    h->span_end(st);
    ORYON_CUDA_CHECK(cudaGetLastError());
    ++h->last_launches;
  }
  return ORYON_OK;
}

int plan_debug(const int32_t* n_a, const int32_t* n_q, int B, int sm_count, int kind, int32_t* segs_out, int seg_cap,
               int32_t* begin_out, int32_t* info_out) {
  const bool pair = (kind & 0x100) != 0;   // the plan of the CTA-pair kernel: 256-column tiles over sm_count / 2 pairs
  kind &= 0xFF;
  ORYON_REQUIRE(n_a && n_q && B > 0 && B < (1 << 24) && sm_count > 0 && begin_out && info_out && (segs_out || seg_cap == 0) &&
                    kind >= kPlanHybrid && kind <= kPlanWholeTasks && (!pair || sm_count >= 2),
                "oryon_match_plan: bad argument");
  std::vector<PairMeta> meta(B);
  int max_a = 0;
  for (int b = 0; b < B; ++b) meta[b].n_a = n_a[b], meta[b].n_q = n_q[b], max_a = std::max(max_a, n_a[b]);
  TcPlan plan;
  build_tc_plan(meta, std::max(1, round_up(max_a, kCtaRows) / kCtaRows), pair ? sm_count / 2 : sm_count, kind, &plan, pair ? kTileN2 : kTileN);
  info_out[0] = plan.grid, info_out[1] = plan.splits, info_out[2] = (int32_t)plan.segs.size();
  for (int c = 0; c <= plan.grid; ++c) begin_out[c] = plan.begin[c];
  ORYON_REQUIRE((int)plan.segs.size() <= seg_cap || seg_cap == 0, "oryon_match_plan: %zu segments, capacity %d", plan.segs.size(), seg_cap);
  if (seg_cap)
    for (size_t i = 0; i < plan.segs.size(); ++i) {
      const Seg& g = plan.segs[i];
      int32_t* o = segs_out + 5 * i;
      o[0] = g.b_slot & 0xFFFFFF, o[1] = g.rb, o[2] = g.j0, o[3] = g.j1, o[4] = g.b_slot >> 24;
    }
  return ORYON_OK;
}

int run_mask_to_roi(oryon_handle* h, const int32_t* mask, int B, int HW, int value, int32_t* roi_out, int32_t* n_out, cudaStream_t st) {
  ORYON_REQUIRE(h && mask && roi_out && n_out && B > 0 && HW > 0, "oryon_mask_to_roi: bad argument");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  h->span_begin(KID_MASK_ROI, st);
  mask_to_roi_kernel<<<B, 1024, 0, st>>>(mask, HW, value, roi_out, n_out);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

int read_stats(oryon_handle* h, int64_t stats[4], cudaStream_t st) {
  ORYON_REQUIRE(h && stats, "oryon_match_last_stats: null argument");
  unsigned long long s[3] = {0, 0, 0};
  if (h->counters.ptr) {
    ORYON_CUDA_CHECK(cudaMemcpyAsync(s, h->counters.as<char>() + 64, sizeof(s), cudaMemcpyDeviceToHost, st));
    ORYON_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  stats[0] = (int64_t)s[0], stats[1] = (int64_t)s[1], stats[2] = (int64_t)s[2], stats[3] = h->last_launches;
  return ORYON_OK;
}

int read_hist(oryon_handle* h, int64_t hist[26], cudaStream_t st) {
  ORYON_REQUIRE(h && hist, "oryon_match_list_hist: null argument");
  unsigned long long v[26] = {0};
  if (h->counters.ptr && h->counters.bytes >= 128 + sizeof(v)) {
    ORYON_CUDA_CHECK(cudaMemcpyAsync(v, h->counters.as<char>() + 128, sizeof(v), cudaMemcpyDeviceToHost, st));
    ORYON_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  for (int i = 0; i < 26; ++i) hist[i] = (int64_t)v[i];
  return ORYON_OK;
}

}  // namespace match
}  // namespace oryon
