// Non-GEMM kernels of the backbone; contract in net_kernels.cuh.  All HBM-bound or latency-bound: coalesced
// along the channel / column index, one warp per row for the normalisations, grids sized by the data.
#include <algorithm>

#include <cstdlib>

#include "net_kernels.cuh"
#include "ptx_sm100.cuh"

namespace oryon {
namespace net {

namespace {

__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ float clamp_h(float x) { return fminf(fmaxf(x, -65504.f), 65504.f); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
inline int blocks_for(int64_t total, int threads, int cap) { return (int)std::min<int64_t>((total + threads - 1) / threads, cap); }

// in-register LayerNorm of one row spread over a warp: v[i] holds column lane + 32*i (i < nper)
__device__ __forceinline__ void warp_layernorm(float (&v)[32], int nper, int C, float eps, const float* gamma, const float* beta, int lane) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nper) s += v[i];
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nper) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nper) {
      const int c = lane + 32 * i;
      v[i] = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// bicubic resize + normalise + patch extraction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

struct ResizeArgs {
  const float* rgb;
  int n, in_size, out_size, patch, grid, align;
  float mean[3], inv_std[3], stdv[3];
  __half* hi;
  __half* lo;
  int ld;
};

__global__ void __launch_bounds__(256) resize_patch_kernel(ResizeArgs a) {
  const int pp = a.patch * a.patch;
  const int64_t total = (int64_t)a.n * a.grid * a.grid * a.ld;
  const float scale = a.align ? (float)(a.in_size - 1) / (float)(a.out_size - 1) : (float)a.in_size / (float)a.out_size;
  const float A = -0.75f;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int col = (int)(e % a.ld);
    const int64_t row = e / a.ld;
    float val = 0.f;
    if (col < 3 * pp) {
      const int c = col / pp, ky = (col % pp) / a.patch, kx = col % a.patch;
      const int px = (int)(row % a.grid), py = (int)((row / a.grid) % a.grid), n = (int)(row / ((int64_t)a.grid * a.grid));
      const int Y = py * a.patch + ky, X = px * a.patch + kx;
      const float sy = a.align ? scale * (float)Y : scale * ((float)Y + 0.5f) - 0.5f;
      const float sx = a.align ? scale * (float)X : scale * ((float)X + 0.5f) - 0.5f;
      const float fy = floorf(sy), fx = floorf(sx);
      const int iy = (int)fy, ix = (int)fx;
      const float ty = sy - fy, tx = sx - fx;
      const float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
      const float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
      const float* img = a.rgb + ((int64_t)n * 3 + c) * a.in_size * a.in_size;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int yy = min(max(iy - 1 + i, 0), a.in_size - 1);
        float r = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int xx = min(max(ix - 1 + j, 0), a.in_size - 1);
          r = fmaf(wx[j], __ldg(img + yy * a.in_size + xx), r);
        }
        acc = fmaf(wy[i], r, acc);
      }
      // selects, not a.mean[c]: a run-time index would copy the argument struct to local memory
      const float mean = c == 0 ? a.mean[0] : c == 1 ? a.mean[1] : a.mean[2];
      const float stdv = c == 0 ? a.stdv[0] : c == 1 ? a.stdv[1] : a.stdv[2];
      val = __fdiv_rn(acc - mean, stdv);
    }
    __half hh, ll;
    split_half(val, hh, ll);
    a.hi[e] = hh;
    if (a.lo) a.lo[e] = ll;
  }
}

int resize_patch(oryon_handle* h, const float* rgb, int n, int in_size, int out_size, int patch, int align_corners, const float mean[3],
                 const float stdv[3], __half* hi, __half* lo, int ld, cudaStream_t st) {
  ResizeArgs a;
  a.rgb = rgb, a.n = n, a.in_size = in_size, a.out_size = out_size, a.patch = patch, a.grid = out_size / patch, a.align = align_corners;
  for (int i = 0; i < 3; ++i) a.mean[i] = mean[i], a.stdv[i] = stdv[i], a.inv_std[i] = 1.f / stdv[i];
  a.hi = hi, a.lo = lo, a.ld = ld;
  const int64_t total = (int64_t)n * a.grid * a.grid * ld;
  h->span_begin(KID_ELTWISE, st);
  resize_patch_kernel<<<blocks_for(total, 256, h->sm_count * 32), 256, 0, st>>>(a);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm (one warp per output row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(LnArgs a) {
  const int lane = threadIdx.x & 31;
  const int nper = a.C >> 5;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < a.rows; r += gridDim.x * 8) {
    const int src = a.row_map ? a.row_map[r] : r;
    float v[32];
    if (src >= 0) {
      const float* xr = a.x + (int64_t)src * a.ldx;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nper) v[i] = xr[lane + 32 * i];
      warp_layernorm(v, nper, a.C, a.eps, a.gamma, a.beta, lane);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
    }
    if (a.out32) {
      float* o = a.out32 + (int64_t)r * a.ld32;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nper) o[lane + 32 * i] = v[i];
    }
    if (a.out_hi) {
      const int64_t o = (int64_t)r * a.ldh;
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < nper) {
          __half hh, ll;
          split_half(v[i], hh, ll);
          a.out_hi[o + lane + 32 * i] = hh;
          if (a.out_lo) a.out_lo[o + lane + 32 * i] = ll;
        }
      for (int c = a.C + lane; c < a.ldh; c += 32) {
        float x = 0.f;
        if (c < a.C + a.cat_C && src >= 0) x = a.cat[(int64_t)src * a.cat_C + (c - a.C)];
        __half hh, ll;
        split_half(x, hh, ll);
        a.out_hi[o + c] = hh;
        if (a.out_lo) a.out_lo[o + c] = ll;
      }
    }
  }
}

// Vectorised form (C % 128 == 0, 16-byte aligned rows): lane holds columns 128*i + 4*lane .. +3, float4 loads / 8-byte split
// stores -- a quarter of the memory instructions of the scalar kernel for the 768 / 1024-wide rows of the towers.
__global__ void __launch_bounds__(256) layernorm_vec_kernel(LnArgs a) {
  const int lane = threadIdx.x & 31;
  const int nq = a.C >> 7;   // float4 per lane, <= 8
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < a.rows; r += gridDim.x * 8) {
    const int src = a.row_map ? a.row_map[r] : r;
    float4 v[8];
    if (src >= 0) {
      const float4* xr = reinterpret_cast<const float4*>(a.x + (int64_t)src * a.ldx) + lane;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) v[i] = xr[32 * i];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      const float mean = warp_sum(s) / (float)a.C;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) {
          const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
          q = fmaf(d0, d0, q), q = fmaf(d1, d1, q), q = fmaf(d2, d2, q), q = fmaf(d3, d3, q);
        }
      const float rstd = rsqrtf(warp_sum(q) / (float)a.C + a.eps);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(a.gamma) + 32 * i + lane);
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.beta) + 32 * i + lane);
          v[i].x = (v[i].x - mean) * rstd * g.x + b.x, v[i].y = (v[i].y - mean) * rstd * g.y + b.y;
          v[i].z = (v[i].z - mean) * rstd * g.z + b.z, v[i].w = (v[i].w - mean) * rstd * g.w + b.w;
        }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (a.out32) {
      float4* o = reinterpret_cast<float4*>(a.out32 + (int64_t)r * a.ld32) + lane;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) o[32 * i] = v[i];
    }
    if (a.out_hi) {
      const int64_t o = (int64_t)r * a.ldh;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nq) {
          __half h0, h1, h2, h3, l0, l1, l2, l3;
          split_half(v[i].x, h0, l0), split_half(v[i].y, h1, l1), split_half(v[i].z, h2, l2), split_half(v[i].w, h3, l3);
          const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned*>(&ha), pk.y = *reinterpret_cast<const unsigned*>(&hb);
          *reinterpret_cast<uint2*>(a.out_hi + o + 128 * i + 4 * lane) = pk;
          if (a.lo_format == gemm::LO_F8X) {
            gemm::store_f8x_act4(a.out_lo + o, 128 * i + 4 * lane, clamp_h(v[i].x), clamp_h(v[i].y), clamp_h(v[i].z), clamp_h(v[i].w));
          } else if (a.out_lo) {
            const __half2 la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
            pk.x = *reinterpret_cast<const unsigned*>(&la), pk.y = *reinterpret_cast<const unsigned*>(&lb);
            *reinterpret_cast<uint2*>(a.out_lo + o + 128 * i + 4 * lane) = pk;
          }
        }
      for (int c = a.C + lane; c < a.ldh; c += 32) {
        float x = 0.f;
        if (c < a.C + a.cat_C && src >= 0) x = a.cat[(int64_t)src * a.cat_C + (c - a.C)];
        __half hh, ll;
        split_half(x, hh, ll);
        a.out_hi[o + c] = hh;
        if (a.out_lo) a.out_lo[o + c] = ll;
      }
    }
  }
}

// Narrow rows (C = 128 * NQ, NQ <= 2: the swin / fusion stages, 300 000 rows of 512 bytes per launch): one row per warp iteration
// keeps a single 16-byte load per lane in flight and the kernel ran at 23 % of the HBM peak (ncu launch list, round 2: ~208 us for
// 314 MB).  Here a warp works on ROWS rows at a time -- all their loads (and the row_map lookups) are issued before the first
// reduction -- with the same per-row arithmetic and summation order as layernorm_vec_kernel (bit-identical results).
template <int NQ, int ROWS>
__global__ void __launch_bounds__(256) layernorm_vec_rows_kernel(LnArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  for (int r0 = wid * ROWS; r0 < a.rows; r0 += nw * ROWS) {
    int src[ROWS];
    float4 v[ROWS][NQ];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const int r = r0 + j;
      src[j] = r < a.rows ? (a.row_map ? a.row_map[r] : r) : -1;
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      if (src[j] >= 0) {
        const float4* xr = reinterpret_cast<const float4*>(a.x + (int64_t)src[j] * a.ldx) + lane;
#pragma unroll
        for (int i = 0; i < NQ; ++i) v[j][i] = xr[32 * i];
      } else {
#pragma unroll
        for (int i = 0; i < NQ; ++i) v[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float4 g[NQ], b[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      g[i] = __ldg(reinterpret_cast<const float4*>(a.gamma) + 32 * i + lane);
      b[i] = __ldg(reinterpret_cast<const float4*>(a.beta) + 32 * i + lane);
    }
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const int r = r0 + j;
      if (r >= a.rows) break;
      if (src[j] >= 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NQ; ++i) s += (v[j][i].x + v[j][i].y) + (v[j][i].z + v[j][i].w);
        const float mean = warp_sum(s) / (float)a.C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          const float d0 = v[j][i].x - mean, d1 = v[j][i].y - mean, d2 = v[j][i].z - mean, d3 = v[j][i].w - mean;
          q = fmaf(d0, d0, q), q = fmaf(d1, d1, q), q = fmaf(d2, d2, q), q = fmaf(d3, d3, q);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)a.C + a.eps);
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          v[j][i].x = (v[j][i].x - mean) * rstd * g[i].x + b[i].x, v[j][i].y = (v[j][i].y - mean) * rstd * g[i].y + b[i].y;
          v[j][i].z = (v[j][i].z - mean) * rstd * g[i].z + b[i].z, v[j][i].w = (v[j][i].w - mean) * rstd * g[i].w + b[i].w;
        }
      }
      if (a.out32) {
        float4* o = reinterpret_cast<float4*>(a.out32 + (int64_t)r * a.ld32) + lane;
#pragma unroll
        for (int i = 0; i < NQ; ++i) o[32 * i] = v[j][i];
      }
      if (a.out_hi) {
        const int64_t o = (int64_t)r * a.ldh;
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
          __half h0, h1, h2, h3, l0, l1, l2, l3;
          split_half(v[j][i].x, h0, l0), split_half(v[j][i].y, h1, l1), split_half(v[j][i].z, h2, l2), split_half(v[j][i].w, h3, l3);
          const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned*>(&ha), pk.y = *reinterpret_cast<const unsigned*>(&hb);
          *reinterpret_cast<uint2*>(a.out_hi + o + 128 * i + 4 * lane) = pk;
          if (a.lo_format == gemm::LO_F8X) {
            gemm::store_f8x_act4(a.out_lo + o, 128 * i + 4 * lane, clamp_h(v[j][i].x), clamp_h(v[j][i].y), clamp_h(v[j][i].z), clamp_h(v[j][i].w));
          } else if (a.out_lo) {
            const __half2 la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
            pk.x = *reinterpret_cast<const unsigned*>(&la), pk.y = *reinterpret_cast<const unsigned*>(&lb);
            *reinterpret_cast<uint2*>(a.out_lo + o + 128 * i + 4 * lane) = pk;
          }
        }
        for (int c = a.C + lane; c < a.ldh; c += 32) {
          float x = 0.f;
          if (c < a.C + a.cat_C && src[j] >= 0) x = a.cat[(int64_t)src[j] * a.cat_C + (c - a.C)];
          __half hh, ll;
          split_half(x, hh, ll);
          a.out_hi[o + c] = hh;
          if (a.out_lo) a.out_lo[o + c] = ll;
        }
      }
    }
  }
}

int layernorm(oryon_handle* h, const LnArgs& a, cudaStream_t st) {
  ORYON_REQUIRE(a.C > 0 && a.C % 32 == 0 && a.C <= 1024, "layernorm: C=%d unsupported", a.C);
  if (a.rows <= 0) return ORYON_OK;
  h->span_begin(KID_NORM, st);
  auto al = [](const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  const bool vec = a.C % 128 == 0 && a.gamma && a.beta && al(a.x, 16) && a.ldx % 4 == 0 && al(a.gamma, 16) && al(a.beta, 16) &&
                   (!a.out32 || (al(a.out32, 16) && a.ld32 % 4 == 0)) &&
                   (!a.out_hi || (al(a.out_hi, 8) && a.ldh % 4 == 0 && (!a.out_lo || al(a.out_lo, 8))));
  ORYON_REQUIRE(a.lo_format == gemm::LO_F16 || (vec && a.out_hi && a.out_lo && a.ldh == a.C && a.cat_C == 0 && a.C % 64 == 0),
                "layernorm: the 8-bit cross-term output needs the vector path, ldh == C and no concatenated source");
  const int grid = std::min((a.rows + 7) / 8, h->sm_count * 16);
  static const bool rows_v1 = getenv("ORYON_LN_V1") != nullptr;   // A/B switch: one row per warp iteration for every width
  if (vec && !rows_v1 && a.C == 128) layernorm_vec_rows_kernel<1, 4><<<std::min((a.rows + 31) / 32, h->sm_count * 16), 256, 0, st>>>(a);
  else if (vec && !rows_v1 && a.C == 256) layernorm_vec_rows_kernel<2, 4><<<std::min((a.rows + 31) / 32, h->sm_count * 16), 256, 0, st>>>(a);
  else if (vec) layernorm_vec_kernel<<<grid, 256, 0, st>>>(a);
  else layernorm_kernel<<<grid, 256, 0, st>>>(a);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// attention: 16 queries of one (sequence, head) per CTA, scores in shared memory (never in HBM)
// ------------------------------------------------------------------------------------------------
constexpr int kQT = 16;    // queries per CTA
constexpr int kQS = 20;    // padded stride of the [key][query] score tile
constexpr int kKC = 64;    // keys per staged chunk

template <int D>
__global__ void __launch_bounds__(256) attention_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;                        // [D][kQS]
  float* KVs = Qs + D * kQS;             // [kKC][D+1]
  float* St = KVs + kKC * (D + 1);       // [S][kQS]
  __shared__ float row_sum[kQT];
  const int q0 = blockIdx.x * kQT, hh = blockIdx.y, seq = blockIdx.z;
  const int S = a.S;
  const float* base = a.qkv + (int64_t)seq * S * a.ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int e = threadIdx.x; e < kQT * D; e += 256) {
    const int qi = e / D, dd = e % D;
    Qs[dd * kQS + qi] = (q0 + qi < S) ? base[(int64_t)(q0 + qi) * a.ld + hh * D + dd] * a.scale : 0.f;
  }
  const float* bias = a.bias ? a.bias + (int64_t)hh * S * S : nullptr;
  const float* mask = a.mask ? a.mask + (int64_t)(seq % a.n_win) * S * S : nullptr;

  for (int c0 = 0; c0 < S; c0 += kKC) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKC * D; e += 256) {
      const int i = e / D, dd = e % D;
      KVs[i * (D + 1) + dd] = (c0 + i < S) ? base[(int64_t)(c0 + i) * a.ld + a.off_k + hh * D + dd] : 0.f;
    }
    __syncthreads();
    const int i = threadIdx.x & (kKC - 1), og = threadIdx.x / kKC;  // 4 query groups of 4
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int dd = 0; dd < D; ++dd) {
      const float kv = KVs[i * (D + 1) + dd];
      const float4 q = *reinterpret_cast<const float4*>(Qs + dd * kQS + og * 4);
      acc[0] = fmaf(q.x, kv, acc[0]), acc[1] = fmaf(q.y, kv, acc[1]), acc[2] = fmaf(q.z, kv, acc[2]), acc[3] = fmaf(q.w, kv, acc[3]);
    }
    const int key = c0 + i;
    if (key < S) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int qi = og * 4 + j, qpos = q0 + qi;
        float s = acc[j];
        if (qpos < S) {
          if (bias) s += __ldg(bias + (int64_t)qpos * S + key);
          if (mask) s += __ldg(mask + (int64_t)qpos * S + key);
          if (a.causal && key > qpos) s = -INFINITY;
        } else {
          s = 0.f;
        }
        St[key * kQS + qi] = s;
      }
    }
  }
  __syncthreads();
  for (int r = 0; r < 2; ++r) {
    const int qi = warp * 2 + r;
    float m = -INFINITY;
    for (int k = lane; k < S; k += 32) m = fmaxf(m, St[k * kQS + qi]);
    m = warp_max(m);
    float s = 0.f;
    for (int k = lane; k < S; k += 32) {
      const float e = expf(St[k * kQS + qi] - m);
      St[k * kQS + qi] = e;
      s += e;
    }
    s = warp_sum(s);
    if (lane == 0) row_sum[qi] = s;
  }
  constexpr int QPT = kQT / (256 / D);   // queries per thread in the PV phase: 4 (D=64) or 2 (D=32)
  const int dd = threadIdx.x % D, og = threadIdx.x / D;
  float acc[QPT];
#pragma unroll
  for (int j = 0; j < QPT; ++j) acc[j] = 0.f;
  for (int c0 = 0; c0 < S; c0 += kKC) {
    __syncthreads();
    for (int e = threadIdx.x; e < kKC * D; e += 256) {
      const int i = e / D, d2 = e % D;
      KVs[i * (D + 1) + d2] = (c0 + i < S) ? base[(int64_t)(c0 + i) * a.ld + a.off_v + hh * D + d2] : 0.f;
    }
    __syncthreads();
    const int lim = min(kKC, S - c0);
#pragma unroll 4
    for (int i = 0; i < lim; ++i) {
      const float v = KVs[i * (D + 1) + dd];
      const float* p = St + (c0 + i) * kQS + og * QPT;
      if constexpr (QPT == 4) {
        const float4 w = *reinterpret_cast<const float4*>(p);
        acc[0] = fmaf(w.x, v, acc[0]), acc[1] = fmaf(w.y, v, acc[1]), acc[2] = fmaf(w.z, v, acc[2]), acc[3] = fmaf(w.w, v, acc[3]);
      } else {
        const float2 w = *reinterpret_cast<const float2*>(p);
        acc[0] = fmaf(w.x, v, acc[0]), acc[1] = fmaf(w.y, v, acc[1]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < QPT; ++j) {
    const int qi = og * QPT + j, qpos = q0 + qi;
    if (qpos < S) {
      __half hv, lv;
      split_half(__fdiv_rn(acc[j], row_sum[qi]), hv, lv);
      const int64_t o = ((int64_t)seq * S + qpos) * a.ldh + hh * D + dd;
      a.out_hi[o] = hv;
      if (a.out_lo) a.out_lo[o] = lv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// window attention (Swin: S = 49 or 144 tokens per window, d = 32): one CTA per (window, head), one query per lane
// ------------------------------------------------------------------------------------------------
// The generic kernel above spends its time in block-wide barriers: a 49-token window gives a 256-thread CTA a few hundred
// FMAs between __syncthreads.  Here the ceil(S / 32) warps of a CTA stage K and V of their (window, head) in shared memory
// once (coalesced 128-byte rows, one barrier), then every lane owns one query: q[32] and the output accumulator stay in
// registers, the key loop reads k_j / v_j as shared-memory broadcasts (no bank conflicts, no further barriers).  First sweep:
// score + bias + mask -> row maximum; second sweep: exp / sum / P V.  KEEP = true parks the scores of the first sweep in a
// per-lane shared-memory row (S <= 64); KEEP = false recomputes the dot product instead (S = 144: the rows would not fit next
// to K and V at a useful occupancy).  Same max-subtracted softmax as the reference.
constexpr int kWinD = 32;

template <bool KEEP>
__global__ void __launch_bounds__(160) window_attention_kernel(AttnArgs a, int s_ld) {
  extern __shared__ __align__(16) float wsm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int seq = blockIdx.x, hh = blockIdx.y;
  const int S = a.S;
  float* Ks = wsm;                  // [S][32]
  float* Vs = Ks + S * kWinD;       // [S][32]
  float* Sc = Vs + S * kWinD;       // KEEP: [warps][32 lanes][s_ld]
  const float* base = a.qkv + (int64_t)seq * S * a.ld;
  {
    // K and V ([S][32] each, adjacent in shared memory) as 2 * S * 8 float4: every thread issues all of its loads before the
    // first store, so the staging costs one memory latency instead of one per row
    constexpr int kMaxIt = 16;   // 2 * 160 * 8 / 160 threads
    const int total = 2 * S * 8;
    float4 tmp[kMaxIt];
#pragma unroll
    for (int i = 0; i < kMaxIt; ++i) {
      const int f = threadIdx.x + i * blockDim.x;
      if (f < total) {
        const int which = f >= S * 8, g = f - which * S * 8;
        tmp[i] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)(g >> 3) * a.ld + (which ? a.off_v : a.off_k) + hh * kWinD) + (g & 7));
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxIt; ++i) {
      const int f = threadIdx.x + i * blockDim.x;
      if (f < total) reinterpret_cast<float4*>(Ks)[f] = tmp[i];
    }
  }
  __syncthreads();
  const int qi = w * 32 + lane;
  const bool ok = qi < S;
  const int qrow = ok ? qi : 0;
  // bias / mask are read with the QUERY index contiguous across lanes (one 128-byte line per instruction instead of 32 scattered
  // sectors): the transposed bias table when the caller has one, and the shift mask through its symmetry mask[i][j] == mask[j][i]
  const bool bt = a.bias_t != nullptr;
  const float* bias = bt ? a.bias_t + (int64_t)hh * S * S + qrow : (a.bias ? a.bias + ((int64_t)hh * S + qrow) * S : nullptr);
  const int bstride = bt ? S : 1;
  const float* mask = a.mask ? a.mask + (int64_t)(seq % a.n_win) * S * S + qrow : nullptr;
  float* my = Sc + (size_t)(w * 32 + lane) * s_ld;
  float q[kWinD];
  {
    const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)qrow * a.ld + hh * kWinD);
#pragma unroll
    for (int i = 0; i < kWinD / 4; ++i) {
      const float4 t = __ldg(qp + i);
      q[4 * i] = t.x * a.scale, q[4 * i + 1] = t.y * a.scale, q[4 * i + 2] = t.z * a.scale, q[4 * i + 3] = t.w * a.scale;
    }
  }
  auto score = [&](int j) {
    const float4* kp = reinterpret_cast<const float4*>(Ks + j * kWinD);
    // packed FFMA2 (fma.rn.f32x2): the same four IEEE accumulation chains, half the issue slots
    uint64_t s01 = ptx::pack_f32x2(0.f, 0.f), s23 = s01;
#pragma unroll
    for (int i = 0; i < kWinD / 4; ++i) {
      const float4 k = kp[i];
      s01 = ptx::fma_f32x2(ptx::pack_f32x2(q[4 * i], q[4 * i + 1]), ptx::pack_f32x2(k.x, k.y), s01);
      s23 = ptx::fma_f32x2(ptx::pack_f32x2(q[4 * i + 2], q[4 * i + 3]), ptx::pack_f32x2(k.z, k.w), s23);
    }
    float s0, s1, s2, s3;
    ptx::unpack_f32x2(s01, s0, s1), ptx::unpack_f32x2(s23, s2, s3);
    float sc = (s0 + s1) + (s2 + s3);
    if (bias) sc += __ldg(bias + j * bstride);
    if (mask) sc += __ldg(mask + j * S);
    return sc;
  };
  float m = -INFINITY;
  for (int j = 0; j < S; ++j) {
    const float sc = score(j);
    if (KEEP) my[j] = sc;
    m = fmaxf(m, sc);
  }
  uint64_t acc2[kWinD / 2];
#pragma unroll
  for (int i = 0; i < kWinD / 2; ++i) acc2[i] = ptx::pack_f32x2(0.f, 0.f);
  float l = 0.f;
  for (int j = 0; j < S; ++j) {
    const float p = expf((KEEP ? my[j] : score(j)) - m);
    l += p;
    const uint64_t pp = ptx::pack_f32x2(p, p);
    const float4* vp = reinterpret_cast<const float4*>(Vs + j * kWinD);
#pragma unroll
    for (int i = 0; i < kWinD / 4; ++i) {
      const float4 v = vp[i];
      acc2[2 * i] = ptx::fma_f32x2(pp, ptx::pack_f32x2(v.x, v.y), acc2[2 * i]);
      acc2[2 * i + 1] = ptx::fma_f32x2(pp, ptx::pack_f32x2(v.z, v.w), acc2[2 * i + 1]);
    }
  }
  float acc[kWinD];
#pragma unroll
  for (int i = 0; i < kWinD / 2; ++i) ptx::unpack_f32x2(acc2[i], acc[2 * i], acc[2 * i + 1]);
  if (ok) {
    const int64_t o = ((int64_t)seq * S + qi) * a.ldh + hh * kWinD;
#pragma unroll
    for (int i = 0; i < kWinD / 8; ++i) {   // 8 outputs = 16 bytes of hi, 16 bytes of lo
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __half h0, l0, h1, l1;
        split_half(__fdiv_rn(acc[8 * i + 2 * e], l), h0, l0);
        split_half(__fdiv_rn(acc[8 * i + 2 * e + 1], l), h1, l1);
        const __half2 hp = __halves2half2(h0, h1), lp = __halves2half2(l0, l1);
        ph[e] = *reinterpret_cast<const uint32_t*>(&hp), pl[e] = *reinterpret_cast<const uint32_t*>(&lp);
      }
      *reinterpret_cast<uint4*>(a.out_hi + o + 8 * i) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      if (a.out_lo) *reinterpret_cast<uint4*>(a.out_lo + o + 8 * i) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

int attention(oryon_handle* h, const AttnArgs& a, cudaStream_t st) {
  ORYON_REQUIRE(a.d == 32 || a.d == 64, "attention: head dim %d unsupported", a.d);
  ORYON_REQUIRE(a.S > 0 && a.S <= 1024, "attention: S=%d unsupported", a.S);
  static const bool generic_only = getenv("ORYON_ATTN_GENERIC") != nullptr;   // A/B switch
  auto al = [](const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  if (!generic_only && a.d == kWinD && !a.causal && a.S <= 160 && a.ld % 4 == 0 && a.off_k % 4 == 0 && a.off_v % 4 == 0 && a.ldh % 8 == 0 &&
      al(a.qkv, 16) && al(a.out_hi, 16) && (!a.out_lo || al(a.out_lo, 16))) {
    const int warps = (a.S + 31) / 32;                                               // <= 5
    const bool keep = a.S <= 64;
    const int s_ld = a.S | 1;                                                        // odd row stride: conflict-free per-lane rows
    const size_t smem = (size_t)(2 * a.S * kWinD + (keep ? warps * 32 * s_ld : 0)) * sizeof(float);
    h->span_begin(KID_ATTN, st);
    if (keep) {
      ORYON_CUDA_CHECK(cudaFuncSetAttribute(window_attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      window_attention_kernel<true><<<dim3(a.n_seq, a.heads), 32 * warps, smem, st>>>(a, s_ld);
    } else {
      ORYON_CUDA_CHECK(cudaFuncSetAttribute(window_attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      window_attention_kernel<false><<<dim3(a.n_seq, a.heads), 32 * warps, smem, st>>>(a, s_ld);
    }
    h->span_end(st);
    ORYON_CUDA_CHECK(cudaGetLastError());
    return ORYON_OK;
  }
  const size_t smem = (size_t)(a.d * kQS + kKC * (a.d + 1) + a.S * kQS) * sizeof(float);
  const dim3 grid((a.S + kQT - 1) / kQT, a.heads, a.n_seq);
  h->span_begin(KID_ATTN, st);
  if (a.d == 64) {
    ORYON_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_kernel<64><<<grid, 256, smem, st>>>(a);
  } else {
    ORYON_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attention_kernel<32><<<grid, 256, smem, st>>>(a);
  }
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// softmax over materialised scores / V transpose (tensor-core attention path)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_split_kernel(const float* scores, int64_t rows, int S, int ld_in, __half* hi, __half* lo, int ld_out) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
    const float* x = scores + r * ld_in;
    float v[32];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < S ? x[c] : -INFINITY;
      m = fmaxf(m, v[i]);
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = (lane + 32 * i < S) ? expf(v[i] - m) : 0.f;
      s += v[i];
    }
    s = warp_sum(s);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + 32 * i;
      if (c < ld_out) {
        __half hh, ll;
        split_half(c < S ? __fdiv_rn(v[i], s) : 0.f, hh, ll);
        hi[r * ld_out + c] = hh;
        if (lo) lo[r * ld_out + c] = ll;
      }
    }
  }
}

int softmax_split(oryon_handle* h, const float* scores, int64_t rows, int S, int ld_in, __half* hi, __half* lo, int ld_out, cudaStream_t st) {
  ORYON_REQUIRE(S <= 1024 && ld_out <= 1024, "softmax_split: S=%d unsupported", S);
  h->span_begin(KID_ATTN, st);
  softmax_split_kernel<<<blocks_for(rows, 8, h->sm_count * 32), 256, 0, st>>>(scores, rows, S, ld_in, hi, lo, ld_out);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

__global__ void __launch_bounds__(256) transpose_v_kernel(const __half* qkv_hi, const __half* qkv_lo, int64_t ld, int off_v, int S, int heads, int d,
                                                         __half* vt_hi, __half* vt_lo, int ld_out) {
  __shared__ __half th[32][34], tl[32][34];
  const int s0 = blockIdx.x * 32, d0 = blockIdx.y * 32, sh = blockIdx.z, seq = sh / heads, hd = sh % heads;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int s = s0 + i;
    __half a = __float2half(0.f), b = a;
    if (s < S && d0 + tx < d) {
      const int64_t o = ((int64_t)seq * S + s) * ld + off_v + hd * d + d0 + tx;
      a = qkv_hi[o];
      if (qkv_lo) b = qkv_lo[o];
    }
    th[i][tx] = a, tl[i][tx] = b;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int dd = d0 + i, s = s0 + tx;
    if (dd < d && s < ld_out) {
      const int64_t o = ((int64_t)sh * d + dd) * ld_out + s;
      vt_hi[o] = th[tx][i];
      if (vt_lo) vt_lo[o] = tl[tx][i];
    }
  }
}

int transpose_v(oryon_handle* h, const __half* qkv_hi, const __half* qkv_lo, int64_t ld, int off_v, int n_seq, int S, int heads, int d,
                __half* vt_hi, __half* vt_lo, int ld_out, cudaStream_t st) {
  h->span_begin(KID_TRANSPOSE, st);
  transpose_v_kernel<<<dim3((ld_out + 31) / 32, (d + 31) / 32, n_seq * heads), 256, 0, st>>>(qkv_hi, qkv_lo, ld, off_v, S, heads, d, vt_hi, vt_lo,
                                                                                         ld_out);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// embeddings
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clip_embed_ln_kernel(const float* patch, const float* cls, const float* pos, const float* gamma,
                                                           const float* beta, int n, int T, int C, float* x) {
  const int lane = threadIdx.x & 31, nper = C >> 5;
  const int rows = n * (T + 1);
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
    const int img = r / (T + 1), tok = r % (T + 1);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nper) {
        const int c = lane + 32 * i;
        const float e = tok == 0 ? cls[c] : patch[((int64_t)img * T + tok - 1) * C + c];
        v[i] = e + pos[(int64_t)tok * C + c];
      }
    warp_layernorm(v, nper, C, 1e-5f, gamma, beta, lane);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nper) x[(int64_t)r * C + lane + 32 * i] = v[i];
  }
}

int clip_embed_ln(oryon_handle* h, const float* patch, const float* cls, const float* pos, const float* gamma, const float* beta, int n,
                  int T, int C, float* x, cudaStream_t st) {
  ORYON_REQUIRE(C % 32 == 0 && C <= 1024, "clip_embed_ln: width %d unsupported", C);
  h->span_begin(KID_NORM, st);
  clip_embed_ln_kernel<<<std::min((n * (T + 1) + 7) / 8, h->sm_count * 16), 256, 0, st>>>(patch, cls, pos, gamma, beta, n, T, C, x);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

__global__ void __launch_bounds__(256) text_embed_kernel(const int32_t* tokens, const float* tok_emb, const float* pos, int64_t total, int L,
                                                        int C, int vocab, float* x) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int c = (int)(e % C);
    const int64_t row = e / C;
    const int l = (int)(row % L);
    int t = tokens[row];
    t = min(max(t, 0), vocab - 1);
    x[e] = tok_emb[(int64_t)t * C + c] + pos[(int64_t)l * C + c];
  }
}

int text_embed(oryon_handle* h, const int32_t* tokens, const float* tok_emb, const float* pos, int n_seq, int L, int C, int vocab, float* x,
               cudaStream_t st) {
  const int64_t total = (int64_t)n_seq * L * C;
  h->span_begin(KID_ELTWISE, st);
  text_embed_kernel<<<blocks_for(total, 256, h->sm_count * 32), 256, 0, st>>>(tokens, tok_emb, pos, total, L, C, vocab, x);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

__global__ void eot_rows_kernel(const int32_t* tokens, int n_seq, int L, int32_t* rows) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seq) return;
  int best = 0, bv = tokens[(int64_t)s * L];
  for (int l = 1; l < L; ++l) {
    const int v = tokens[(int64_t)s * L + l];
    if (v > bv) bv = v, best = l;  // first maximum, as torch.argmax
  }
  rows[s] = s * L + best;
}

int eot_rows(oryon_handle* h, const int32_t* tokens, int n_seq, int L, int32_t* rows, cudaStream_t st) {
  eot_rows_kernel<<<(n_seq + 127) / 128, 128, 0, st>>>(tokens, n_seq, L, rows);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// im2col
// ------------------------------------------------------------------------------------------------
// One warp per output row (pixel): taps in the outer loop (bounds test and source address once per tap), lanes
// over the contiguous channel run of the tap -> coalesced 128-byte reads and 64-byte writes, no per-element division.
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs a) {
  const int Ct = a.C0 + a.C1, kk = a.k * a.k, half = a.k / 2;
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)a.n * a.H * a.W;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    const int x = (int)(row % a.W), y = (int)((row / a.W) % a.H), n = (int)(row / ((int64_t)a.W * a.H));
    __half* hi = a.hi + row * a.ld;
    __half* lo = a.lo ? a.lo + row * a.ld : nullptr;
    for (int tap = 0; tap < kk; ++tap) {
      const int yy = y + tap / a.k - half, xx = x + tap % a.k - half;
      const bool inb = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
      const float* s0 = nullptr;
      const float* s1 = nullptr;
      if (inb) {
        if (a.shuffle0) {
          const int64_t p = ((int64_t)n * (a.H / 2) + yy / 2) * (a.W / 2) + xx / 2;
          s0 = a.src0 + p * 4 * a.C0 + ((yy & 1) * 2 + (xx & 1)) * a.C0;
        } else {
          s0 = a.src0 + (((int64_t)n * a.H + yy) * a.W + xx) * a.C0;
        }
        if (a.C1) s1 = a.src1 + (((int64_t)n * a.H + yy) * a.W + xx) * a.C1;
      }
      for (int c = lane; c < Ct; c += 32) {
        float val = 0.f;
        if (inb) val = c < a.C0 ? __ldg(s0 + c) : __ldg(s1 + (c - a.C0));
        __half hh, ll;
        split_half(val, hh, ll);
        hi[tap * Ct + c] = hh;
        if (lo) lo[tap * Ct + c] = ll;
      }
    }
    for (int c = kk * Ct + lane; c < a.ld; c += 32) {
      hi[c] = __float2half(0.f);
      if (lo) lo[c] = __float2half(0.f);
    }
  }
}

// Vectorised form (channel counts and row pitch multiples of 4): one warp per output row, every lane moves 4 consecutive
// channels of one tap per step -- a float4 load, an 8-byte hi store and an 8-byte lo store (256 contiguous bytes per warp
// store instead of 64), the zero padding up to `ld` included in the same loop.
__global__ void __launch_bounds__(256) im2col_vec_kernel(Im2colArgs a) {
  const int Ct = a.C0 + a.C1, Cq = Ct >> 2, C0q = a.C0 >> 2, kk = a.k * a.k, half = a.k / 2;
  const int nvec = kk * Cq, ldq = a.ld >> 2;
  const int lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)a.n * a.H * a.W;
  for (int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * 8) {
    const int x = (int)(row % a.W), y = (int)((row / a.W) % a.H), n = (int)(row / ((int64_t)a.W * a.H));
    uint2* hi = reinterpret_cast<uint2*>(a.hi + row * a.ld);
    uint2* lo = a.lo ? reinterpret_cast<uint2*>(a.lo + row * a.ld) : nullptr;
    for (int v = lane; v < ldq; v += 32) {
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v < nvec) {
        const int tap = v / Cq, cq = v - tap * Cq;
        const int yy = y + tap / a.k - half, xx = x + tap % a.k - half;
        if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
          const float* src;
          if (cq < C0q) {
            if (a.shuffle0) {
              const int64_t p = ((int64_t)n * (a.H / 2) + yy / 2) * (a.W / 2) + xx / 2;
              src = a.src0 + p * 4 * a.C0 + ((yy & 1) * 2 + (xx & 1)) * a.C0 + 4 * cq;
            } else {
              src = a.src0 + (((int64_t)n * a.H + yy) * a.W + xx) * a.C0 + 4 * cq;
            }
          } else {
            src = a.src1 + (((int64_t)n * a.H + yy) * a.W + xx) * a.C1 + 4 * (cq - C0q);
          }
          val = __ldg(reinterpret_cast<const float4*>(src));
        }
      }
      __half h0, h1, h2, h3, l0, l1, l2, l3;
      split_half(val.x, h0, l0), split_half(val.y, h1, l1), split_half(val.z, h2, l2), split_half(val.w, h3, l3);
      const __half2 ha = __halves2half2(h0, h1), hb = __halves2half2(h2, h3);
      uint2 pk;
      pk.x = *reinterpret_cast<const unsigned*>(&ha), pk.y = *reinterpret_cast<const unsigned*>(&hb);
      hi[v] = pk;
      if (lo) {
        const __half2 la = __halves2half2(l0, l1), lb = __halves2half2(l2, l3);
        pk.x = *reinterpret_cast<const unsigned*>(&la), pk.y = *reinterpret_cast<const unsigned*>(&lb);
        lo[v] = pk;
      }
    }
  }
}

int im2col(oryon_handle* h, const Im2colArgs& a, cudaStream_t st) {
  ORYON_REQUIRE(a.ld >= a.k * a.k * (a.C0 + a.C1), "im2col: ld too small");
  const int64_t rows = (int64_t)a.n * a.H * a.W;
  h->span_begin(KID_IM2COL, st);
  auto al = [](const void* p, size_t n) { return (reinterpret_cast<uintptr_t>(p) % n) == 0; };
  const bool vec = a.C0 % 4 == 0 && a.C1 % 4 == 0 && a.ld % 4 == 0 && al(a.src0, 16) && (!a.C1 || al(a.src1, 16)) && al(a.hi, 8) &&
                   (!a.lo || al(a.lo, 8));
  if (vec) im2col_vec_kernel<<<blocks_for(rows, 8, h->sm_count * 64), 256, 0, st>>>(a);
  else im2col_kernel<<<blocks_for(rows, 8, h->sm_count * 64), 256, 0, st>>>(a);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (16 channels per group) + ReLU, NHWC
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* x, int HW, int C, double* stats) {
  // grid (chunks, n); thread's channel is fixed because 256 % C == 0
  __shared__ double sh[8][2];  // up to 4 groups (C <= 64) x {sum, sumsq}
  const int n = blockIdx.y, groups = C / 16;
  if (threadIdx.x < 16) (&sh[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  const int64_t per_img = (int64_t)HW * C;
  const int64_t chunk = (per_img + gridDim.x - 1) / gridDim.x;
  const int64_t chunk_al = (chunk + 255) / 256 * 256;
  const int64_t e0 = (int64_t)blockIdx.x * chunk_al, e1 = min(per_img, e0 + chunk_al);
  const float* xi = x + (int64_t)n * per_img;
  float s = 0.f, q = 0.f;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += 256) {
    const float v = xi[e];
    s += v, q = fmaf(v, v, q);
  }
  const int g = (threadIdx.x % C) / 16;
  atomicAdd(&sh[g][0], (double)s);
  atomicAdd(&sh[g][1], (double)q);
  __syncthreads();
  if (threadIdx.x < groups * 2) atomicAdd(stats + ((int64_t)n * groups + threadIdx.x / 2) * 2 + (threadIdx.x & 1), sh[threadIdx.x / 2][threadIdx.x & 1]);
}

__global__ void __launch_bounds__(256) gn_apply_kernel(float* x, int HW, int C, const float* gamma, const float* beta, const double* stats,
                                                      int64_t total) {
  const int groups = C / 16;
  const double cnt = (double)HW * 16.0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int c = (int)(e % C);
    const int n = (int)(e / ((int64_t)HW * C));
    const double* st = stats + ((int64_t)n * groups + c / 16) * 2;
    const double mean = st[0] / cnt;
    const double var = fmax(st[1] / cnt - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float y = (x[e] - (float)mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    x[e] = fmaxf(y, 0.f);
  }
}

int groupnorm_relu(oryon_handle* h, float* x, int n, int HW, int C, const float* gamma, const float* beta, double* stats, cudaStream_t st) {
  ORYON_REQUIRE(C % 16 == 0 && C <= 64 && 256 % C == 0, "groupnorm_relu: C=%d unsupported", C);
  const int groups = C / 16;
  ORYON_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * groups * n, st));
  const int chunks = std::max(1, std::min(64, (int)(((int64_t)HW * C) / 16384)));
  h->span_begin(KID_NORM, st);
  gn_stats_kernel<<<dim3(chunks, n), 256, 0, st>>>(x, HW, C, stats);
  const int64_t total = (int64_t)n * HW * C;
  gn_apply_kernel<<<blocks_for(total, 256, h->sm_count * 32), 256, 0, st>>>(x, HW, C, gamma, beta, stats, total);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// PatchMerging gather + LayerNorm(4C)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_merge_ln_kernel(const float* x, int n, int H, int W, int C, const float* gamma, const float* beta,
                                                            __half* hi, __half* lo) {
  const int lane = threadIdx.x & 31;
  const int H2 = H / 2, W2 = W / 2, C4 = 4 * C, nper = C4 >> 5;
  const int rows = n * H2 * W2;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
    const int x2 = r % W2, y2 = (r / W2) % H2, img = r / (W2 * H2);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nper) {
        const int c4 = lane + 32 * i, part = c4 / C, c = c4 % C;
        // x0 (even,even) x1 (odd row, even col) x2 (even row, odd col) x3 (odd, odd)
        const int yy = 2 * y2 + (part & 1), xx = 2 * x2 + (part >> 1);
        v[i] = x[(((int64_t)img * H + yy) * W + xx) * C + c];
      }
    warp_layernorm(v, nper, C4, 1e-5f, gamma, beta, lane);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nper) {
        __half hh, ll;
        split_half(v[i], hh, ll);
        hi[(int64_t)r * C4 + lane + 32 * i] = hh;
        if (lo) lo[(int64_t)r * C4 + lane + 32 * i] = ll;
      }
  }
}

int patch_merge_ln(oryon_handle* h, const float* x, int n, int H, int W, int C, const float* gamma, const float* beta, __half* hi, __half* lo,
                   cudaStream_t st) {
  ORYON_REQUIRE(H % 2 == 0 && W % 2 == 0 && 4 * C <= 1024 && C % 8 == 0, "patch_merge_ln: unsupported shape");
  h->span_begin(KID_NORM, st);
  patch_merge_ln_kernel<<<std::min((n * (H / 2) * (W / 2) + 7) / 8, h->sm_count * 16), 256, 0, st>>>(x, n, H, W, C, gamma, beta, hi, lo);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// L2 normalisation of rows -> split
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2norm_split_kernel(const float* x, int rows, int C, __half* hi, __half* lo, int ld) {
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8) {
    const float* xr = x + (int64_t)r * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
    const float nrm = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
    for (int c = lane; c < ld; c += 32) {
      __half hh, ll;
      split_half(c < C ? __fdiv_rn(xr[c], nrm) : 0.f, hh, ll);
      hi[(int64_t)r * ld + c] = hh;
      if (lo) lo[(int64_t)r * ld + c] = ll;
    }
  }
}

int l2norm_split(oryon_handle* h, const float* x, int rows, int C, __half* hi, __half* lo, int ld, cudaStream_t st) {
  h->span_begin(KID_NORM, st);
  l2norm_split_kernel<<<std::min((rows + 7) / 8, h->sm_count * 16), 256, 0, st>>>(x, rows, C, hi, lo, ld);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// text guidance: mean over prompts, renormalise, Linear + ReLU
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) text_guidance_kernel(const float* text, int P, int C, const float* w, const float* b, int out_c, float* out) {
  extern __shared__ float m[];  // [C]
  __shared__ float red[8];
  const int bi = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* t = text + (int64_t)bi * P * C;
  float ss = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += t[(int64_t)p * C + c];
    s = s / (float)P;
    m[c] = s;
    ss = fmaf(s, s, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float nrm = sqrtf(tot);
  for (int j = warp; j < out_c; j += 8) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(w[(int64_t)j * C + c], __fdiv_rn(m[c], nrm), s);
    s = warp_sum(s);
    if (lane == 0) out[(int64_t)bi * out_c + j] = fmaxf(s + b[j], 0.f);
  }
}

int text_guidance(oryon_handle* h, const float* text, int B, int P, int C, const float* w, const float* b, int out_c, float* out,
                  cudaStream_t st) {
  h->span_begin(KID_ELTWISE, st);
  text_guidance_kernel<<<B, 256, C * sizeof(float), st>>>(text, P, C, w, b, out_c, out);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// ClassTransformerLayer, T = 1 (fusion.py:409-434); one CTA per image.  Weights are stored transposed
// ([in][out]) by the loader so that consecutive threads read consecutive addresses.
// ------------------------------------------------------------------------------------------------
constexpr int kCT = 128;   // hidden dim
constexpr int kTok = 16;   // 4 x 4 pooled tokens

__device__ __forceinline__ void ct_layernorm(const float* in, float* out, const float* g, const float* b, int warp, int lane) {
  for (int r = 0; r < 2; ++r) {
    const int tok = warp * 2 + r;
    float v[4], s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = in[tok * kCT + lane + 32 * i], s += v[i];
    const float mean = warp_sum(s) / (float)kCT;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) q = fmaf(v[i] - mean, v[i] - mean, q);
    const float rstd = rsqrtf(warp_sum(q) / (float)kCT + 1e-5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[tok * kCT + lane + 32 * i] = (v[i] - mean) * rstd * g[lane + 32 * i] + b[lane + 32 * i];
  }
}

// out[tok][j] = b[j] + sum_ci in[tok][ci] * Wt[ci][j], j < n_out (multiple of 128); thread <-> (j % 128, 8 tokens)
__device__ __forceinline__ void ct_linear(const float* in, int ld_in, int n_in, const float* Wt, const float* b, int n_out, float* out, int ld_out,
                                          bool relu) {
  for (int j0 = 0; j0 < n_out; j0 += kCT) {
    const int j = j0 + (threadIdx.x & (kCT - 1)), tg = threadIdx.x / kCT;
    float acc[8];
    const float bj = b[j];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = bj;
    for (int ci = 0; ci < n_in; ++ci) {
      const float w = __ldg(Wt + (int64_t)ci * n_out + j);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(in[(tg * 8 + t) * ld_in + ci], w, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) out[(tg * 8 + t) * ld_out + j] = relu ? fmaxf(acc[t], 0.f) : acc[t];
  }
}

__global__ void __launch_bounds__(256) class_transformer_kernel(float* x, const float* text_guid, int B, ClassTfW w) {
  extern __shared__ __align__(16) float ct_sm[];
  float* xp = ct_sm;                        // [16][128] pooled tokens (residual stream)
  float* cat = xp + kTok * kCT;             // [16][256] [LN(x) | guidance]
  float* q = cat + kTok * 2 * kCT;          // [16][128]
  float* k = q + kTok * kCT;
  float* v = k + kTok * kCT;
  float* hid = v + kTok * kCT;              // [16][512]
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xi = x + (int64_t)n * 576 * kCT;
  const float* g = text_guid + (int64_t)(n % B) * kCT;
  {  // AvgPool2d(6): 24x24 -> 4x4
    const int c = threadIdx.x & (kCT - 1), tg = threadIdx.x / kCT;
    for (int t = 0; t < 8; ++t) {
      const int tok = tg * 8 + t, ph = tok / 4, pw = tok % 4;
      float s = 0.f;
      for (int dy = 0; dy < 6; ++dy)
        for (int dx = 0; dx < 6; ++dx) s += xi[((ph * 6 + dy) * 24 + pw * 6 + dx) * kCT + c];
      xp[tok * kCT + c] = s / 36.f;
    }
  }
  __syncthreads();
  ct_layernorm(xp, q /*scratch*/, w.n1_g, w.n1_b, warp, lane);
  __syncthreads();
  for (int e = threadIdx.x; e < kTok * 2 * kCT; e += 256) {
    const int tok = e / (2 * kCT), c = e % (2 * kCT);
    cat[e] = c < kCT ? q[tok * kCT + c] : g[c - kCT];
  }
  __syncthreads();
  ct_linear(cat, 2 * kCT, 2 * kCT, w.q_w, w.q_b, kCT, q, kCT, false);
  ct_linear(cat, 2 * kCT, 2 * kCT, w.k_w, w.k_b, kCT, k, kCT, false);
  ct_linear(cat, 2 * kCT, kCT, w.v_w, w.v_b, kCT, v, kCT, false);
  __syncthreads();
  // linear attention with L = S = 1 (fusion.py:246-266): out = v * (Q.K) / (Q.K + eps) per head (4 heads x 32)
  for (int e = threadIdx.x; e < kTok * kCT; e += 256) {
    const float qq = q[e], kk = k[e];
    q[e] = (qq > 0.f ? qq : expm1f(qq)) + 1.f;
    k[e] = (kk > 0.f ? kk : expm1f(kk)) + 1.f;
  }
  __syncthreads();
  for (int item = warp; item < kTok * 4; item += 8) {  // (token, head): dot over 32 channels
    const int tok = item / 4, hd = item % 4;
    const float s = warp_sum(q[tok * kCT + hd * 32 + lane] * k[tok * kCT + hd * 32 + lane]);
    const float z = 1.f / (s + 1e-6f);
    xp[tok * kCT + hd * 32 + lane] += v[tok * kCT + hd * 32 + lane] * s * z;
  }
  __syncthreads();
  ct_layernorm(xp, q, w.n2_g, w.n2_b, warp, lane);
  __syncthreads();
  ct_linear(q, kCT, kCT, w.m0_w, w.m0_b, 4 * kCT, hid, 4 * kCT, true);
  __syncthreads();
  ct_linear(hid, 4 * kCT, 4 * kCT, w.m2_w, w.m2_b, kCT, k /*scratch*/, kCT, false);
  __syncthreads();
  for (int e = threadIdx.x; e < kTok * kCT; e += 256) xp[e] += k[e];
  __syncthreads();
  // bilinear 4x4 -> 24x24, align_corners=True, added to x (fusion.py:430-433)
  const float scale = 3.f / 23.f;
  for (int e = threadIdx.x; e < 576 * kCT; e += 256) {
    const int c = e % kCT, pix = e / kCT, py = pix / 24, px = pix % 24;
    const float sy = scale * (float)py, sx = scale * (float)px;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, 3), x1 = min(x0 + 1, 3);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float top = (1.f - lx) * xp[(y0 * 4 + x0) * kCT + c] + lx * xp[(y0 * 4 + x1) * kCT + c];
    const float bot = (1.f - lx) * xp[(y1 * 4 + x0) * kCT + c] + lx * xp[(y1 * 4 + x1) * kCT + c];
    xi[e] += (1.f - ly) * top + ly * bot;
  }
}

int class_transformer(oryon_handle* h, float* x, const float* text_guid, int n, int B, const ClassTfW& w, cudaStream_t st) {
  h->span_begin(KID_ELTWISE, st);
  const size_t smem = (size_t)kTok * kCT * (1 + 2 + 3 + 4) * sizeof(float);
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(class_transformer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  class_transformer_kernel<<<n, 256, smem, st>>>(x, text_guid, B, w);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// decoder head + layout change
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_conv_kernel(const float* x, int n, int H, int W, const float* w, const float* b, float* logits) {
  __shared__ float ws[9 * 32];  // [tap][c]
  for (int e = threadIdx.x; e < 288; e += 256) {
    const int tap = e / 32, c = e % 32;
    ws[e] = w[c * 9 + tap];  // weight [1][32][3][3]
  }
  __syncthreads();
  const int64_t total = (int64_t)n * H * W;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < total; p += (int64_t)gridDim.x * 256) {
    const int xx = (int)(p % W), yy = (int)((p / W) % H), img = (int)(p / ((int64_t)W * H));
    float acc = b[0];
    for (int tap = 0; tap < 9; ++tap) {
      const int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
      if (y2 < 0 || y2 >= H || x2 < 0 || x2 >= W) continue;
      const float4* src = reinterpret_cast<const float4*>(x + (((int64_t)img * H + y2) * W + x2) * 32);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = __ldg(src + i);
        acc = fmaf(v.x, ws[tap * 32 + 4 * i], acc), acc = fmaf(v.y, ws[tap * 32 + 4 * i + 1], acc);
        acc = fmaf(v.z, ws[tap * 32 + 4 * i + 2], acc), acc = fmaf(v.w, ws[tap * 32 + 4 * i + 3], acc);
      }
    }
    logits[p] = acc;
  }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* x, int HW, int C, float* out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (p0 + i < HW && c0 + tx < C) tile[i][tx] = x[((int64_t)n * HW + p0 + i) * C + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < C && p0 + tx < HW) out[((int64_t)n * C + c0 + i) * HW + p0 + tx] = tile[tx][i];
}

int nhwc_to_nchw(oryon_handle* h, const float* x, int n, int HW, int C, float* out, cudaStream_t st) {
  h->span_begin(KID_ELTWISE, st);
  nhwc_to_nchw_kernel<<<dim3((HW + 31) / 32, (C + 31) / 32, n), 256, 0, st>>>(x, HW, C, out);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

int decoder_head(oryon_handle* h, const float* x, int n, int H, int W, const float* w, const float* b, float* logits, float* featmap,
                 cudaStream_t st) {
  const int64_t total = (int64_t)n * H * W;
  h->span_begin(KID_ELTWISE, st);
  head_conv_kernel<<<blocks_for(total, 256, h->sm_count * 32), 256, 0, st>>>(x, n, H, W, w, b, logits);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return nhwc_to_nchw(h, x, n, H * W, 32, featmap, st);
}

}  // namespace net
}  // namespace oryon
