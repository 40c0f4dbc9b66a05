// a11 -- PointDSC registration, test mode (reference models/pointdsc/PointDSC.py:128-197,
// models/pointdsc/common.py:7-69, utils/pointdsc/init.py:10-29), for P independent correspondence
// sets at once.  Everything stays on the device: no SVD round trip to the host (reference
// common.py:36-37), no host-synchronised early exits (PointDSC.py:354, :426).
//
//   pdsc_sc_kernel        spatial-consistency matrix SC[n][n] (PointDSC.py:150-153)                 HBM/L2
//   pdsc_prologue_kernel  corr_pos - mean, layer0, PointCN_0, q/k/v projections of layer 0          fp32 FMA
//   pdsc_layer_kernel     one NonLocal layer for a tile of 16 points: SC-gated attention over all
//                         n points, message MLP, residual, then the per-point part of the next
//                         layer (PointCN + q/k/v) or, after the last layer, L2-normalised features
//                         and the confidence MLP                                                   fp32 FMA
//   pdsc_seeds_kernel     parallel NMS + ordered top-k (pick_seeds, PointDSC.py:199-217)
//   pdsc_knn_kernel       feature-space kNN of each seed (common.py:48-69)
//   pdsc_compat_kernel    k x k feature (x) spatial compatibility of each seed (PointDSC.py:258-278)
//   pdsc_power_kernel     power iteration with the reference's global early exit (PointDSC.py:338-358)
//                         + weighted Kabsch with an in-register 3x3 SVD (common.py:7-45)          warp level
//   pdsc_fitness_kernel   inlier count of every hypothesis (PointDSC.py:325-332)
//   pdsc_refine_kernel    argmax hypothesis + <=20 re-weighted Kabsch iterations (PointDSC.py:403-438)
//
// Layouts (per pair p, npad = cap rounded up to 16): SC [npad][npad]; F1/Q/V point-major [npad][C];
// K channel-major [C][npad]; two generations of F1/Q/K/V (layer l writes generation (l+1)&1 while every
// CTA still reads generation l&1).  Weights: conv weights transposed to [Cin][Cout] with eval-mode
// BatchNorm folded in (done once on the host in load_weights).
#include <algorithm>
#include <cmath>

#include <cstdlib>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx_sm100.cuh"

namespace oryon {
namespace pdsc {

constexpr int C = 128;       // num_channels of the released model (checked at load time)
constexpr int TP = 16;       // points per tile
constexpr int TPS = 20;      // padded shared-memory row stride (floats; keeps float4 alignment, spreads banks)
constexpr int NT = 256;      // threads per CTA of the per-point kernels
constexpr int kMaxK = 64;    // neighbourhood size limit (release: 40)

// (a0, a1) += (x0, x1) * (y, y) as ONE packed FFMA2 (sm_100 fma.rn.f32x2): the same two IEEE fused multiply-adds, half the issue
// slots.  The per-point layers and the attention of the NonLocal blocks are issue-bound fp32 FMA loops.
__device__ __forceinline__ void fma2(float& a0, float& a1, float x0, float x1, float y) {
  ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(x0, x1), ptx::pack_f32x2(y, y), ptx::pack_f32x2(a0, a1)), a0, a1);
}

struct LayerW {
  const float *pcn_w, *pcn_b;            // [C][C], [C]      PointCN conv + BN folded
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b;
  const float *m0_w, *m0_b;              // [C][C/2]         fc_message.0 + BN folded
  const float *m1_w, *m1_b;              // [C/2][C/2]       fc_message.3 + BN folded
  const float *m2_w, *m2_b;              // [C/2][C]         fc_message.6
};

struct Model {
  oryon_pointdsc_config cfg;
  DeviceBuffer blob;
  const float *l0_w = nullptr, *l0_b = nullptr;   // [in_dim][C]
  std::vector<LayerW> layers;
  const float *c0_w = nullptr, *c0_b = nullptr, *c1_w = nullptr, *c1_b = nullptr, *c2_w = nullptr, *c2_b = nullptr;
  DeviceBuffer layer_table;                        // LayerW[num_layers] on the device
  // tensor-core path: the same folded weights as GEMM operands, [cout][cin] fp16 split pairs (cin = 128 or 64: already a multiple of 64)
  DeviceBuffer tc_blob;
  struct TcLayer {
    const __half *pcn_hi, *pcn_lo, *qk_hi, *qk_lo, *v_hi, *v_lo, *m0_hi, *m0_lo, *m1_hi, *m1_lo, *m2_hi, *m2_lo;
    const float *pcn_b, *qk_b, *v_b, *m0_b, *m1_b, *m2_b;
  };
  std::vector<TcLayer> tc;
};

struct PairMeta {
  int n;        // correspondences of this pair
  int n_seeds;  // int(n * ratio)
  int k;        // min(k, n-1)
};

// ------------------------------------------------------------------------------------------------
// per-point linear layer on a tile held in shared memory:  Y[co][p] = act(b[co] + sum_ci W[ci][co] X[ci][p])
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, bool RELU>
__device__ __forceinline__ void point_linear(const float* __restrict__ Wt, const float* __restrict__ bias, const float* Xs, float* Ys) {
  constexpr int G = NT / COUT;   // point groups
  constexpr int PPT = TP / G;    // points per thread
  static_assert(NT % COUT == 0 && TP % G == 0 && (PPT == 2 || PPT == 4 || PPT == 8), "tile mapping");
  const int co = threadIdx.x % COUT, pg = threadIdx.x / COUT;
  float acc[PPT];
  const float b = __ldg(bias + co);
#pragma unroll
  for (int p = 0; p < PPT; ++p) acc[p] = b;
  const float* xs = Xs + pg * PPT;
#pragma unroll 16
  for (int ci = 0; ci < CIN; ++ci) {
    const float w = __ldg(Wt + ci * COUT + co);
    if constexpr (PPT == 2) {
      const float2 x = *reinterpret_cast<const float2*>(xs + ci * TPS);
      fma2(acc[0], acc[1], x.x, x.y, w);
    } else {
#pragma unroll
      for (int v = 0; v < PPT / 4; ++v) {
        const float4 x = *reinterpret_cast<const float4*>(xs + ci * TPS + v * 4);
        fma2(acc[v * 4 + 0], acc[v * 4 + 1], x.x, x.y, w);
        fma2(acc[v * 4 + 2], acc[v * 4 + 3], x.z, x.w, w);
      }
    }
  }
#pragma unroll
  for (int p = 0; p < PPT; ++p) Ys[co * TPS + pg * PPT + p] = RELU ? fmaxf(acc[p], 0.f) : acc[p];
}

// tile [C][TPS] in shared memory -> point-major global [npad][C]
__device__ __forceinline__ void store_point_major(const float* buf, float* g, int base) {
  const int c = threadIdx.x % C, pg = threadIdx.x / C;
#pragma unroll
  for (int p = 0; p < TP / (NT / C); ++p) {
    const int pt = pg * (TP / (NT / C)) + p;
    g[(size_t)(base + pt) * C + c] = buf[c * TPS + pt];
  }
}
// tile -> channel-major global [C][npad]
__device__ __forceinline__ void store_channel_major(const float* buf, float* g, int base, int npad) {
  const int c = threadIdx.x >> 1, half = threadIdx.x & 1;
  const float4 a = *reinterpret_cast<const float4*>(buf + c * TPS + half * 8);
  const float4 b = *reinterpret_cast<const float4*>(buf + c * TPS + half * 8 + 4);
  float4* dst = reinterpret_cast<float4*>(g + (size_t)c * npad + base + half * 8);
  dst[0] = a, dst[1] = b;
}

struct Buffers {
  float* sc;          // [P][npad][npad]
  float* f1[2];       // [P][npad][C]
  float* q[2];        // [P][npad][C]
  float* kt[2];       // [P][C][npad]
  float* v[2];        // [P][npad][C]
  float* fn;          // [P][npad][C]  L2-normalised final features
  float* conf;        // [P][npad]
  int32_t* seeds;     // [P][smax]
  int32_t* knn;       // [P][smax][kMaxK]
  float* M;           // [P][smax][kMaxK*kMaxK]
  float* seed_trans;  // [P][smax][12]  R row-major (9) + t (3)
  int32_t* fit;       // [P][smax] inlier counts
};

struct Args {
  const float* src;   // [P][cap][3]
  const float* tgt;
  const PairMeta* meta;
  int cap, npad, smax, in_dim;
  Buffers b;
  const LayerW* layers;
  const float *l0_w, *l0_b, *c0_w, *c0_b, *c1_w, *c1_b, *c2_w, *c2_b;
  float sigma_d2, sigma2, nms_radius, inlier_th;
  int num_iterations;
};

// ------------------------------------------------------------------------------------------------
// SC matrix (PointDSC.py:150-153):  clamp(1 - (|si-sj| - |ti-tj|)^2 / sigma_spat^2, 0)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pdsc_sc_kernel(Args a) {
  const int p = blockIdx.z;
  const int n = a.meta[p].n;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int o0 = blockIdx.y * 32 + (threadIdx.x >> 5) * 4;
  if (i >= a.npad) return;
  const float* S = a.src + (size_t)p * a.cap * 3;
  const float* T = a.tgt + (size_t)p * a.cap * 3;
  float six = 0, siy = 0, siz = 0, tix = 0, tiy = 0, tiz = 0;
  if (i < n) six = S[i * 3], siy = S[i * 3 + 1], siz = S[i * 3 + 2], tix = T[i * 3], tiy = T[i * 3 + 1], tiz = T[i * 3 + 2];
  for (int r = 0; r < 4; ++r) {
    const int o = o0 + r;
    if (o >= a.npad) break;
    float val = 0.f;
    if (o < n && i < n) {
      const float dx = S[o * 3] - six, dy = S[o * 3 + 1] - siy, dz = S[o * 3 + 2] - siz;
      const float ex = T[o * 3] - tix, ey = T[o * 3 + 1] - tiy, ez = T[o * 3 + 2] - tiz;
      const float sd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      const float td = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
      const float c = __fsub_rn(sd, td);
      val = fmaxf(__fsub_rn(1.0f, __fdiv_rn(__fmul_rn(c, c), a.sigma_d2)), 0.f);
    }
    a.b.sc[((size_t)p * a.npad + o) * a.npad + i] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// prologue: corr_pos - mean(0) (init.py:18-19), layer0, PointCN_0, q/k/v of layer 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) pdsc_prologue_kernel(Args a) {
  __shared__ __align__(16) float bufA[C * TPS];
  __shared__ __align__(16) float bufB[C * TPS];
  __shared__ double red[8][6];
  __shared__ float mean[6];
  const int p = blockIdx.y, base = blockIdx.x * TP;
  const int n = a.meta[p].n;
  if (base >= n) return;
  const float* S = a.src + (size_t)p * a.cap * 3;
  const float* T = a.tgt + (size_t)p * a.cap * 3;
  {  // column means of cat(src, tgt)
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += NT) {
#pragma unroll
      for (int d = 0; d < 3; ++d) s[d] += S[i * 3 + d], s[3 + d] += T[i * 3 + d];
    }
#pragma unroll
    for (int d = 0; d < 6; ++d) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s[d] += __shfl_xor_sync(0xffffffffu, s[d], off);
    }
    if ((threadIdx.x & 31) == 0)
      for (int d = 0; d < 6; ++d) red[threadIdx.x >> 5][d] = s[d];
    __syncthreads();
    if (threadIdx.x < 6) {
      double t = 0;
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
      mean[threadIdx.x] = (float)(t / n);
    }
    __syncthreads();
  }
  if (threadIdx.x < 6 * TP) {
    const int d = threadIdx.x / TP, pt = threadIdx.x % TP;
    float v = 0.f;
    if (base + pt < n) v = (d < 3 ? S[(base + pt) * 3 + d] : T[(base + pt) * 3 + d - 3]) - mean[d];
    bufA[d * TPS + pt] = v;
  }
  __syncthreads();
  point_linear<6, C, false>(a.l0_w, a.l0_b, bufA, bufB);
  __syncthreads();
  const LayerW L = a.layers[0];
  point_linear<C, C, true>(L.pcn_w, L.pcn_b, bufB, bufA);  // F1
  __syncthreads();
  const size_t po = (size_t)p * a.npad * C;
  store_point_major(bufA, a.b.f1[0] + po, base);
  point_linear<C, C, false>(L.q_w, L.q_b, bufA, bufB);
  __syncthreads();
  store_point_major(bufB, a.b.q[0] + po, base);
  __syncthreads();
  point_linear<C, C, false>(L.k_w, L.k_b, bufA, bufB);
  __syncthreads();
  store_channel_major(bufB, a.b.kt[0] + po, base, a.npad);
  __syncthreads();
  point_linear<C, C, false>(L.v_w, L.v_b, bufA, bufB);
  __syncthreads();
  store_point_major(bufB, a.b.v[0] + po, base);
}

// after the last layer (PointDSC.py:156, :171): L2-normalised features and the confidence MLP for the tile in bufA ([C][TPS])
__device__ __forceinline__ void final_part(const Args& a, int p, int base, int n, size_t po, float* bufA, float* bufB) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // F.normalize(p=2, dim=-1): x / max(|x|, 1e-12); warp w owns points 2w, 2w+1
  for (int r = 0; r < 2; ++r) {
    const int pt = warp * 2 + r;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(bufA[c * TPS + pt], bufA[c * TPS + pt], ss);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    if (base + pt < n)
      for (int c = lane; c < C; c += 32) a.b.fn[po + (size_t)(base + pt) * C + c] = __fdiv_rn(bufA[c * TPS + pt], nrm);
  }
  point_linear<C, 32, true>(a.c0_w, a.c0_b, bufA, bufB);
  __syncthreads();
  point_linear<32, 32, true>(a.c1_w, a.c1_b, bufB, bufA);
  __syncthreads();
  if (threadIdx.x < TP && base + threadIdx.x < n) {
    float s = __ldg(a.c2_b);
    for (int ci = 0; ci < 32; ++ci) s = fmaf(__ldg(a.c2_w + ci), bufA[ci * TPS + threadIdx.x], s);
    a.b.conf[(size_t)p * a.npad + base + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// one NonLocal layer for a tile of 16 points (PointDSC.py:26-45), fused with the per-point part of the
// next layer (PointDSC.py:73-76) or with feature normalisation + confidence MLP (PointDSC.py:156, :171)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) pdsc_layer_kernel(Args a, int layer, int last) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                  // [C][TPS]
  float* bufA = Qs + C * TPS;        // [C][TPS]
  float* bufB = bufA + C * TPS;      // [C][TPS]
  float* St = bufB + C * TPS;        // [npad][TPS]  scores / softmax weights, i-major
  __shared__ float row_inv[TP];

  const int p = blockIdx.y, base = blockIdx.x * TP;
  const int n = a.meta[p].n;
  if (base >= n) return;
  const int gen = layer & 1;
  const size_t po = (size_t)p * a.npad * C;
  const float* Qg = a.b.q[gen] + po;
  const float* Kt = a.b.kt[gen] + po;
  const float* Vg = a.b.v[gen] + po;
  const float* F1 = a.b.f1[gen] + po;
  const float* SC = a.b.sc + (size_t)p * a.npad * a.npad;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // q tile -> shared, transposed to [c][o]
  {
    const int c = threadIdx.x % C, pg = threadIdx.x / C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pt = pg * 8 + j;
      Qs[c * TPS + pt] = Qg[(size_t)(base + pt) * C + c];
    }
  }
  __syncthreads();

  // scores: thread <-> column i, all 16 rows o of the tile
  const float inv_scale = 11.313708498984761f;  // (num_channels // head) ** 0.5, a Python float
  for (int i = threadIdx.x; i < n; i += NT) {
    float acc[TP];
#pragma unroll
    for (int o = 0; o < TP; ++o) acc[o] = 0.f;
#pragma unroll 16
    for (int c = 0; c < C; ++c) {
      const float kv = __ldg(Kt + (size_t)c * a.npad + i);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 q = *reinterpret_cast<const float4*>(Qs + c * TPS + v * 4);
        fma2(acc[v * 4 + 0], acc[v * 4 + 1], q.x, q.y, kv);
        fma2(acc[v * 4 + 2], acc[v * 4 + 3], q.z, q.w, kv);
      }
    }
#pragma unroll
    for (int o = 0; o < TP; ++o) {
      const float sc = (base + o < n) ? __ldg(SC + (size_t)(base + o) * a.npad + i) : 0.f;
      St[i * TPS + o] = __fmul_rn(sc, __fdiv_rn(acc[o], inv_scale));
    }
  }
  __syncthreads();

  // softmax over i for each of the 16 rows: warp w owns rows 2w, 2w+1
  for (int r = 0; r < 2; ++r) {
    const int o = warp * 2 + r;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, St[i * TPS + o]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float s = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float e = expf(St[i * TPS + o] - m);
      St[i * TPS + o] = e;
      s += e;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) row_inv[o] = s;
  }
  __syncthreads();

  // message[c][o] = sum_i w[o][i] V[i][c]; thread <-> (c, 8 rows)
  {
    const int c = threadIdx.x % C, og = threadIdx.x / C;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 8
    for (int i = 0; i < n; ++i) {
      const float v = __ldg(Vg + (size_t)i * C + c);
      const float4 w0 = *reinterpret_cast<const float4*>(St + i * TPS + og * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(St + i * TPS + og * 8 + 4);
      fma2(acc[0], acc[1], w0.x, w0.y, v), fma2(acc[2], acc[3], w0.z, w0.w, v);
      fma2(acc[4], acc[5], w1.x, w1.y, v), fma2(acc[6], acc[7], w1.z, w1.w, v);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bufA[c * TPS + og * 8 + j] = __fdiv_rn(acc[j], row_inv[og * 8 + j]);
  }
  __syncthreads();

  const LayerW L = a.layers[layer];
  point_linear<C, C / 2, true>(L.m0_w, L.m0_b, bufA, bufB);
  __syncthreads();
  point_linear<C / 2, C / 2, true>(L.m1_w, L.m1_b, bufB, bufA);
  __syncthreads();
  point_linear<C / 2, C, false>(L.m2_w, L.m2_b, bufA, bufB);
  __syncthreads();
  {  // residual: feat = F1 + message -> bufA
    const int c = threadIdx.x % C, pg = threadIdx.x / C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pt = pg * 8 + j;
      const float f = (base + pt < n) ? F1[(size_t)(base + pt) * C + c] : 0.f;
      bufA[c * TPS + pt] = (base + pt < n) ? f + bufB[c * TPS + pt] : 0.f;
    }
  }
  __syncthreads();

  if (!last) {
    const LayerW Nx = a.layers[layer + 1];
    const int g2 = gen ^ 1;
    point_linear<C, C, true>(Nx.pcn_w, Nx.pcn_b, bufA, bufB);  // F1 of the next layer
    __syncthreads();
    store_point_major(bufB, a.b.f1[g2] + po, base);
    point_linear<C, C, false>(Nx.q_w, Nx.q_b, bufB, bufA);
    __syncthreads();
    store_point_major(bufA, a.b.q[g2] + po, base);
    __syncthreads();
    point_linear<C, C, false>(Nx.k_w, Nx.k_b, bufB, bufA);
    __syncthreads();
    store_channel_major(bufA, a.b.kt[g2] + po, base, a.npad);
    __syncthreads();
    point_linear<C, C, false>(Nx.v_w, Nx.v_b, bufB, bufA);
    __syncthreads();
    store_point_major(bufA, a.b.v[g2] + po, base);
  } else {
    final_part(a, p, base, n, po, bufA, bufB);
  }
}

// ------------------------------------------------------------------------------------------------
// tensor-core path of the NonLocal network (run_network_tc): the two kernels that are not GEMMs
// ------------------------------------------------------------------------------------------------
// weight[o][i] = softmax_i( SC[o][i] * S[o][i] ) for i < n (PointDSC.py:39: attention * feat_attention, softmax over the last dim),
// written as fp16 split pairs with the K extent (i) zero padded to kp: the A operand of the message GEMM.  One warp per row.
template <int kMaxPer>   // keys per lane held in registers: 16 covers kp <= 512 (the reference's 500 correspondences), 64 the 2048 limit
__global__ void __launch_bounds__(256) pdsc_softmax_kernel(Args a, const float* __restrict__ S, __half* __restrict__ p_hi, __half* __restrict__ p_lo, int kp) {
  const int p = blockIdx.y, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= a.npad) return;
  const int n = a.meta[p].n;
  const size_t row = (size_t)p * a.npad + o;
  __half* ph = p_hi + row * kp;
  __half* pl = p_lo + row * kp;
  float x[kMaxPer];
  const int per = (kp + 31) / 32;
  if (o >= n) {
    for (int t = 0; t < per; ++t) {
      const int i = lane + 32 * t;
      if (i < kp) ph[i] = __float2half_rn(0.f), pl[i] = __float2half_rn(0.f);
    }
    return;
  }
  const float* sr = S + row * a.npad;
  const float* cr = a.b.sc + row * a.npad;
  float m = -INFINITY;
#pragma unroll
  for (int t = 0; t < kMaxPer; ++t) {
    if (t >= per) break;
    const int i = lane + 32 * t;
    x[t] = i < n ? __fmul_rn(__ldg(cr + i), __ldg(sr + i)) : -INFINITY;
    m = fmaxf(m, x[t]);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < kMaxPer; ++t) {
    if (t >= per) break;
    const int i = lane + 32 * t;
    x[t] = i < n ? expf(x[t] - m) : 0.f;
    sum += x[t];
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll
  for (int t = 0; t < kMaxPer; ++t) {
    if (t >= per) break;
    const int i = lane + 32 * t;
    if (i < kp) {
      const float w = __fdiv_rn(x[t], sum);
      const __half hi = __float2half_rn(w);
      ph[i] = hi, pl[i] = __float2half_rn(w - __half2float(hi));
    }
  }
}

// last layer: features [P][npad][C] fp32 (point-major) -> L2-normalised features + confidence (the tail of pdsc_layer_kernel)
__global__ void __launch_bounds__(NT) pdsc_final_kernel(Args a, const float* __restrict__ feat) {
  __shared__ __align__(16) float bufA[C * TPS];
  __shared__ __align__(16) float bufB[C * TPS];
  const int p = blockIdx.y, base = blockIdx.x * TP;
  const int n = a.meta[p].n;
  if (base >= n) return;
  const size_t po = (size_t)p * a.npad * C;
  {
    const int c = threadIdx.x % C, pg = threadIdx.x / C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int pt = pg * 8 + j;
      bufA[c * TPS + pt] = (base + pt < n) ? feat[po + (size_t)(base + pt) * C + c] : 0.f;
    }
  }
  __syncthreads();
  final_part(a, p, base, n, po, bufA, bufB);
}

// ------------------------------------------------------------------------------------------------
// seeds (PointDSC.py:199-217): local maxima under parallel NMS, then argsort(descending)[:max_num].
// Ties are ordered by ascending index (stable descending sort, as ATen's CPU sort).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) pdsc_seeds_kernel(Args a) {
  extern __shared__ float sm[];
  const int p = blockIdx.x;
  const PairMeta pm = a.meta[p];
  const int n = pm.n;
  float* px = sm;
  float* py = px + a.npad;
  float* pz = py + a.npad;
  float* sc = pz + a.npad;
  float* val = sc + a.npad;
  const float* S = a.src + (size_t)p * a.cap * 3;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    px[i] = S[i * 3], py[i] = S[i * 3 + 1], pz[i] = S[i * 3 + 2];
    sc[i] = a.b.conf[(size_t)p * a.npad + i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = px[i], y = py[i], z = pz[i], s = sc[i];
    bool is_max = true;
    for (int j = 0; j < n; ++j) {
      const float dx = x - px[j], dy = y - py[j], dz = z - pz[j];
      const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      is_max &= (s >= sc[j]) | (d >= a.nms_radius);
    }
    val[i] = __fmul_rn(s, is_max ? 1.f : 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = val[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (val[j] > v) | ((val[j] == v) & (j < i));
    if (rank < pm.n_seeds) a.b.seeds[(size_t)p * a.smax + rank] = i;
  }
}

// ------------------------------------------------------------------------------------------------
// kNN of each seed in normalised feature space (common.py:48-69): topk(k+1, smallest) of 2 - 2 x.x^T,
// first entry dropped.  Ordered by (distance, index).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pdsc_knn_kernel(Args a) {
  extern __shared__ float dist[];
  const int p = blockIdx.y, s = blockIdx.x;
  const PairMeta pm = a.meta[p];
  if (s >= pm.n_seeds) return;
  const int n = pm.n;
  const int seed = a.b.seeds[(size_t)p * a.smax + s];
  const float* fn = a.b.fn + (size_t)p * a.npad * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 me = *reinterpret_cast<const float4*>(fn + (size_t)seed * C + lane * 4);
  for (int j = warp; j < n; j += 8) {
    const float4 o = *reinterpret_cast<const float4*>(fn + (size_t)j * C + lane * 4);
    float d = me.x * o.x;
    d = fmaf(me.y, o.y, d), d = fmaf(me.z, o.z, d), d = fmaf(me.w, o.w, d);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    if (lane == 0) dist[j] = __fsub_rn(2.f, __fmul_rn(2.f, d));
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += 256) {
    const float d = dist[j];
    int rank = 0;
    for (int l = 0; l < n; ++l) rank += (dist[l] < d) | ((dist[l] == d) & (l < j));
    if (rank >= 1 && rank <= pm.k) a.b.knn[((size_t)p * a.smax + s) * kMaxK + rank - 1] = j;
  }
}

// ------------------------------------------------------------------------------------------------
// k x k compatibility of one seed's neighbourhood (PointDSC.py:258-281)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pdsc_compat_kernel(Args a) {
  extern __shared__ float sm[];
  const int p = blockIdx.y, s = blockIdx.x;
  const PairMeta pm = a.meta[p];
  if (s >= pm.n_seeds) return;
  const int k = pm.k;
  float* kf = sm;                       // [k][C+1]
  float* ks = kf + kMaxK * (C + 1);     // [k][3]
  float* kt = ks + kMaxK * 3;           // [k][3]
  const int32_t* idx = a.b.knn + ((size_t)p * a.smax + s) * kMaxK;
  const float* fn = a.b.fn + (size_t)p * a.npad * C;
  const float* S = a.src + (size_t)p * a.cap * 3;
  const float* T = a.tgt + (size_t)p * a.cap * 3;
  for (int e = threadIdx.x; e < k * C; e += 256) {
    const int r = e / C, c = e % C;
    kf[r * (C + 1) + c] = fn[(size_t)idx[r] * C + c];
  }
  for (int e = threadIdx.x; e < k * 3; e += 256) ks[e] = S[idx[e / 3] * 3 + e % 3], kt[e] = T[idx[e / 3] * 3 + e % 3];
  __syncthreads();
  float* M = a.b.M + ((size_t)p * a.smax + s) * kMaxK * kMaxK;
  for (int e = threadIdx.x; e < k * k; e += 256) {
    const int r = e / k, c = e % k;
    float val = 0.f;
    if (r != c) {
      const int lo = min(r, c), hi = max(r, c);  // evaluate (lo,hi) so that M is bitwise symmetric
      float dot = 0.f;
      for (int ch = 0; ch < C; ++ch) dot = fmaf(kf[lo * (C + 1) + ch], kf[hi * (C + 1) + ch], dot);
      const float fM = fmaxf(__fsub_rn(1.f, __fdiv_rn(__fsub_rn(1.f, dot), a.sigma2)), 0.f);
      const float dx = ks[lo * 3] - ks[hi * 3], dy = ks[lo * 3 + 1] - ks[hi * 3 + 1], dz = ks[lo * 3 + 2] - ks[hi * 3 + 2];
      const float ex = kt[lo * 3] - kt[hi * 3], ey = kt[lo * 3 + 1] - kt[hi * 3 + 1], ez = kt[lo * 3 + 2] - kt[hi * 3 + 2];
      const float sd = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      const float td = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez)));
      const float cdiff = __fsub_rn(sd, td);
      const float sM = fmaxf(__fsub_rn(1.f, __fdiv_rn(__fmul_rn(cdiff, cdiff), a.sigma_d2)), 0.f);
      val = __fmul_rn(fM, sM);
    }
    M[r * k + c] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// 3x3 machinery (double precision, one thread): proper rotation R = V diag(1,1,det(V U^T)) U^T of
// H = U S V^T (common.py:35-40).  Eigen-decomposition of H^T H by cyclic Jacobi gives V and the singular
// values; the two leading left vectors are u_i = H v_i / s_i and the third of each basis is completed
// right-handed, which yields exactly V diag(1,1,d) U^T without needing the smallest singular triplet.
// ------------------------------------------------------------------------------------------------
__device__ void jacobi_eig3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      if (fabs(A[p][q]) < 1e-300) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int r = 0; r < 3; ++r) {  // A <- A J
        const double arp = A[r][p], arq = A[r][q];
        A[r][p] = c * arp - s * arq, A[r][q] = s * arp + c * arq;
      }
      for (int r = 0; r < 3; ++r) {  // A <- J^T A
        const double apr = A[p][r], aqr = A[q][r];
        A[p][r] = c * apr - s * aqr, A[q][r] = s * apr + c * aqr;
      }
      for (int r = 0; r < 3; ++r) {
        const double vrp = V[r][p], vrq = V[r][q];
        V[r][p] = c * vrp - s * vrq, V[r][q] = s * vrp + c * vrq;
      }
    }
  }
}

__device__ inline void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1], o[1] = a[2] * b[0] - a[0] * b[2], o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline double normalize3(double v[3]) {
  const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (nn > 0) v[0] /= nn, v[1] /= nn, v[2] /= nn;
  return nn;
}

// R (row-major 3x3) from the covariance H = sum w (a - ca)(b - cb)^T
__device__ void kabsch_rotation(const double H[3][3], double R[3][3]) {
  double A[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];  // H^T H
  jacobi_eig3(A, V);
  int o0 = 0, o1 = 1, o2 = 2;  // sort eigenvalues descending
  double l[3] = {A[0][0], A[1][1], A[2][2]};
  if (l[o0] < l[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (l[o0] < l[o2]) { int t = o0; o0 = o2; o2 = t; }
  if (l[o1] < l[o2]) { int t = o1; o1 = o2; o2 = t; }
  double v0[3] = {V[0][o0], V[1][o0], V[2][o0]}, v1[3] = {V[0][o1], V[1][o1], V[2][o1]}, v2[3];
  cross3(v0, v1, v2);
  double u0[3], u1[3], u2[3];
  for (int i = 0; i < 3; ++i) u0[i] = H[i][0] * v0[0] + H[i][1] * v0[1] + H[i][2] * v0[2];
  for (int i = 0; i < 3; ++i) u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
  if (normalize3(u0) <= 0) { u0[0] = 1, u0[1] = 0, u0[2] = 0; }
  const double d01 = u0[0] * u1[0] + u0[1] * u1[1] + u0[2] * u1[2];
  for (int i = 0; i < 3; ++i) u1[i] -= d01 * u0[i];
  if (normalize3(u1) <= 1e-300) {  // rank-1 covariance: any unit vector orthogonal to u0
    const double ax[3] = {fabs(u0[0]) < 0.9 ? 1.0 : 0.0, fabs(u0[0]) < 0.9 ? 0.0 : 1.0, 0.0};
    cross3(u0, ax, u1);
    normalize3(u1);
  }
  cross3(u0, u1, u2);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = v0[i] * u0[j] + v1[i] * u1[j] + v2[i] * u2[j];
}

// ------------------------------------------------------------------------------------------------
// power iteration (global early exit over all seeds of the pair) + weighted Kabsch per seed
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) pdsc_power_kernel(Args a) {
  extern __shared__ float vbuf[];  // [2][smax][kMaxK]
  const int p = blockIdx.x;
  const PairMeta pm = a.meta[p];
  const int S = pm.n_seeds, k = pm.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* v_old = vbuf;
  float* v_new = vbuf + (size_t)a.smax * kMaxK;
  for (int e = threadIdx.x; e < S * kMaxK; e += blockDim.x) v_old[e] = 1.f;
  __syncthreads();
  for (int it = 0; it < a.num_iterations; ++it) {
    int ok = 1;
    for (int s = warp; s < S; s += 16) {
      const float* M = a.b.M + ((size_t)p * a.smax + s) * kMaxK * kMaxK;
      float acc[2] = {0.f, 0.f};
      for (int c = 0; c < k; ++c) {
        const float vc = v_old[s * kMaxK + c];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lane + 32 * h;
          if (r < k) acc[h] = fmaf(__ldg(M + c * k + r), vc, acc[h]);  // M is symmetric: column read, coalesced
        }
      }
      float ss = acc[0] * acc[0] + acc[1] * acc[1];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      const float den = sqrtf(ss) + 1e-6f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if (r < k) {
          const float nv = __fdiv_rn(acc[h], den), ov = v_old[s * kMaxK + r];
          v_new[s * kMaxK + r] = nv;
          ok &= fabsf(nv - ov) <= 1e-8f + 1e-5f * fabsf(ov);  // torch.allclose defaults
        }
      }
    }
    const int all_ok = __syncthreads_and(ok);
    float* t = v_old;
    v_old = v_new, v_new = t;
    if (all_ok) break;
  }
  // weighted Kabsch per seed (PointDSC.py:282-319, common.py:7-45)
  const float* Sp = a.src + (size_t)p * a.cap * 3;
  const float* Tp = a.tgt + (size_t)p * a.cap * 3;
  for (int s = warp; s < S; s += 16) {
    const int32_t* idx = a.b.knn + ((size_t)p * a.smax + s) * kMaxK;
    float wsum = 0.f;
    float w[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      if (r < k) w[h] = v_old[s * kMaxK + r], wsum += w[h];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, off);
    double red[7] = {0, 0, 0, 0, 0, 0, 0};
    double ax[2][3], bx[2][3], wd[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lane + 32 * h;
      wd[h] = 0;
      for (int d = 0; d < 3; ++d) ax[h][d] = bx[h][d] = 0;
      if (r < k) {
        float wn = __fdiv_rn(w[h], wsum + 1e-6f);
        if (wn < 0.f) wn = 0.f;
        wd[h] = wn;
        const int id = idx[r];
        for (int d = 0; d < 3; ++d) ax[h][d] = Sp[id * 3 + d], bx[h][d] = Tp[id * 3 + d];
        red[0] += wd[h];
        for (int d = 0; d < 3; ++d) red[1 + d] += wd[h] * ax[h][d], red[4 + d] += wd[h] * bx[h][d];
      }
    }
#pragma unroll
    for (int q = 0; q < 7; ++q)
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) red[q] += __shfl_xor_sync(0xffffffffu, red[q], off);
    const double den = red[0] + 1e-6;
    const double ca[3] = {red[1] / den, red[2] / den, red[3] / den}, cb[3] = {red[4] / den, red[5] / den, red[6] / den};
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int h = 0; h < 2; ++h)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) H[i * 3 + j] += wd[h] * (ax[h][i] - ca[i]) * (bx[h][j] - cb[j]);
#pragma unroll
    for (int q = 0; q < 9; ++q)
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) H[q] += __shfl_xor_sync(0xffffffffu, H[q], off);
    if (lane == 0) {
      double Hm[3][3], R[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Hm[i][j] = H[i * 3 + j];
      kabsch_rotation(Hm, R);
      float* out = a.b.seed_trans + ((size_t)p * a.smax + s) * 12;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out[i * 3 + j] = (float)R[i][j];
        out[9 + i] = (float)(cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// hypothesis scoring (PointDSC.py:325-332): inliers of every seed transform over all correspondences
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pdsc_fitness_kernel(Args a) {
  __shared__ int wsum[4];
  const int p = blockIdx.y, s = blockIdx.x;
  const PairMeta pm = a.meta[p];
  if (s >= pm.n_seeds) return;
  const float* T = a.b.seed_trans + ((size_t)p * a.smax + s) * 12;
  float R[9], t[3];
  for (int i = 0; i < 9; ++i) R[i] = T[i];
  for (int i = 0; i < 3; ++i) t[i] = T[9 + i];
  const float* Sp = a.src + (size_t)p * a.cap * 3;
  const float* Tp = a.tgt + (size_t)p * a.cap * 3;
  int cnt = 0;
  for (int i = threadIdx.x; i < pm.n; i += 128) {
    const float x = Sp[i * 3], y = Sp[i * 3 + 1], z = Sp[i * 3 + 2];
    const float dx = fmaf(R[2], z, fmaf(R[1], y, R[0] * x)) + t[0] - Tp[i * 3];
    const float dy = fmaf(R[5], z, fmaf(R[4], y, R[3] * x)) + t[1] - Tp[i * 3 + 1];
    const float dz = fmaf(R[8], z, fmaf(R[7], y, R[6] * x)) + t[2] - Tp[i * 3 + 2];
    cnt += sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))) < a.inlier_th;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) a.b.fit[(size_t)p * a.smax + s] = wsum[0] + wsum[1] + wsum[2] + wsum[3];
}

// ------------------------------------------------------------------------------------------------
// best hypothesis + post-refinement (PointDSC.py:333-335, :403-438); one CTA per pair
// ------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* scratch /*[16][NV]*/, double* result /*[NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NV; ++q)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  __syncthreads();  // scratch / result reuse
  if (lane == 0)
    for (int q = 0; q < NV; ++q) scratch[warp * NV + q] = v[q];
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += scratch[w * NV + threadIdx.x];
    result[threadIdx.x] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) pdsc_refine_kernel(Args a, float* out_T, float* out_initial, int32_t* out_best) {
  __shared__ double scratch[16 * 9];
  __shared__ double res[9];
  __shared__ float Tm[12];
  __shared__ int s_cnt[16];
  __shared__ int s_total;
  const int p = blockIdx.x;
  const PairMeta pm = a.meta[p];
  const int n = pm.n;
  if (threadIdx.x == 0) {
    int best = 0, bc = -1;
    for (int s = 0; s < pm.n_seeds; ++s) {
      const int c = a.b.fit[(size_t)p * a.smax + s];
      if (c > bc) bc = c, best = s;  // argmax keeps the first maximum
    }
    const float* T = a.b.seed_trans + ((size_t)p * a.smax + best) * 12;
    for (int i = 0; i < 12; ++i) Tm[i] = T[i];
    if (out_best) out_best[p] = best;
    if (out_initial) {
      float* o = out_initial + (size_t)p * 16;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) o[i * 4 + j] = T[i * 3 + j];
        o[i * 4 + 3] = T[9 + i];
      }
      o[12] = 0.f, o[13] = 0.f, o[14] = 0.f, o[15] = 1.f;
    }
  }
  __syncthreads();
  const float* Sp = a.src + (size_t)p * a.cap * 3;
  const float* Tp = a.tgt + (size_t)p * a.cap * 3;
  // PointDSC.py:415-418: the model's own inlier_threshold stays at the constructor default 0.10
  const float th = a.inlier_th == 0.10f ? 0.10f : 1.2f;
  int prev = 0;
  for (int it = 0; it < 20; ++it) {
    // this thread's points (n <= 4 * 512 is enforced by the host)
    float wgt[4];
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 512;
      wgt[q] = -1.f;
      if (i < n) {
        const float x = Sp[i * 3], y = Sp[i * 3 + 1], z = Sp[i * 3 + 2];
        const float dx = fmaf(Tm[2], z, fmaf(Tm[1], y, Tm[0] * x)) + Tm[9] - Tp[i * 3];
        const float dy = fmaf(Tm[5], z, fmaf(Tm[4], y, Tm[3] * x)) + Tm[10] - Tp[i * 3 + 1];
        const float dz = fmaf(Tm[8], z, fmaf(Tm[7], y, Tm[6] * x)) + Tm[11] - Tp[i * 3 + 2];
        const float l2 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        if (l2 < th) {
          const float r = __fdiv_rn(l2, th);
          wgt[q] = __fdiv_rn(1.f, __fadd_rn(1.f, __fmul_rn(r, r)));
          ++cnt;
        }
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 16; ++w) t += s_cnt[w];
      s_total = t;
    }
    __syncthreads();
    const int total = s_total;
    if (abs(total - prev) < 1) break;
    prev = total;
    double s7[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 512;
      if (wgt[q] >= 0.f) {
        const double w = wgt[q];
        s7[0] += w;
        for (int d = 0; d < 3; ++d) s7[1 + d] += w * Sp[i * 3 + d], s7[4 + d] += w * Tp[i * 3 + d];
      }
    }
    block_reduce<7>(s7, scratch, res);
    const double den = res[0] + 1e-6;
    const double ca[3] = {res[1] / den, res[2] / den, res[3] / den}, cb[3] = {res[4] / den, res[5] / den, res[6] / den};
    double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * 512;
      if (wgt[q] >= 0.f) {
        const double w = wgt[q];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) H[r * 3 + c] += w * ((double)Sp[i * 3 + r] - ca[r]) * ((double)Tp[i * 3 + c] - cb[c]);
      }
    }
    block_reduce<9>(H, scratch, res);
    if (threadIdx.x == 0) {
      double Hm[3][3], R[3][3];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Hm[i][j] = res[i * 3 + j];
      kabsch_rotation(Hm, R);
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Tm[i * 3 + j] = (float)R[i][j];
        Tm[9 + i] = (float)(cb[i] - (R[i][0] * ca[0] + R[i][1] * ca[1] + R[i][2] * ca[2]));
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float* o = out_T + (size_t)p * 16;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) o[i * 4 + j] = Tm[i * 3 + j];
      o[i * 4 + 3] = Tm[9 + i];
    }
    o[12] = 0.f, o[13] = 0.f, o[14] = 0.f, o[15] = 1.f;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static Model* model_of(oryon_handle* h) { return static_cast<Model*>(h->pointdsc_model); }

void destroy_model(oryon_handle* h) {
  Model* m = model_of(h);
  if (!m) return;
  m->blob.release();
  m->layer_table.release();
  m->tc_blob.release();
  delete m;
  h->pointdsc_model = nullptr;
}

static void destroy_model_ptr(Model* m) {
  m->blob.release(), m->layer_table.release(), m->tc_blob.release();
  delete m;
}

// Packed host weights, reference state_dict order and shapes (models/pointdsc/PointDSC.py:9-113):
//   encoder.layer0 {weight [C][in_dim], bias [C]}
//   per layer i:  PointCN_layer_i.0 {weight [C][C], bias}, PointCN_layer_i.1 {weight, bias, running_mean, running_var} [C]
//                 NonLocal_layer_i.fc_message.0 {w [C/2][C], b}, .1 BN [C/2] x4, .3 {w [C/2][C/2], b}, .4 BN x4, .6 {w [C][C/2], b}
//                 NonLocal_layer_i.projection_q {w [C][C], b}, projection_k, projection_v
//   classification.0 {w [32][C], b}, .2 {w [32][32], b}, .4 {w [1][32], b}
int load_weights(oryon_handle* h, const oryon_pointdsc_config* cfg, const float* w, int64_t n_floats, cudaStream_t st) {
  ORYON_REQUIRE(h && cfg && w, "oryon_pointdsc_load: null argument");
  ORYON_REQUIRE(cfg->num_channels == C, "oryon_pointdsc_load: num_channels=%d not supported (kernels are built for %d)", cfg->num_channels, C);
  ORYON_REQUIRE(cfg->in_dim == 6, "oryon_pointdsc_load: in_dim=%d not supported (6)", cfg->in_dim);
  ORYON_REQUIRE(cfg->num_layers >= 1 && cfg->num_layers <= 64, "oryon_pointdsc_load: num_layers=%d out of range", cfg->num_layers);
  ORYON_REQUIRE(cfg->k >= 1 && cfg->k <= kMaxK, "oryon_pointdsc_load: k=%d out of range (1..%d)", cfg->k, kMaxK);
  ORYON_REQUIRE(cfg->num_iterations >= 1, "oryon_pointdsc_load: num_iterations must be >= 1");
  const int L = cfg->num_layers, H = C / 2;
  const int64_t per_layer = (int64_t)(C * C + C) + 4 * C + (H * C + H) + 4 * H + (H * H + H) + 4 * H + (C * H + C) + 3 * (C * C + C);
  const int64_t expect = (int64_t)(C * 6 + C) + L * per_layer + (32 * C + 32) + (32 * 32 + 32) + (32 + 1);
  ORYON_REQUIRE(n_floats == expect, "oryon_pointdsc_load: got %lld floats, the configured model has %lld", (long long)n_floats,
                (long long)expect);
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  destroy_model(h);

  std::vector<float> packed;
  packed.reserve((size_t)expect);
  struct Off { size_t w, b; };
  const float* cur = w;
  const double eps = 1e-5;
  // conv [cout][cin] (+ optional BN) -> transposed [cin][cout] folded, 16-byte aligned segments
  auto fold = [&](int cout, int cin, bool bn) -> Off {
    const float* W = cur;
    const float* B = cur + (size_t)cout * cin;
    cur += (size_t)cout * cin + cout;
    const float *g = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr;
    if (bn) g = cur, beta = cur + cout, mean = cur + 2 * cout, var = cur + 3 * cout, cur += 4 * cout;
    while (packed.size() % 4) packed.push_back(0.f);
    Off o;
    o.w = packed.size();
    packed.resize(o.w + (size_t)cin * cout);
    for (int co = 0; co < cout; ++co) {
      const double sc = bn ? (double)g[co] / std::sqrt((double)var[co] + eps) : 1.0;
      for (int ci = 0; ci < cin; ++ci) packed[o.w + (size_t)ci * cout + co] = (float)(sc * W[(size_t)co * cin + ci]);
    }
    while (packed.size() % 4) packed.push_back(0.f);
    o.b = packed.size();
    for (int co = 0; co < cout; ++co) {
      const double sc = bn ? (double)g[co] / std::sqrt((double)var[co] + eps) : 1.0;
      packed.push_back((float)(bn ? ((double)B[co] - mean[co]) * sc + beta[co] : (double)B[co]));
    }
    return o;
  };
  // the same folded weights once more in GEMM order [cout][cin] (row-major, K = cin contiguous) for the tensor-core path
  std::vector<float> rowmajor;
  auto rm_of = [&](const Off& o, int cout, int cin) -> size_t {
    const size_t at = rowmajor.size();
    rowmajor.resize(at + (size_t)cout * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) rowmajor[at + (size_t)co * cin + ci] = packed[o.w + (size_t)ci * cout + co];
    return at;
  };
  struct LayerOff { Off pcn, m0, m1, m2, q, k, v; size_t r_pcn, r_qk, r_v, r_m0, r_m1, r_m2, b_qk; };
  const Off l0 = fold(C, 6, false);
  std::vector<LayerOff> lo_(L);
  for (int i = 0; i < L; ++i) {
    lo_[i].pcn = fold(C, C, true);
    lo_[i].m0 = fold(H, C, true);
    lo_[i].m1 = fold(H, H, true);
    lo_[i].m2 = fold(C, H, false);
    lo_[i].q = fold(C, C, false);
    lo_[i].k = fold(C, C, false);
    lo_[i].v = fold(C, C, false);
    lo_[i].r_pcn = rm_of(lo_[i].pcn, C, C);
    lo_[i].r_qk = rm_of(lo_[i].q, C, C), rm_of(lo_[i].k, C, C);      // q rows then k rows: one [2C][C] operand
    lo_[i].r_v = rm_of(lo_[i].v, C, C);
    lo_[i].r_m0 = rm_of(lo_[i].m0, H, C), lo_[i].r_m1 = rm_of(lo_[i].m1, H, H), lo_[i].r_m2 = rm_of(lo_[i].m2, C, H);
    while (packed.size() % 4) packed.push_back(0.f);
    lo_[i].b_qk = packed.size();                                    // q bias followed by k bias
    for (int c = 0; c < C; ++c) packed.push_back(packed[lo_[i].q.b + c]);
    for (int c = 0; c < C; ++c) packed.push_back(packed[lo_[i].k.b + c]);
  }
  const Off c0 = fold(32, C, false), c1 = fold(32, 32, false), c2 = fold(1, 32, false);
  if (cur != w + n_floats) {
    set_error("oryon_pointdsc_load: internal packing mismatch");
    return ORYON_ERR_INVALID_ARGUMENT;
  }

  Model* m = new Model();
  m->cfg = *cfg;
  int rc;
  if ((rc = m->blob.reserve(packed.size() * sizeof(float), st)) || (rc = m->layer_table.reserve(sizeof(LayerW) * L, st))) {
    m->blob.release(), m->layer_table.release();
    delete m;
    return rc;
  }
  ORYON_CUDA_CHECK(cudaMemcpyAsync(m->blob.ptr, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  const float* base = m->blob.as<float>();
  m->l0_w = base + l0.w, m->l0_b = base + l0.b;
  m->layers.resize(L);
  for (int i = 0; i < L; ++i) {
    LayerW& lw = m->layers[i];
    lw.pcn_w = base + lo_[i].pcn.w, lw.pcn_b = base + lo_[i].pcn.b;
    lw.q_w = base + lo_[i].q.w, lw.q_b = base + lo_[i].q.b;
    lw.k_w = base + lo_[i].k.w, lw.k_b = base + lo_[i].k.b;
    lw.v_w = base + lo_[i].v.w, lw.v_b = base + lo_[i].v.b;
    lw.m0_w = base + lo_[i].m0.w, lw.m0_b = base + lo_[i].m0.b;
    lw.m1_w = base + lo_[i].m1.w, lw.m1_b = base + lo_[i].m1.b;
    lw.m2_w = base + lo_[i].m2.w, lw.m2_b = base + lo_[i].m2.b;
  }
  m->c0_w = base + c0.w, m->c0_b = base + c0.b, m->c1_w = base + c1.w, m->c1_b = base + c1.b, m->c2_w = base + c2.w, m->c2_b = base + c2.b;
  ORYON_CUDA_CHECK(cudaMemcpyAsync(m->layer_table.ptr, m->layers.data(), sizeof(LayerW) * L, cudaMemcpyHostToDevice, st));
  {  // tensor-core operands: fp32 row-major staging (front of the blob) -> fp16 split pairs (every K here is 128 or 64: no padding)
    const size_t nrm = rowmajor.size();
    if ((rc = m->tc_blob.reserve(nrm * sizeof(float) + 2 * nrm * sizeof(__half), st))) {
      destroy_model_ptr(m);
      return rc;
    }
    float* stage = m->tc_blob.as<float>();
    __half* hi = reinterpret_cast<__half*>(stage + nrm);
    __half* lo = hi + nrm;
    ORYON_CUDA_CHECK(cudaMemcpyAsync(stage, rowmajor.data(), nrm * sizeof(float), cudaMemcpyHostToDevice, st));
    if ((rc = gemm::split_rows(h, stage, (int64_t)nrm, 1, (int)nrm, hi, lo, (int64_t)nrm, st))) {
      destroy_model_ptr(m);
      return rc;
    }
    m->tc.resize(L);
    for (int i = 0; i < L; ++i) {
      Model::TcLayer& t = m->tc[i];
      t.pcn_hi = hi + lo_[i].r_pcn, t.pcn_lo = lo + lo_[i].r_pcn, t.qk_hi = hi + lo_[i].r_qk, t.qk_lo = lo + lo_[i].r_qk;
      t.v_hi = hi + lo_[i].r_v, t.v_lo = lo + lo_[i].r_v, t.m0_hi = hi + lo_[i].r_m0, t.m0_lo = lo + lo_[i].r_m0;
      t.m1_hi = hi + lo_[i].r_m1, t.m1_lo = lo + lo_[i].r_m1, t.m2_hi = hi + lo_[i].r_m2, t.m2_lo = lo + lo_[i].r_m2;
      t.pcn_b = base + lo_[i].pcn.b, t.qk_b = base + lo_[i].b_qk, t.v_b = base + lo_[i].v.b;
      t.m0_b = base + lo_[i].m0.b, t.m1_b = base + lo_[i].m1.b, t.m2_b = base + lo_[i].m2.b;
    }
  }
  ORYON_CUDA_CHECK(cudaStreamSynchronize(st));  // `packed`, `rowmajor` and `m->layers` are read by the copies
  h->pointdsc_model = m;
  return ORYON_OK;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------
// The NonLocal network on the tcgen05 GEMM (gemm.cuh), batched over the P pairs of the call.
// ------------------------------------------------------------------------------------------------
// Every per-point layer of the network is a 1x1 convolution = a linear layer over the P * npad points of the call, and the two
// products of a NonLocal block are per-pair GEMMs (M = N = npad points, K = C / K = npad): per layer
//   QK    [P*R][2C]   = X Wqk^T + b            (split pair out: Q | K, the operands of the score product)
//   V^T   [C][P*R]    = (X Wv^T + b)^T         (split pair out, transposed: the W operand of the message product)
//   S     [P][R][R]   = Q K^T / sqrt(C)        (fp32 out, batched over pairs)
//   W     = softmax(SC * S) over the keys      (pdsc_softmax_kernel: split pair out, K extent zero padded)
//   MSG   [P*R][C]    = W V                    (batched over pairs)
//   fc_message: 128 -> 64 (ReLU) -> 64 (ReLU) -> 128, + F1  (feat, fp32 and split pair)
//   PointCN of the next layer: 128 -> 128 (ReLU)            (F1 of the next layer, fp32 and split pair)
// with R = npad.  Operands are fp16 split pairs and every product is issued three times (float32-equivalent, gemm.cuh), like the
// network GEMMs; rows of points >= n are computed on finite garbage and never read.  Opt-in (ORYON_PDSC_TC=1), see run_pose.
struct TcBuffers {
  __half *xa_hi, *xa_lo, *xb_hi, *xb_lo;     // [P*R][C]   activations (A operands), ping-pong
  __half *qk_hi, *qk_lo;                     // [P*R][2C]
  __half *vt_hi, *vt_lo;                     // [C][ldv]   V transposed over ALL points of the call (+ 64 zero columns)
  float* S;                                  // [P][R][R]
  __half *pm_hi, *pm_lo;                     // [P][R][kp] softmax weights
  __half *msg_hi, *msg_lo;                   // [P*R][C]
  __half *h1_hi, *h1_lo, *h2_hi, *h2_lo;     // [P*R][C/2]
  float* feat;                               // [P*R][C]
  int R, kp, ldv;
};

static int run_network_tc(oryon_handle* h, Model* m, const Args& a, const TcBuffers& b, int P, cudaStream_t st) {
  const int R = b.R, H2 = C / 2, rows = P * R;
  const int L = m->cfg.num_layers;
  int rc;
  auto lin = [&](const __half* a_hi, const __half* a_lo, int K, const __half* w_hi, const __half* w_lo, int N, const gemm::Epilogue& ep) -> int {
    gemm::Problem p;
    p.M = rows, p.N = N, p.K = K, p.precision = 3;
    p.A.hi = a_hi, p.A.lo = a_lo, p.A.ld = K;
    p.W.hi = w_hi, p.W.lo = w_lo, p.W.ld = K;
    p.ep = ep;
    return gemm::launch(h, p, st);
  };
  auto ep_split = [](__half* hi, __half* lo, int ld, const float* bias, int act) {
    gemm::Epilogue e;
    e.out_hi = hi, e.out_lo = lo, e.ldh = ld, e.bias = bias, e.act = act;
    return e;
  };
  // layer 0's F1 comes from the prologue kernel in fp32
  if ((rc = gemm::split_rows(h, a.b.f1[0], C, rows, C, b.xa_hi, b.xa_lo, C, st))) return rc;
  const float* F1 = a.b.f1[0];
  __half *x_hi = b.xa_hi, *x_lo = b.xa_lo, *y_hi = b.xb_hi, *y_lo = b.xb_lo;
  for (int l = 0; l < L; ++l) {
    const Model::TcLayer& t = m->tc[l];
    if ((rc = lin(x_hi, x_lo, C, t.qk_hi, t.qk_lo, 2 * C, ep_split(b.qk_hi, b.qk_lo, 2 * C, t.qk_b, gemm::ACT_NONE)))) return rc;
    {
      gemm::Epilogue e = ep_split(b.vt_hi, b.vt_lo, b.ldv, t.v_b, gemm::ACT_NONE);
      e.transpose_h = 1;                      // element (point, channel) -> vt[channel][point]
      if ((rc = lin(x_hi, x_lo, C, t.v_hi, t.v_lo, C, e))) return rc;
    }
    {  // scores, batched over pairs: A = Q (columns 0..C-1 of QK), W = K (columns C..2C-1)
      gemm::Problem p;
      p.M = R, p.N = R, p.K = C, p.nb0 = P, p.precision = 3;
      p.A.hi = b.qk_hi, p.A.lo = b.qk_lo, p.A.ld = 2 * C, p.A.stride_b0 = (int64_t)R * 2 * C;
      p.W.hi = b.qk_hi + C, p.W.lo = b.qk_lo + C, p.W.ld = 2 * C, p.W.stride_b0 = (int64_t)R * 2 * C;
      p.ep.alpha = 1.f / 11.313708498984761f;  // / (num_channels // head) ** 0.5
      p.ep.out32 = b.S, p.ep.ld32 = R, p.ep.out_b0 = (int64_t)R * R;
      if ((rc = gemm::launch(h, p, st))) return rc;
    }
    h->span_begin(KID_PDSC_NET, st);
    if (b.kp <= 512) pdsc_softmax_kernel<16><<<dim3((R + 7) / 8, P), 256, 0, st>>>(a, b.S, b.pm_hi, b.pm_lo, b.kp);
    else pdsc_softmax_kernel<64><<<dim3((R + 7) / 8, P), 256, 0, st>>>(a, b.S, b.pm_hi, b.pm_lo, b.kp);
    h->span_end(st);
    ORYON_CUDA_CHECK(cudaGetLastError());
    {  // message = W V, batched over pairs: W operand = this pair's columns of V^T
      gemm::Problem p;
      p.M = R, p.N = C, p.K = b.kp, p.nb0 = P, p.precision = 3;
      p.A.hi = b.pm_hi, p.A.lo = b.pm_lo, p.A.ld = b.kp, p.A.stride_b0 = (int64_t)R * b.kp;
      p.W.hi = b.vt_hi, p.W.lo = b.vt_lo, p.W.ld = b.ldv, p.W.stride_b0 = R;
      p.ep.out_hi = b.msg_hi, p.ep.out_lo = b.msg_lo, p.ep.ldh = C, p.ep.outh_b0 = (int64_t)R * C;
      if ((rc = gemm::launch(h, p, st))) return rc;
    }
    if ((rc = lin(b.msg_hi, b.msg_lo, C, t.m0_hi, t.m0_lo, H2, ep_split(b.h1_hi, b.h1_lo, H2, t.m0_b, gemm::ACT_RELU)))) return rc;
    if ((rc = lin(b.h1_hi, b.h1_lo, H2, t.m1_hi, t.m1_lo, H2, ep_split(b.h2_hi, b.h2_lo, H2, t.m1_b, gemm::ACT_RELU)))) return rc;
    {  // feat = F1 + fc_message(message)
      gemm::Epilogue e = ep_split(y_hi, y_lo, C, t.m2_b, gemm::ACT_NONE);
      e.residual = F1, e.out32 = b.feat, e.ld32 = C;
      if ((rc = lin(b.h2_hi, b.h2_lo, H2, t.m2_hi, t.m2_lo, C, e))) return rc;
    }
    if (l + 1 < L) {  // PointCN of the next layer: its F1 (fp32 for the residual, split pair for the projections)
      float* f1n = a.b.f1[(l + 1) & 1];
      gemm::Epilogue e = ep_split(x_hi, x_lo, C, m->tc[l + 1].pcn_b, gemm::ACT_RELU);
      e.out32 = f1n, e.ld32 = C;
      if ((rc = lin(y_hi, y_lo, C, m->tc[l + 1].pcn_hi, m->tc[l + 1].pcn_lo, C, e))) return rc;
      F1 = f1n;
    }
  }
  const int tiles = (a.npad + TP - 1) / TP;
  h->span_begin(KID_PDSC_NET, st);
  pdsc_final_kernel<<<dim3(tiles, P), NT, 0, st>>>(a, b.feat);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

int run_pose(oryon_handle* h, const float* src, const float* tgt, const int32_t* n_host, int P, int cap, float* out_T,
             const oryon_pointdsc_debug* dbg, cudaStream_t st) {
  ORYON_REQUIRE(h && src && tgt && n_host && out_T, "oryon_pointdsc_pose: null argument");
  Model* m = model_of(h);
  if (!m) {
    set_error("oryon_pointdsc_pose: no PointDSC weights loaded (oryon_pointdsc_load)");
    return ORYON_ERR_NOT_LOADED;
  }
  ORYON_REQUIRE(P > 0 && cap > 0 && cap <= 2048, "oryon_pointdsc_pose: P=%d cap=%d out of range (cap <= 2048)", P, cap);
  const oryon_pointdsc_config& cfg = m->cfg;
  std::vector<PairMeta> meta(P);
  int smax = 1, nmax = 0;
  for (int p = 0; p < P; ++p) {
    const int n = n_host[p];
    ORYON_REQUIRE(n >= 2 && n <= cap, "oryon_pointdsc_pose: pair %d has %d correspondences (need 2..cap)", p, n);
    meta[p].n = n;
    meta[p].n_seeds = (int)((double)n * cfg.ratio);  // int(num_corr * self.ratio), PointDSC.py:174
    // the reference's argmax over an empty seed dimension raises (PointDSC.py:333)
    ORYON_REQUIRE(meta[p].n_seeds >= 1, "oryon_pointdsc_pose: pair %d: int(%d * ratio) == 0 seeds (the reference raises here)", p, n);
    meta[p].k = std::min(cfg.k, n - 1);
    smax = std::max(smax, meta[p].n_seeds);
    nmax = std::max(nmax, n);
  }
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  const int npad = round_up(cap, TP);
  int rc;
  // workspace carve-up
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  };
  const size_t feat = (size_t)P * npad * C * 4;
  const size_t o_sc = take((size_t)P * npad * npad * 4);
  size_t o_f1[2], o_q[2], o_k[2], o_v[2];
  for (int g = 0; g < 2; ++g) o_f1[g] = take(feat), o_q[g] = take(feat), o_k[g] = take(feat), o_v[g] = take(feat);
  const size_t o_fn = take(feat), o_conf = take((size_t)P * npad * 4), o_seeds = take((size_t)P * smax * 4);
  const size_t o_knn = take((size_t)P * smax * kMaxK * 4), o_M = take((size_t)P * smax * kMaxK * kMaxK * 4);
  const size_t o_st = take((size_t)P * smax * 12 * 4), o_fit = take((size_t)P * smax * 4), o_meta = take(sizeof(PairMeta) * P);
  // Network path: the fp32 CUDA-core layer kernel (default), or -- ORYON_PDSC_TC=1, read per call -- the tcgen05 GEMM form.
  // Measured on B200, 32 pairs x 500 correspondences (profiles/r02_pointdsc_tc.md): 4.45 ms against 4.93 ms, and on the hardest
  // parity case (333 correspondences, 50 % outliers) a pose entry 1.04e-4 from the reference where the fp32 kernel is within 1e-4:
  // the tensor core accumulates in TRUNCATED fp32 (~1e-5 relative at K = 512), ten times the error of fp32 FMAs, and twelve layers
  // of it reach the pose.  Half a millisecond does not buy that; the fp32 kernel stays the default.
  const char* tc_env = std::getenv("ORYON_PDSC_TC");
  const bool use_tc = tc_env && tc_env[0] == '1';
  TcBuffers tb{};
  size_t o_tc[16] = {0};
  tb.R = npad, tb.kp = round_up(npad, 64), tb.ldv = round_up(P * npad + 64, 64);
  if (use_tc) {
    const size_t rows = (size_t)P * npad;
    o_tc[0] = take(rows * C * 2), o_tc[1] = take(rows * C * 2), o_tc[2] = take(rows * C * 2), o_tc[3] = take(rows * C * 2);   // xa, xb
    o_tc[4] = take(rows * 2 * C * 2), o_tc[5] = take(rows * 2 * C * 2);                                                       // qk
    o_tc[6] = take((size_t)C * tb.ldv * 2), o_tc[7] = take((size_t)C * tb.ldv * 2);                                           // vt
    o_tc[8] = take(rows * npad * 4);                                                                                          // S
    o_tc[9] = take(rows * tb.kp * 2), o_tc[10] = take(rows * tb.kp * 2);                                                      // pm
    o_tc[11] = take(rows * C * 2), o_tc[12] = take(rows * C * 2);                                                             // msg
    o_tc[13] = take(rows * C * 2), o_tc[14] = take(rows * C * 2);                                                             // h1 | h2 (hi and lo halves inside)
    o_tc[15] = take(rows * C * 4);                                                                                            // feat
  }
  if ((rc = h->pdsc_ws.reserve(off, st))) return rc;
  char* ws = h->pdsc_ws.as<char>();
  if (use_tc) {
    auto hp = [&](int i) { return reinterpret_cast<__half*>(ws + o_tc[i]); };
    const size_t rows = (size_t)P * npad;
    tb.xa_hi = hp(0), tb.xa_lo = hp(1), tb.xb_hi = hp(2), tb.xb_lo = hp(3), tb.qk_hi = hp(4), tb.qk_lo = hp(5), tb.vt_hi = hp(6), tb.vt_lo = hp(7);
    tb.S = reinterpret_cast<float*>(ws + o_tc[8]);
    tb.pm_hi = hp(9), tb.pm_lo = hp(10), tb.msg_hi = hp(11), tb.msg_lo = hp(12);
    tb.h1_hi = hp(13), tb.h1_lo = hp(13) + rows * (C / 2), tb.h2_hi = hp(14), tb.h2_lo = hp(14) + rows * (C / 2);
    tb.feat = reinterpret_cast<float*>(ws + o_tc[15]);
    // the grow-only workspace may hold anything from an earlier call's layout: the K padding of V^T (points past the last pair)
    // must be finite, and rows the prologue does not write (points >= n of every pair) feed GEMM rows that must stay finite too
    ORYON_CUDA_CHECK(cudaMemsetAsync(tb.vt_hi, 0, (size_t)C * tb.ldv * 2, st));
    ORYON_CUDA_CHECK(cudaMemsetAsync(tb.vt_lo, 0, (size_t)C * tb.ldv * 2, st));
    ORYON_CUDA_CHECK(cudaMemsetAsync(ws + o_f1[0], 0, feat, st));
  }
  ORYON_CUDA_CHECK(cudaMemcpyAsync(ws + o_meta, meta.data(), sizeof(PairMeta) * P, cudaMemcpyHostToDevice, st));

  Args a;
  a.src = src, a.tgt = tgt;
  a.meta = reinterpret_cast<const PairMeta*>(ws + o_meta);
  a.cap = cap, a.npad = npad, a.smax = smax, a.in_dim = cfg.in_dim;
  a.b.sc = reinterpret_cast<float*>(ws + o_sc);
  for (int g = 0; g < 2; ++g) {
    a.b.f1[g] = reinterpret_cast<float*>(ws + o_f1[g]), a.b.q[g] = reinterpret_cast<float*>(ws + o_q[g]);
    a.b.kt[g] = reinterpret_cast<float*>(ws + o_k[g]), a.b.v[g] = reinterpret_cast<float*>(ws + o_v[g]);
  }
  a.b.fn = reinterpret_cast<float*>(ws + o_fn), a.b.conf = reinterpret_cast<float*>(ws + o_conf);
  a.b.seeds = reinterpret_cast<int32_t*>(ws + o_seeds), a.b.knn = reinterpret_cast<int32_t*>(ws + o_knn);
  a.b.M = reinterpret_cast<float*>(ws + o_M), a.b.seed_trans = reinterpret_cast<float*>(ws + o_st);
  a.b.fit = reinterpret_cast<int32_t*>(ws + o_fit);
  a.layers = m->layer_table.as<LayerW>();
  a.l0_w = m->l0_w, a.l0_b = m->l0_b, a.c0_w = m->c0_w, a.c0_b = m->c0_b, a.c1_w = m->c1_w, a.c1_b = m->c1_b, a.c2_w = m->c2_w, a.c2_b = m->c2_b;
  // sigma_spat ** 2 and sigma ** 2 are float32 tensor powers in the reference
  const float sd32 = (float)cfg.sigma_d, s32 = (float)cfg.sigma;
  a.sigma_d2 = sd32 * sd32, a.sigma2 = s32 * s32;
  a.nms_radius = (float)cfg.nms_radius, a.inlier_th = (float)cfg.inlier_threshold;
  a.num_iterations = cfg.num_iterations;

  const int tiles = (nmax + TP - 1) / TP;
  h->span_begin(KID_PDSC_SC, st);
  pdsc_sc_kernel<<<dim3((npad + 31) / 32, (npad + 31) / 32, P), 256, 0, st>>>(a);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  h->span_begin(KID_PDSC_NET, st);
  pdsc_prologue_kernel<<<dim3(tiles, P), NT, 0, st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  const size_t layer_smem = (size_t)(3 * C * TPS + npad * TPS) * sizeof(float);
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(pdsc_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)layer_smem));
  if (use_tc) {
    h->span_end(st);                                        // the prologue's span; the GEMMs open their own, booked as pointdsc_net
    h->span_alias = KID_PDSC_NET, h->gemm_uncounted = true;
    rc = run_network_tc(h, m, a, tb, P, st);
    h->span_alias = -1, h->gemm_uncounted = false;
    if (rc) return rc;
    h->span_begin(KID_PDSC_NET, st);                        // (empty) span closed below
  } else {
    for (int l = 0; l < cfg.num_layers; ++l) {
      pdsc_layer_kernel<<<dim3(tiles, P), NT, layer_smem, st>>>(a, l, l == cfg.num_layers - 1 ? 1 : 0);
      ORYON_CUDA_CHECK(cudaGetLastError());
    }
  }
  h->span_end(st);
  h->span_begin(KID_PDSC_SEEDS, st);
  pdsc_seeds_kernel<<<P, 1024, (size_t)npad * 5 * sizeof(float), st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  pdsc_knn_kernel<<<dim3(smax, P), 256, (size_t)npad * sizeof(float), st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  const size_t compat_smem = (size_t)(kMaxK * (C + 1) + kMaxK * 6) * sizeof(float);
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(pdsc_compat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)compat_smem));
  pdsc_compat_kernel<<<dim3(smax, P), 256, compat_smem, st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  const size_t power_smem = (size_t)2 * smax * kMaxK * sizeof(float);
  ORYON_CUDA_CHECK(cudaFuncSetAttribute(pdsc_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)power_smem));
  pdsc_power_kernel<<<P, 512, power_smem, st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  pdsc_fitness_kernel<<<dim3(smax, P), 128, 0, st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  h->span_end(st);
  h->span_begin(KID_PDSC_REFINE, st);
  pdsc_refine_kernel<<<P, 512, 0, st>>>(a, out_T, dbg ? dbg->initial_trans : nullptr, dbg ? dbg->best_seed : nullptr);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());

  if (dbg) {  // copies for the parity tests (device -> caller's device buffers)
    if (dbg->conf)
      ORYON_CUDA_CHECK(cudaMemcpy2DAsync(dbg->conf, (size_t)cap * 4, a.b.conf, (size_t)npad * 4, (size_t)cap * 4, P, cudaMemcpyDeviceToDevice, st));
    if (dbg->seeds) {
      ORYON_REQUIRE(dbg->seeds_cap >= smax, "oryon_pointdsc_pose: debug seeds_cap %d < %d", dbg->seeds_cap, smax);
      ORYON_CUDA_CHECK(cudaMemcpy2DAsync(dbg->seeds, (size_t)dbg->seeds_cap * 4, a.b.seeds, (size_t)smax * 4, (size_t)smax * 4, P,
                                         cudaMemcpyDeviceToDevice, st));
    }
    if (dbg->fitness) {
      ORYON_REQUIRE(dbg->seeds_cap >= smax, "oryon_pointdsc_pose: debug seeds_cap %d < %d", dbg->seeds_cap, smax);
      ORYON_CUDA_CHECK(cudaMemcpy2DAsync(dbg->fitness, (size_t)dbg->seeds_cap * 4, a.b.fit, (size_t)smax * 4, (size_t)smax * 4, P,
                                         cudaMemcpyDeviceToDevice, st));
    }
    if (dbg->features)
      ORYON_CUDA_CHECK(cudaMemcpy2DAsync(dbg->features, (size_t)cap * C * 4, a.b.fn, (size_t)npad * C * 4, (size_t)cap * C * 4, P,
                                         cudaMemcpyDeviceToDevice, st));
  }
  return ORYON_OK;
}

}  // namespace pdsc
}  // namespace oryon
