// N2 -- pose-error metrics of Evaluator.register_eval (reference utils/evaluator.py:206-288), for P pose pairs per call:
//   R / T error          utils/metrics.py:236-259 compute_RT_distances (float64 here; the reference normalises the float32
//                        prediction in float32, which shows up as ~1e-7 / sin(theta) rad in its angle)
//   ADD                  utils/metrics.py:205-218 + utils/pcd.py:127-133: the model is transformed in FLOAT16 (operands
//                        rounded to half, float32 accumulation over k, result rounded to half, translation added in half);
//                        difference, squares, their sum and the square root are rounded to half, the mean is a float32
//                        sum rounded to half.  Emulated operation by operation.
//   ADD-S                utils/metrics.py:220-234: float64 nearest-neighbour distances between the two float16 clouds
//                        (the reference asks a KDTree; brute force over the model here), float64 mean.
//   MSSD / MSPD          bop_toolkit_lib/pose_error.py:370-426 as called from evaluator.py:258-266: poses rounded to half,
//                        translation to half millimetres, float64 arithmetic, over the FIRST THREE model points only
//                        (np_transform's pts[:, :3] slices the point axis of a [1,N,3] array, pose_error.py:349).
// One CTA column per pose (grid.y), kChunks CTAs share the model points of a pose; partial sums are combined in a fixed
// order by the last CTA to finish, so results are deterministic.
#include <map>

#include "common.cuh"

namespace oryon {
namespace eval {

constexpr int kThreads = 256;
constexpr int kChunks = 8;
constexpr int kPosesPerLaunch = 64;
constexpr int kTile = 1024;   // ground-truth points staged in shared memory per sweep

struct Object {
  double* pts = nullptr;    // [n][3] mm
  double* syms = nullptr;   // [s][12] rows of [R | t]
  int32_t* faces = nullptr; // [f][3] vertex indices (VSD only)
  int n = 0, s = 0, f = 0;
};

struct State {
  std::map<int, Object> objects;
  DeviceBuffer ws;          // partial sums [P][kChunks] doubles + counters [P] ints
  DeviceBuffer vsd_ws;      // projected vertices, depth buffers and counters of the VSD pass
};

struct PoseObj {
  const double* pts;
  const double* syms;
  int n, s;
};

struct Args {
  const double* pred;   // [P][16]
  const double* gt;     // [P][16]
  const double* cams;   // [P][9]
  double* out;          // [P][6]
  double* partial;      // [P][kChunks]
  unsigned* counters;   // [P]
  int p0;
  PoseObj obj[kPosesPerLaunch];
};

__device__ __forceinline__ float h2f(double x) { return __half2float(__double2half(x)); }   // round a float64 to half, widen
__device__ __forceinline__ float rh(float x) { return __half2float(__float2half_rn(x)); }    // round a float32 to half, widen

struct Pose16 {
  float r[9], t[3];
};

__device__ __forceinline__ Pose16 load_pose16(const double* T) {
  Pose16 p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) p.r[3 * i + j] = h2f(T[4 * i + j]);
    p.t[i] = h2f(T[4 * i + 3]);
  }
  return p;
}

// np.dot(pcd.astype(f16), r.astype(f16).T) + t.astype(f16): products of halves are exact in float32; float32 adds over k.
__device__ __forceinline__ void transform16(const Pose16& p, const double* pt_mm, float out[3]) {
  const float x = h2f(pt_mm[0] / 1000.0), y = h2f(pt_mm[1] / 1000.0), z = h2f(pt_mm[2] / 1000.0);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float acc = __fmul_rn(x, p.r[3 * j]);
    acc = __fadd_rn(acc, __fmul_rn(y, p.r[3 * j + 1]));
    acc = __fadd_rn(acc, __fmul_rn(z, p.r[3 * j + 2]));
    out[j] = rh(__fadd_rn(rh(acc), p.t[j]));
  }
}

__device__ double block_sum(double v, double* red) {
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < kThreads / 32; ++w) t += red[w];   // fixed order
  return t;
}

__device__ double block_min(double v, double* red) {
  for (int off = 16; off >= 1; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[0];
  for (int w = 1; w < kThreads / 32; ++w) t = fmin(t, red[w]);
  return t;
}

__global__ void __launch_bounds__(kThreads) pose_errors_kernel(const __grid_constant__ Args a) {
  __shared__ float gts[kTile][3];
  __shared__ double red[kThreads / 32];
  __shared__ bool last;
  const int lp = blockIdx.y, p = a.p0 + lp, chunk = blockIdx.x;
  const PoseObj o = a.obj[lp];
  const double* Tp = a.pred + 16 * (size_t)p;
  const double* Tg = a.gt + 16 * (size_t)p;
  double* out = a.out + 6 * (size_t)p;
  const bool symmetric = o.s > 1;
  const Pose16 pp = load_pose16(Tp), pg = load_pose16(Tg);

  // ---- ADD / ADD-S partial sum over this chunk's predicted points ----
  const int per = (o.n + kChunks - 1) / kChunks;
  const int i0 = chunk * per, i1 = min(o.n, i0 + per);
  double acc = 0.0;
  if (!symmetric) {
    for (int i = i0 + threadIdx.x; i < i1; i += kThreads) {
      float a3[3], b3[3];
      transform16(pp, o.pts + 3 * (size_t)i, a3);
      transform16(pg, o.pts + 3 * (size_t)i, b3);
      float sq[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float d = rh(__fsub_rn(a3[j], b3[j]));
        sq[j] = rh(__fmul_rn(d, d));
      }
      const float s = rh(__fadd_rn(sq[0], __fadd_rn(sq[1], sq[2])));   // numpy's half add.reduce: first + float32 sum of the rest
      acc += (double)rh(__fsqrt_rn(s));
    }
  } else {
    // every thread owns up to kOwn predicted points; the ground-truth cloud streams through shared memory
    constexpr int kOwn = 4;
    for (int base = i0; base < i1; base += kThreads * kOwn) {
      float q[kOwn][3];
      double best[kOwn];
#pragma unroll
      for (int u = 0; u < kOwn; ++u) {
        const int i = base + u * kThreads + threadIdx.x;
        best[u] = 1e300;
        if (i < i1) transform16(pp, o.pts + 3 * (size_t)i, q[u]);
        else q[u][0] = q[u][1] = q[u][2] = 0.f;
      }
      for (int g0 = 0; g0 < o.n; g0 += kTile) {
        __syncthreads();
        for (int g = threadIdx.x; g < min(kTile, o.n - g0); g += kThreads) transform16(pg, o.pts + 3 * (size_t)(g0 + g), gts[g]);
        __syncthreads();
        const int ng = min(kTile, o.n - g0);
        for (int g = 0; g < ng; ++g) {
          const float gx = gts[g][0], gy = gts[g][1], gz = gts[g][2];
#pragma unroll
          for (int u = 0; u < kOwn; ++u) {
            // differences of halves are exact in float32; squares and their sum in float64
            const double dx = (double)(q[u][0] - gx), dy = (double)(q[u][1] - gy), dz = (double)(q[u][2] - gz);
            best[u] = fmin(best[u], dx * dx + dy * dy + dz * dz);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kOwn; ++u)
        if (base + u * kThreads + threadIdx.x < i1) acc += sqrt(best[u]);
    }
  }
  const double part = block_sum(acc, red);
  if (threadIdx.x == 0) a.partial[(size_t)p * kChunks + chunk] = part;

  // ---- chunk 0: R / T error, MSSD, MSPD ----
  if (chunk == 0) {
    if (threadIdx.x == 0) {
      double R1[9], R2[9];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R1[3 * i + j] = Tp[4 * i + j], R2[3 * i + j] = Tg[4 * i + j];
      auto det3 = [](const double* m) {
        return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
      };
      const double c1 = cbrt(det3(R1)), c2 = cbrt(det3(R2));
      double tr = 0.0;
      for (int i = 0; i < 9; ++i) tr += (R1[i] / c1) * (R2[i] / c2);   // trace(R1 R2^T)
      double arg = (tr - 1.0) / 2.0;
      arg = fmin(fmax(arg, -1.0 + 1e-12), 1.0 - 1e-12);
      double theta = acos(arg) * 180.0 / 3.14159265358979323846;
      if (isnan(theta)) theta = 180.0;
      const double dx = Tp[3] - Tg[3], dy = Tp[7] - Tg[7], dz = Tp[11] - Tg[11];
      out[0] = theta;
      out[1] = sqrt(dx * dx + dy * dy + dz * dz) * 100.0;
      out[3] = symmetric ? 1.0 : 0.0;
    }
    // poses in half, translations in half millimetres (evaluator.py:258-261), then float64
    double Rp[9], tp[3], Rg[9], tg[3];
    for (int i = 0; i < 9; ++i) Rp[i] = (double)pp.r[i], Rg[i] = (double)pg.r[i];
    for (int i = 0; i < 3; ++i) tp[i] = (double)rh(__fmul_rn(pp.t[i], 1000.f)), tg[i] = (double)rh(__fmul_rn(pg.t[i], 1000.f));
    const double* K = a.cams + 9 * (size_t)p;
    const int m = min(3, o.n);
    double est[3][3], pe[3][2];
    for (int i = 0; i < m; ++i) {
      const double* x = o.pts + 3 * (size_t)i;
      for (int j = 0; j < 3; ++j) est[i][j] = x[0] * Rp[3 * j] + x[1] * Rp[3 * j + 1] + x[2] * Rp[3 * j + 2] + tp[j];
      double pr[3];
      for (int j = 0; j < 3; ++j) pr[j] = est[i][0] * K[3 * j] + est[i][1] * K[3 * j + 1] + est[i][2] * K[3 * j + 2];
      pe[i][0] = pr[0] / pr[2], pe[i][1] = pr[1] / pr[2];
    }
    double best_s = 1e300, best_p = 1e300;
    for (int s = threadIdx.x; s < o.s; s += kThreads) {
      const double* S = o.syms + 12 * (size_t)s;
      double Rs[9], ts[3];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rs[3 * i + j] = Rg[3 * i] * S[j] + Rg[3 * i + 1] * S[4 + j] + Rg[3 * i + 2] * S[8 + j];
        ts[i] = Rg[3 * i] * S[3] + Rg[3 * i + 1] * S[7] + Rg[3 * i + 2] * S[11] + tg[i];
      }
      double es = 0.0, ep = 0.0;
      for (int i = 0; i < m; ++i) {
        const double* x = o.pts + 3 * (size_t)i;
        double q[3], pr[3];
        for (int j = 0; j < 3; ++j) q[j] = x[0] * Rs[3 * j] + x[1] * Rs[3 * j + 1] + x[2] * Rs[3 * j + 2] + ts[j];
        const double d0 = est[i][0] - q[0], d1 = est[i][1] - q[1], d2 = est[i][2] - q[2];
        es = fmax(es, sqrt(d0 * d0 + d1 * d1 + d2 * d2));
        for (int j = 0; j < 3; ++j) pr[j] = q[0] * K[3 * j] + q[1] * K[3 * j + 1] + q[2] * K[3 * j + 2];
        const double u = pe[i][0] - pr[0] / pr[2], v = pe[i][1] - pr[1] / pr[2];
        ep = fmax(ep, sqrt(u * u + v * v));
      }
      best_s = fmin(best_s, es), best_p = fmin(best_p, ep);
    }
    best_s = block_min(best_s, red);
    best_p = block_min(best_p, red);
    if (threadIdx.x == 0) out[4] = best_s, out[5] = best_p;
  }

  // ---- the last CTA of this pose combines the partial sums in chunk order ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&a.counters[p], 1u) == kChunks - 1;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0.0;
    for (int c = 0; c < kChunks; ++c) tot += reinterpret_cast<volatile double*>(a.partial)[(size_t)p * kChunks + c];
    if (symmetric) out[2] = tot / (double)o.n;
    else out[2] = (double)rh(__fdiv_rn((float)tot, (float)o.n));   // np.mean of halves: float32 sum / n, rounded to half
  }
}

static State* state(oryon_handle* h) {
  if (!h->eval_state) h->eval_state = new State();
  return static_cast<State*>(h->eval_state);
}

void destroy_state(oryon_handle* h) {
  if (!h->eval_state) return;
  State* s = static_cast<State*>(h->eval_state);
  for (auto& kv : s->objects) cudaFree(kv.second.pts), cudaFree(kv.second.syms), cudaFree(kv.second.faces);
  s->ws.release();
  s->vsd_ws.release();
  delete s;
  h->eval_state = nullptr;
}

int set_object(oryon_handle* h, int obj_id, const double* pts, int n, const double* syms, int n_sym) {
  ORYON_REQUIRE(h && pts && syms && n > 0 && n_sym > 0, "oryon_eval_set_object: bad argument");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  State* s = state(h);
  Object& o = s->objects[obj_id];
  if (o.pts) cudaFree(o.pts), cudaFree(o.syms), cudaFree(o.faces);
  o = Object();
  ORYON_CUDA_CHECK(cudaMalloc(&o.pts, sizeof(double) * 3 * (size_t)n));
  ORYON_CUDA_CHECK(cudaMalloc(&o.syms, sizeof(double) * 12 * (size_t)n_sym));
  ORYON_CUDA_CHECK(cudaMemcpy(o.pts, pts, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice));
  ORYON_CUDA_CHECK(cudaMemcpy(o.syms, syms, sizeof(double) * 12 * (size_t)n_sym, cudaMemcpyHostToDevice));
  o.n = n, o.s = n_sym;
  return ORYON_OK;
}

int pose_errors(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams, double* out,
                cudaStream_t st) {
  ORYON_REQUIRE(h && obj_ids && pred && gt && cams && out && P > 0, "oryon_eval_pose_errors: bad argument");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  State* s = state(h);
  const size_t part_bytes = sizeof(double) * kChunks * (size_t)P;
  if (int rc = s->ws.reserve(part_bytes + sizeof(unsigned) * (size_t)P, st)) return rc;
  Args a;
  a.pred = pred, a.gt = gt, a.cams = cams, a.out = out;
  a.partial = s->ws.as<double>();
  a.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->ws.ptr) + part_bytes);
  ORYON_CUDA_CHECK(cudaMemsetAsync(a.counters, 0, sizeof(unsigned) * (size_t)P, st));
  for (int p0 = 0; p0 < P; p0 += kPosesPerLaunch) {
    const int np = std::min(kPosesPerLaunch, P - p0);
    a.p0 = p0;
    for (int i = 0; i < np; ++i) {
      auto it = s->objects.find(obj_ids[p0 + i]);
      ORYON_REQUIRE(it != s->objects.end(), "oryon_eval_pose_errors: object %d was not registered (oryon_eval_set_object)", obj_ids[p0 + i]);
      a.obj[i] = PoseObj{it->second.pts, it->second.syms, it->second.n, it->second.s};
    }
    pose_errors_kernel<<<dim3(kChunks, np), kThreads, 0, st>>>(a);
  }
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

// ------------------------------------------------------------------------------------------------
// VSD (bop_toolkit_lib/pose_error.py:17-96 as called from utils/evaluator.py:279-286)
//   1. depth images of the model in the estimated and the ground-truth pose.  The reference renders them with OpenGL
//      (renderer_vispy.py); here a z-buffer rasteriser with the reference's conventions: sample of pixel (r, c) at image
//      coordinates (c + 0.5, r + 0.5) (projection matrix of renderer_vispy.py:186-231), no culling, nearest surface,
//      eye-space depth at the sample, float32, 0 = background.  float64 edge functions in a fixed operation order
//      (explicit _rn intrinsics, no FMA) so coverage is bit-identical with oracle/vsd_oracle.py.  The rendering step has no
//      reference output to compare with (OpenGL is not available offline): parity of this step is unpinned.
//   2. distance images (misc.py:137-163), visibility masks 'bop19' (visibility.py:9-75, float32 difference), step cost per
//      tau, (#cost + #complement) / #union.  Integer counts: identical to the reference arithmetic given the depth images.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxTaus = 16;

struct VsdObj {
  const double* pts;
  const int32_t* faces;
  int n, f;
};

struct VsdArgs {
  const double* pred;     // [P][16] metres
  const double* gt;
  const double* cams;     // [P][9]
  double* proj;           // [P][2][maxv][3]  (u, v, z)
  unsigned* zbuf;         // [P][2][H][W] float bits, 0xffffffff = empty
  unsigned long long* counts;  // [P][kMaxTaus + 2]: cost counts per tau, complement, union
  const void* depth_test; // [P][H][W] int32 or float32 (mm)
  int depth_is_f32;
  int H, W, maxv, n_tau, p0;
  float delta;
  double taus[kMaxTaus];
  double* out;            // [P][n_tau]
  VsdObj obj[kPosesPerLaunch];
  double diam[kPosesPerLaunch];
};

__global__ void __launch_bounds__(256) vsd_project_kernel(const __grid_constant__ VsdArgs a) {
  const int lp = blockIdx.z, p = a.p0 + lp, which = blockIdx.y;
  const VsdObj o = a.obj[lp];
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= o.n) return;
  const double* T = (which == 0 ? a.pred : a.gt) + 16 * (size_t)p;
  const double* K = a.cams + 9 * (size_t)p;
  // evaluator.py:258-261 + renderer_vispy.py:520-521: pose rounded to half, translation to half millimetres, held in float32
  double R[9], t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) R[3 * r + c] = (double)h2f(T[4 * r + c]);
    t[r] = (double)rh(__fmul_rn(h2f(T[4 * r + 3]), 1000.f));
  }
  const double* x = o.pts + 3 * (size_t)i;
  double cam[3];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    cam[j] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(R[3 * j], x[0]), __dmul_rn(R[3 * j + 1], x[1])), __dmul_rn(R[3 * j + 2], x[2])), t[j]);
  double* q = a.proj + (((size_t)p * 2 + which) * a.maxv + i) * 3;
  q[0] = __dadd_rn(__ddiv_rn(__dmul_rn(K[0], cam[0]), cam[2]), K[2]);
  q[1] = __dadd_rn(__ddiv_rn(__dmul_rn(K[4], cam[1]), cam[2]), K[5]);
  q[2] = cam[2];
}

__global__ void __launch_bounds__(128) vsd_raster_kernel(const __grid_constant__ VsdArgs a) {
  const int lp = blockIdx.z, p = a.p0 + lp, which = blockIdx.y;
  const VsdObj o = a.obj[lp];
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= o.f) return;
  const double* pr = a.proj + ((size_t)p * 2 + which) * a.maxv * 3;
  const int ia = o.faces[3 * f], ib = o.faces[3 * f + 1], ic = o.faces[3 * f + 2];
  const double ua = pr[3 * ia], va = pr[3 * ia + 1], za = pr[3 * ia + 2];
  const double ub = pr[3 * ib], vb = pr[3 * ib + 1], zb = pr[3 * ib + 2];
  const double uc = pr[3 * ic], vc = pr[3 * ic + 1], zc = pr[3 * ic + 2];
  if (!(za > 0 && zb > 0 && zc > 0)) return;
  const int c0 = max(0, (int)ceil(fmin(fmin(ua, ub), uc) - 0.5)), c1 = min(a.W - 1, (int)floor(fmax(fmax(ua, ub), uc) - 0.5));
  const int r0 = max(0, (int)ceil(fmin(fmin(va, vb), vc) - 0.5)), r1 = min(a.H - 1, (int)floor(fmax(fmax(va, vb), vc) - 0.5));
  unsigned* zb_img = a.zbuf + ((size_t)p * 2 + which) * a.H * a.W;
  for (int r = r0; r <= r1; ++r) {
    const double py = (double)r + 0.5;
    for (int c = c0; c <= c1; ++c) {
      const double px = (double)c + 0.5;
      const double w0 = __dsub_rn(__dmul_rn(__dsub_rn(ub, px), __dsub_rn(vc, py)), __dmul_rn(__dsub_rn(uc, px), __dsub_rn(vb, py)));
      const double w1 = __dsub_rn(__dmul_rn(__dsub_rn(uc, px), __dsub_rn(va, py)), __dmul_rn(__dsub_rn(ua, px), __dsub_rn(vc, py)));
      const double w2 = __dsub_rn(__dmul_rn(__dsub_rn(ua, px), __dsub_rn(vb, py)), __dmul_rn(__dsub_rn(ub, px), __dsub_rn(va, py)));
      const bool inside = (w0 >= 0 && w1 >= 0 && w2 >= 0) || (w0 <= 0 && w1 <= 0 && w2 <= 0);
      const double s = __dadd_rn(__dadd_rn(w0, w1), w2);
      if (!inside || s == 0) continue;
      const double invz = __ddiv_rn(__dadd_rn(__dadd_rn(__ddiv_rn(w0, za), __ddiv_rn(w1, zb)), __ddiv_rn(w2, zc)), s);
      const float d = (float)__ddiv_rn(1.0, invz);
      if (d == d) atomicMin(zb_img + (size_t)r * a.W + c, __float_as_uint(d));   // positive floats order like their bit patterns
    }
  }
}

__global__ void __launch_bounds__(256) vsd_reduce_kernel(const __grid_constant__ VsdArgs a) {
  __shared__ unsigned long long sh[kMaxTaus + 2];
  const int lp = blockIdx.y, p = a.p0 + lp;
  if (threadIdx.x < kMaxTaus + 2) sh[threadIdx.x] = 0;
  __syncthreads();
  const double* K = a.cams + 9 * (size_t)p;
  const unsigned* z_est = a.zbuf + ((size_t)p * 2) * a.H * a.W;
  const unsigned* z_gt = z_est + (size_t)a.H * a.W;
  const double diam = a.diam[lp];
  unsigned cost[kMaxTaus];
#pragma unroll
  for (int k = 0; k < kMaxTaus; ++k) cost[k] = 0;
  unsigned n_union = 0, n_comp = 0;
  for (int e = blockIdx.x * 256 + threadIdx.x; e < a.H * a.W; e += gridDim.x * 256) {
    const int r = e / a.W, c = e - r * a.W;
    const double pre_x = __ddiv_rn(__dsub_rn((double)c, K[2]), K[0]), pre_y = __ddiv_rn(__dsub_rn((double)r, K[5]), K[4]);
    auto dist = [&](double d) {
      const double x = __dmul_rn(pre_x, d), y = __dmul_rn(pre_y, d);
      return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(d, d)));
    };
    const size_t o = (size_t)p * a.H * a.W + e;
    const double dt = a.depth_is_f32 ? (double)reinterpret_cast<const float*>(a.depth_test)[o] : (double)reinterpret_cast<const int32_t*>(a.depth_test)[o];
    const unsigned be = z_est[e], bg = z_gt[e];
    const double d_test = dist(dt), d_est = dist(be == 0xffffffffu ? 0.0 : (double)__uint_as_float(be)),
                 d_gt = dist(bg == 0xffffffffu ? 0.0 : (double)__uint_as_float(bg));
    const bool v_gt = (__fsub_rn((float)d_gt, (float)d_test) <= a.delta || d_test == 0) && d_gt > 0;
    const bool v_est = ((__fsub_rn((float)d_est, (float)d_test) <= a.delta || d_test == 0) && d_est > 0) || (v_gt && d_est > 0);
    if (v_gt || v_est) ++n_union;
    if (v_gt && v_est) {
      const double dd = __ddiv_rn(fabs(__dsub_rn(d_gt, d_est)), diam);
#pragma unroll
      for (int k = 0; k < kMaxTaus; ++k)
        if (k < a.n_tau && dd >= a.taus[k]) ++cost[k];
    } else if (v_gt || v_est) {
      ++n_comp;
    }
  }
#pragma unroll
  for (int k = 0; k < kMaxTaus; ++k)
    if (k < a.n_tau && cost[k]) atomicAdd(&sh[k], (unsigned long long)cost[k]);
  if (n_comp) atomicAdd(&sh[kMaxTaus], (unsigned long long)n_comp);
  if (n_union) atomicAdd(&sh[kMaxTaus + 1], (unsigned long long)n_union);
  __syncthreads();
  if (threadIdx.x < kMaxTaus + 2 && sh[threadIdx.x]) atomicAdd(a.counts + (size_t)p * (kMaxTaus + 2) + threadIdx.x, sh[threadIdx.x]);
}

__global__ void vsd_finish_kernel(const unsigned long long* counts, int P, int n_tau, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * n_tau) return;
  const int p = i / n_tau, k = i - p * n_tau;
  const unsigned long long* c = counts + (size_t)p * (kMaxTaus + 2);
  out[i] = c[kMaxTaus + 1] == 0 ? 1.0 : (double)(c[k] + c[kMaxTaus]) / (double)c[kMaxTaus + 1];
}

int set_object_mesh(oryon_handle* h, int obj_id, const int32_t* faces, int n_faces) {
  ORYON_REQUIRE(h && faces && n_faces > 0, "oryon_eval_set_object_mesh: bad argument");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  State* s = state(h);
  auto it = s->objects.find(obj_id);
  ORYON_REQUIRE(it != s->objects.end(), "oryon_eval_set_object_mesh: object %d has no points yet (oryon_eval_set_object)", obj_id);
  for (int i = 0; i < 3 * n_faces; ++i)
    ORYON_REQUIRE(faces[i] >= 0 && faces[i] < it->second.n, "oryon_eval_set_object_mesh: face index %d out of range", faces[i]);
  if (it->second.faces) cudaFree(it->second.faces);
  it->second.faces = nullptr;
  ORYON_CUDA_CHECK(cudaMalloc(&it->second.faces, sizeof(int32_t) * 3 * (size_t)n_faces));
  ORYON_CUDA_CHECK(cudaMemcpy(it->second.faces, faces, sizeof(int32_t) * 3 * (size_t)n_faces, cudaMemcpyHostToDevice));
  it->second.f = n_faces;
  return ORYON_OK;
}

int vsd(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams, const void* depth_test,
        int depth_is_f32, int H, int W, double delta, const double* taus, int n_tau, const double* diameters, double* out,
        cudaStream_t st) {
  ORYON_REQUIRE(h && obj_ids && pred && gt && cams && depth_test && taus && diameters && out, "oryon_eval_vsd: null argument");
  ORYON_REQUIRE(P > 0 && H > 0 && W > 0 && n_tau > 0 && n_tau <= kMaxTaus, "oryon_eval_vsd: bad sizes (n_tau <= %d)", kMaxTaus);
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  State* s = state(h);
  int maxv = 0, maxf = 0;
  for (int p = 0; p < P; ++p) {
    auto it = s->objects.find(obj_ids[p]);
    ORYON_REQUIRE(it != s->objects.end() && it->second.faces, "oryon_eval_vsd: object %d has no mesh (oryon_eval_set_object_mesh)", obj_ids[p]);
    maxv = std::max(maxv, it->second.n), maxf = std::max(maxf, it->second.f);
  }
  const size_t proj_bytes = sizeof(double) * 3 * (size_t)maxv * 2 * P;
  const size_t z_bytes = sizeof(unsigned) * (size_t)H * W * 2 * P;
  const size_t cnt_bytes = sizeof(unsigned long long) * (kMaxTaus + 2) * (size_t)P;
  if (int rc = s->vsd_ws.reserve(proj_bytes + z_bytes + cnt_bytes, st)) return rc;
  VsdArgs a;
  a.pred = pred, a.gt = gt, a.cams = cams;
  a.proj = s->vsd_ws.as<double>();
  a.zbuf = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(s->vsd_ws.ptr) + proj_bytes);
  a.counts = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(s->vsd_ws.ptr) + proj_bytes + z_bytes);
  a.depth_test = depth_test, a.depth_is_f32 = depth_is_f32;
  a.H = H, a.W = W, a.maxv = maxv, a.n_tau = n_tau, a.delta = (float)delta, a.out = out;
  for (int k = 0; k < kMaxTaus; ++k) a.taus[k] = k < n_tau ? taus[k] : 0.0;
  ORYON_CUDA_CHECK(cudaMemsetAsync(a.zbuf, 0xff, z_bytes, st));
  ORYON_CUDA_CHECK(cudaMemsetAsync(a.counts, 0, cnt_bytes, st));
  for (int p0 = 0; p0 < P; p0 += kPosesPerLaunch) {
    const int np = std::min(kPosesPerLaunch, P - p0);
    a.p0 = p0;
    for (int i = 0; i < np; ++i) {
      const Object& o = s->objects[obj_ids[p0 + i]];
      a.obj[i] = VsdObj{o.pts, o.faces, o.n, o.f};
      a.diam[i] = diameters[p0 + i];
    }
    vsd_project_kernel<<<dim3((maxv + 255) / 256, 2, np), 256, 0, st>>>(a);
    vsd_raster_kernel<<<dim3((maxf + 127) / 128, 2, np), 128, 0, st>>>(a);
    vsd_reduce_kernel<<<dim3(64, np), 256, 0, st>>>(a);
  }
  vsd_finish_kernel<<<(P * n_tau + 127) / 128, 128, 0, st>>>(a.counts, P, n_tau, out);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

}  // namespace eval
}  // namespace oryon
