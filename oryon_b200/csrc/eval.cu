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
  int n = 0, s = 0;
};

struct State {
  std::map<int, Object> objects;
  DeviceBuffer ws;          // partial sums [P][kChunks] doubles + counters [P] ints
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
  for (auto& kv : s->objects) cudaFree(kv.second.pts), cudaFree(kv.second.syms);
  s->ws.release();
  delete s;
  h->eval_state = nullptr;
}

int set_object(oryon_handle* h, int obj_id, const double* pts, int n, const double* syms, int n_sym) {
  ORYON_REQUIRE(h && pts && syms && n > 0 && n_sym > 0, "oryon_eval_set_object: bad argument");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  State* s = state(h);
  Object& o = s->objects[obj_id];
  if (o.pts) cudaFree(o.pts), cudaFree(o.syms);
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

}  // namespace eval
}  // namespace oryon
