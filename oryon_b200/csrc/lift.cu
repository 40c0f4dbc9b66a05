// a9 + a10 -- coordinate scaling, bounds test, truncation and pin-hole lifting
// (reference pipeline.py:447-460, utils/coordinates.py:5-13,36-47, utils/pcd.py:35-81).
//
// Every float operation is written with explicit round-to-nearest intrinsics in the reference's
// order (no FMA contraction), so the output is bit-identical to the PyTorch CPU evaluation:
//   y' = f32(y) * f32(ratio)      valid = 0 <= y' < H      yi = trunc(y')      z = f32(depth[yi][xi])
//   X = ((f32(xi) - f32(cx)) * z) / f32(fx)    ...    out = X / 1000f
#include "common.cuh"

namespace oryon {
namespace lift {

struct Cam {
  float fx, fy, cx, cy;
};

__device__ __forceinline__ float load_depth(const void* depth, int dtype, size_t i) {
  switch (dtype) {
    case ORYON_DEPTH_I32: return static_cast<float>(reinterpret_cast<const int32_t*>(depth)[i]);
    case ORYON_DEPTH_F32: return reinterpret_cast<const float*>(depth)[i];
    case ORYON_DEPTH_I16: return static_cast<float>(reinterpret_cast<const int16_t*>(depth)[i]);
    default: return static_cast<float>(reinterpret_cast<const uint16_t*>(depth)[i]);
  }
}

__device__ __forceinline__ void lift_point(const void* depth, int dtype, int W, Cam c, long long x, long long y, float scale_div,
                                           float* out) {
  const float z = load_depth(depth, dtype, (size_t)y * W + x);
  const float px = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, c.cx), z), c.fx);
  const float py = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, c.cy), z), c.fy);
  if (scale_div != 0.f) {
    out[0] = __fdiv_rn(px, scale_div), out[1] = __fdiv_rn(py, scale_div), out[2] = __fdiv_rn(z, scale_div);
  } else {
    out[0] = px, out[1] = py, out[2] = z;
  }
}

__global__ void lift_pcd_kernel(const void* depth, int dtype, int H, int W, Cam c, const long long* xs, const long long* ys, int n,
                                float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long x = xs[i], y = ys[i];
  // torch advanced indexing wraps negative indices; anything else out of range is a caller error.
  if (x < 0) x += W;
  if (y < 0) y += H;
  lift_point(depth, dtype, W, c, x, y, 0.f, out + 3 * (size_t)i);
}

struct CorrArgs {
  const long long* corrs;
  int n;
  float ry_a, rx_a, ry_q, rx_q;  // float32(target/source) ratios
  const void* depth_a;
  const void* depth_q;
  int dtype, Ha, Wa, Hq, Wq;
  Cam cam_a, cam_q;
  float* pcd_a;
  float* pcd_q;
  int32_t* n_valid;
};

// One CTA; stable compaction of the bounds mask with a block-wide scan, 1024 rows per pass.
__global__ void __launch_bounds__(1024) corrs_to_pcd_kernel(CorrArgs a) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < a.n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    bool ok = false;
    float ya = 0, xa = 0, yq = 0, xq = 0;
    if (i < a.n) {
      ya = __fmul_rn((float)a.corrs[4 * (size_t)i + 0], a.ry_a);
      xa = __fmul_rn((float)a.corrs[4 * (size_t)i + 1], a.rx_a);
      yq = __fmul_rn((float)a.corrs[4 * (size_t)i + 2], a.ry_q);
      xq = __fmul_rn((float)a.corrs[4 * (size_t)i + 3], a.rx_q);
      ok = ya >= 0.f && ya < (float)a.Ha && xa >= 0.f && xa < (float)a.Wa && yq >= 0.f && yq < (float)a.Hq && xq >= 0.f &&
           xq < (float)a.Wq;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (ok) {
      const int o = off + __popc(bal & ((1u << lane) - 1u));
      lift_point(a.depth_a, a.dtype, a.Wa, a.cam_a, (long long)xa, (long long)ya, 1000.f, a.pcd_a + 3 * (size_t)o);
      lift_point(a.depth_q, a.dtype, a.Wq, a.cam_q, (long long)xq, (long long)yq, 1000.f, a.pcd_q + 3 * (size_t)o);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_tot[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.n_valid = base;
}

// Batched tail of nn_correspondences + pipeline.py:447-460 (one CTA per pair): gathers the caller-selected rows
// into (y1,x1,y2,x2) correspondences, then scales / bounds-tests / compacts / lifts exactly like corrs_to_pcd_kernel.
static Cam make_cam(const double* k) {
  Cam c;
  c.fx = (float)k[0], c.cx = (float)k[2], c.fy = (float)k[4], c.cy = (float)k[5];
  return c;
}

constexpr int kSelectPairsPerLaunch = 64;   // 2 KB of intrinsics ride in the kernel parameters (no staging copy, no sync)
struct SelectArgs {
  const int32_t* rows;      // [B][n]
  const int32_t* roi_a;     // [B][cap_a]
  const int32_t* roi_q;     // [B][cap_q]
  const int32_t* nn_idx;    // [B][cap_a]
  int n, cap_a, cap_q, feat_w;
  float ry_a, rx_a, ry_q, rx_q;
  const void* depth_a;      // [B][Ha][Wa]
  const void* depth_q;
  int dtype, Ha, Wa, Hq, Wq;
  long long* corrs;         // [B][n][4]
  float* pcd_a;             // [B][n][3]
  float* pcd_q;
  int32_t* n_valid;         // [B]
  int b0;                   // first pair of this launch
  Cam cams[2 * kSelectPairsPerLaunch];   // (anchor, query) intrinsics of pairs b0 .. b0 + gridDim.x
};

__global__ void __launch_bounds__(1024) select_lift_kernel(const __grid_constant__ SelectArgs a) {
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int b = a.b0 + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t* rows = a.rows + (size_t)b * a.n;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  if (rows[0] < 0) {  // pair without correspondences
    if (threadIdx.x == 0) a.n_valid[b] = -1;
    return;
  }
  const size_t esz = (a.dtype == ORYON_DEPTH_I16 || a.dtype == ORYON_DEPTH_U16) ? 2 : 4;
  const char* da = reinterpret_cast<const char*>(a.depth_a) + (size_t)b * a.Ha * a.Wa * esz;
  const char* dq = reinterpret_cast<const char*>(a.depth_q) + (size_t)b * a.Hq * a.Wq * esz;
  const Cam ca = a.cams[2 * blockIdx.x], cq = a.cams[2 * blockIdx.x + 1];
  for (int i0 = 0; i0 < a.n; i0 += 1024) {
    const int i = i0 + threadIdx.x;
    bool ok = false;
    float ya = 0, xa = 0, yq = 0, xq = 0;
    if (i < a.n) {
      const int r = rows[i];
      const int p1 = a.roi_a[(size_t)b * a.cap_a + r];
      const int p2 = a.roi_q[(size_t)b * a.cap_q + a.nn_idx[(size_t)b * a.cap_a + r]];
      const long long y1 = p1 / a.feat_w, x1 = p1 % a.feat_w, y2 = p2 / a.feat_w, x2 = p2 % a.feat_w;
      long long* c = a.corrs + ((size_t)b * a.n + i) * 4;
      c[0] = y1, c[1] = x1, c[2] = y2, c[3] = x2;
      ya = __fmul_rn((float)y1, a.ry_a), xa = __fmul_rn((float)x1, a.rx_a);
      yq = __fmul_rn((float)y2, a.ry_q), xq = __fmul_rn((float)x2, a.rx_q);
      ok = ya >= 0.f && ya < (float)a.Ha && xa >= 0.f && xa < (float)a.Wa && yq >= 0.f && yq < (float)a.Hq && xq >= 0.f &&
           xq < (float)a.Wq;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (ok) {
      const size_t o = (size_t)b * a.n + off + __popc(bal & ((1u << lane) - 1u));
      lift_point(da, a.dtype, a.Wa, ca, (long long)xa, (long long)ya, 1000.f, a.pcd_a + 3 * o);
      lift_point(dq, a.dtype, a.Wq, cq, (long long)xq, (long long)yq, 1000.f, a.pcd_q + 3 * o);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_tot[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) a.n_valid[b] = base;
}

int run_lift(oryon_handle* h, const void* depth, int dtype, int H, int W, const double* cam, const int64_t* xs, const int64_t* ys, int n,
             float* out, cudaStream_t st) {
  ORYON_REQUIRE(h && depth && cam && out && H > 0 && W > 0 && n >= 0, "oryon_lift_pcd: bad argument");
  ORYON_REQUIRE(dtype >= ORYON_DEPTH_I32 && dtype <= ORYON_DEPTH_U16, "oryon_lift_pcd: unknown depth dtype %d", dtype);
  if (n == 0) return ORYON_OK;
  ORYON_REQUIRE(xs && ys, "oryon_lift_pcd: null index arrays");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  lift_pcd_kernel<<<(n + 255) / 256, 256, 0, st>>>(depth, dtype, H, W, make_cam(cam), reinterpret_cast<const long long*>(xs),
                                                 reinterpret_cast<const long long*>(ys), n, out);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

int run_corrs_to_pcd(oryon_handle* h, const int64_t* corrs, int n, int feat_h, int feat_w, const void* depth_a, const void* depth_q,
                     int dtype, int Ha, int Wa, int Hq, int Wq, const double* cam_a, const double* cam_q, float* pcd_a, float* pcd_q,
                     int32_t* n_valid, cudaStream_t st) {
  ORYON_REQUIRE(h && depth_a && depth_q && cam_a && cam_q && pcd_a && pcd_q && n_valid, "oryon_corrs_to_pcd: null argument");
  ORYON_REQUIRE(n >= 0 && feat_h > 0 && feat_w > 0 && Ha > 0 && Wa > 0 && Hq > 0 && Wq > 0, "oryon_corrs_to_pcd: bad sizes");
  ORYON_REQUIRE(dtype >= ORYON_DEPTH_I32 && dtype <= ORYON_DEPTH_U16, "oryon_corrs_to_pcd: unknown depth dtype %d", dtype);
  ORYON_REQUIRE(n == 0 || corrs, "oryon_corrs_to_pcd: null corrs");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  CorrArgs a;
  a.corrs = reinterpret_cast<const long long*>(corrs);
  a.n = n;
  // reference: new_coords * (target / source) with a Python-float ratio, evaluated in float32
  a.ry_a = (float)((double)Ha / (double)feat_h), a.rx_a = (float)((double)Wa / (double)feat_w);
  a.ry_q = (float)((double)Hq / (double)feat_h), a.rx_q = (float)((double)Wq / (double)feat_w);
  a.depth_a = depth_a, a.depth_q = depth_q, a.dtype = dtype;
  a.Ha = Ha, a.Wa = Wa, a.Hq = Hq, a.Wq = Wq;
  a.cam_a = make_cam(cam_a), a.cam_q = make_cam(cam_q);
  a.pcd_a = pcd_a, a.pcd_q = pcd_q, a.n_valid = n_valid;
  corrs_to_pcd_kernel<<<1, 1024, 0, st>>>(a);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

int run_select_lift(oryon_handle* h, const int32_t* rows, int B, int n, const int32_t* roi_a, const int32_t* roi_q, const int32_t* nn_idx,
                    int cap_a, int cap_q, int feat_h, int feat_w, const void* depth_a, const void* depth_q, int dtype, int Ha, int Wa,
                    int Hq, int Wq, const double* cams_a, const double* cams_q, int64_t* corrs, float* pcd_a, float* pcd_q,
                    int32_t* n_valid, cudaStream_t st) {
  ORYON_REQUIRE(h && rows && roi_a && roi_q && nn_idx && depth_a && depth_q && cams_a && cams_q && corrs && pcd_a && pcd_q && n_valid,
                "oryon_select_lift: null argument");
  ORYON_REQUIRE(B > 0 && n > 0 && cap_a > 0 && cap_q > 0 && feat_h > 0 && feat_w > 0 && Ha > 0 && Wa > 0 && Hq > 0 && Wq > 0,
                "oryon_select_lift: bad sizes");
  ORYON_REQUIRE(dtype >= ORYON_DEPTH_I32 && dtype <= ORYON_DEPTH_U16, "oryon_select_lift: unknown depth dtype %d", dtype);
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  SelectArgs a;
  a.rows = rows, a.roi_a = roi_a, a.roi_q = roi_q, a.nn_idx = nn_idx;
  a.n = n, a.cap_a = cap_a, a.cap_q = cap_q, a.feat_w = feat_w;
  a.ry_a = (float)((double)Ha / (double)feat_h), a.rx_a = (float)((double)Wa / (double)feat_w);
  a.ry_q = (float)((double)Hq / (double)feat_h), a.rx_q = (float)((double)Wq / (double)feat_w);
  a.depth_a = depth_a, a.depth_q = depth_q, a.dtype = dtype;
  a.Ha = Ha, a.Wa = Wa, a.Hq = Hq, a.Wq = Wq;
  a.corrs = reinterpret_cast<long long*>(corrs);
  a.pcd_a = pcd_a, a.pcd_q = pcd_q, a.n_valid = n_valid;
  h->span_begin(KID_LIFT, st);
  for (int b0 = 0; b0 < B; b0 += kSelectPairsPerLaunch) {
    const int nb = B - b0 < kSelectPairsPerLaunch ? B - b0 : kSelectPairsPerLaunch;
    a.b0 = b0;
    for (int b = 0; b < nb; ++b) a.cams[2 * b] = make_cam(cams_a + 9 * (b0 + b)), a.cams[2 * b + 1] = make_cam(cams_q + 9 * (b0 + b));
    select_lift_kernel<<<nb, 1024, 0, st>>>(a);
  }
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

}  // namespace lift
}  // namespace oryon
