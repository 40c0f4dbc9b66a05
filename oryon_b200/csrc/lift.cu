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

static Cam make_cam(const double* k) {
  Cam c;
  c.fx = (float)k[0], c.cx = (float)k[2], c.fy = (float)k[4], c.cy = (float)k[5];
  return c;
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

}  // namespace lift
}  // namespace oryon
