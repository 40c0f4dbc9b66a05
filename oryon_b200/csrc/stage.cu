// N1 -- batch staging on the GPU: the per-sample CPU work the reference's loader does between the decoded frame and the
// network input, for a whole batch in one launch.
//   rgb    preprocess_item (utils/data/common.py:48-49): uint8 HWC -> CHW / 255. in FLOAT64, then resize to img_size
//          (utils/augmentations.py:137: torchvision F.resize BILINEAR on a tensor; the reference pins torchvision 0.13,
//          where that is F.interpolate(mode='bilinear', align_corners=False) without antialiasing), evaluated in float64
//          as ATen's CPU kernel does for a float64 tensor, then CollateWrapper's .to(torch.float32) (datasets.py:205)
//   mask   (mask == mask_id) -> 0/1 (common.py:62-64), NEAREST resize (augmentations.py:138: floor(dst * float(in/out))),
//          .to(torch.uint8) (datasets.py:207)
// HBM-bound, read once / written once: 0.9 MB uint8 in and 0.6 MB float32 out per 480x640 frame.
#include "common.cuh"

namespace oryon {
namespace stage {

struct Args {
  const uint8_t* rgb;    // [B][H][W][3]
  const void* mask;      // [B][H][W] uint8 or int32, or nullptr
  int mask_is_i32;
  const int32_t* mask_ids;   // [B] device, or nullptr (then any non-zero value counts, id == 1 semantics on a 0/1 mask)
  int B, H, W, S_h, S_w;
  float* out_rgb;        // [B][3][S_h][S_w]
  uint8_t* out_mask;     // [B][S_h][S_w]
};

// area_pixel_compute_source_index (ATen UpSample.h), align_corners = False, in double
__device__ __forceinline__ void src_index(int dst, double scale, int in_size, int& i0, int& i1, double& w0, double& w1) {
  double s = scale * ((double)dst + 0.5) - 0.5;
  if (s < 0.0) s = 0.0;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  w1 = s - (double)i0;
  w0 = 1.0 - w1;
}

__global__ void __launch_bounds__(256) stage_kernel(Args a) {
  const int64_t per_img = (int64_t)a.S_h * a.S_w;
  const int64_t total = (int64_t)a.B * per_img;
  const double sy = (double)a.H / (double)a.S_h, sx = (double)a.W / (double)a.S_w;
  const float ny = (float)a.H / (float)a.S_h, nx = (float)a.W / (float)a.S_w;   // nearest: compute_scales_value<float>
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int b = (int)(e / per_img);
    const int r = (int)(e - (int64_t)b * per_img);
    const int oy = r / a.S_w, ox = r - oy * a.S_w;
    int y0, y1, x0, x1;
    double wy0, wy1, wx0, wx1;
    src_index(oy, sy, a.H, y0, y1, wy0, wy1);
    src_index(ox, sx, a.W, x0, x1, wx0, wx1);
    const uint8_t* img = a.rgb + (int64_t)b * a.H * a.W * 3;
    const uint8_t* p00 = img + ((int64_t)y0 * a.W + x0) * 3;
    const uint8_t* p01 = img + ((int64_t)y0 * a.W + x1) * 3;
    const uint8_t* p10 = img + ((int64_t)y1 * a.W + x0) * 3;
    const uint8_t* p11 = img + ((int64_t)y1 * a.W + x1) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double v00 = (double)p00[c] / 255.0, v01 = (double)p01[c] / 255.0, v10 = (double)p10[c] / 255.0, v11 = (double)p11[c] / 255.0;
      // ATen: wy0 * (wx0 * v00 + wx1 * v01) + wy1 * (wx0 * v10 + wx1 * v11); no contraction
      const double top = __dadd_rn(__dmul_rn(wx0, v00), __dmul_rn(wx1, v01));
      const double bot = __dadd_rn(__dmul_rn(wx0, v10), __dmul_rn(wx1, v11));
      const double v = __dadd_rn(__dmul_rn(wy0, top), __dmul_rn(wy1, bot));
      a.out_rgb[((int64_t)b * 3 + c) * per_img + r] = (float)v;
    }
    if (a.mask) {
      const int my = min((int)floorf((float)oy * ny), a.H - 1), mx = min((int)floorf((float)ox * nx), a.W - 1);
      const int64_t o = ((int64_t)b * a.H + my) * a.W + mx;
      const int v = a.mask_is_i32 ? reinterpret_cast<const int32_t*>(a.mask)[o] : (int)reinterpret_cast<const uint8_t*>(a.mask)[o];
      const int id = a.mask_ids ? a.mask_ids[b] : 1;
      a.out_mask[e] = v == id ? 1 : 0;
    }
  }
}

int run(oryon_handle* h, const uint8_t* rgb, const void* mask, int mask_is_i32, const int32_t* mask_ids, int B, int H, int W, int S_h, int S_w,
        float* out_rgb, uint8_t* out_mask, cudaStream_t st) {
  ORYON_REQUIRE(h && rgb && out_rgb && B > 0 && H > 0 && W > 0 && S_h > 0 && S_w > 0, "oryon_stage_inputs: bad argument");
  ORYON_REQUIRE(!mask || out_mask, "oryon_stage_inputs: mask given without an output buffer");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  Args a;
  a.rgb = rgb, a.mask = mask, a.mask_is_i32 = mask_is_i32, a.mask_ids = mask_ids;
  a.B = B, a.H = H, a.W = W, a.S_h = S_h, a.S_w = S_w, a.out_rgb = out_rgb, a.out_mask = out_mask;
  const int64_t total = (int64_t)B * S_h * S_w;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sm_count * 16);
  h->span_begin(KID_ELTWISE, st);
  stage_kernel<<<grid, 256, 0, st>>>(a);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

}  // namespace stage
}  // namespace oryon
