// a7 -- mask post-processing (reference losses.py:40-62 no-grad part, utils/metrics.py:18-40, pipeline.py:372-395):
// predicted mask = sigmoid(logit) > th, nearest resize of the ground-truth / external mask to the feature-map size,
// per-image IoU and pixel counts (the validity test of is_detection_valid).  One CTA per image, HBM-bound.
#include "common.cuh"

namespace oryon {
namespace maskpost {

__global__ void __launch_bounds__(1024) mask_postproc_kernel(const float* logits, int H, int W, float th, const uint8_t* gt, int Hg, int Wg,
                                                            int32_t* pred, int32_t* gt_resized, int32_t* n_pred, int32_t* n_gt, float* iou) {
  __shared__ int red[32][4];
  const int b = blockIdx.x, HW = H * W;
  // F.interpolate(mode='nearest'): src = min(floor(dst * (float)in / out), in - 1)
  const float sy = (float)Hg / (float)H, sx = (float)Wg / (float)W;
  int c_pred = 0, c_gt = 0, c_inter = 0, c_union = 0;
  for (int i = threadIdx.x; i < HW; i += 1024) {
    int p = 0;
    if (logits) {
      const float x = logits[(size_t)b * HW + i];
      p = (1.f / (1.f + expf(-x))) > th ? 1 : 0;
      pred[(size_t)b * HW + i] = p;
    }
    int g = 0;
    if (gt) {
      const int y = i / W, xw = i % W;
      const int ys = min((int)floorf((float)y * sy), Hg - 1), xs = min((int)floorf((float)xw * sx), Wg - 1);
      g = gt[((size_t)b * Hg + ys) * Wg + xs];
      if (gt_resized) gt_resized[(size_t)b * HW + i] = g;
    }
    c_pred += p, c_gt += (g == 1), c_inter += (p != 0 && g != 0), c_union += (p != 0 || g != 0);
  }
  int v[4] = {c_pred, c_gt, c_inter, c_union};
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
  if ((threadIdx.x & 31) == 0)
    for (int q = 0; q < 4; ++q) red[threadIdx.x >> 5][q] = v[q];
  __syncthreads();
  if (threadIdx.x == 0) {
    int t[4] = {0, 0, 0, 0};
    for (int w = 0; w < 32; ++w)
      for (int q = 0; q < 4; ++q) t[q] += red[w][q];
    if (n_pred) n_pred[b] = t[0];
    if (n_gt) n_gt[b] = t[1];
    if (iou) iou[b] = (float)t[2] / (float)t[3];  // 0/0 -> NaN as the reference's true division
  }
}

int run(oryon_handle* h, const float* logits, int B, int H, int W, float th, const uint8_t* gt, int Hg, int Wg, int32_t* pred,
        int32_t* gt_resized, int32_t* n_pred, int32_t* n_gt, float* iou, cudaStream_t st) {
  ORYON_REQUIRE(h && B > 0 && H > 0 && W > 0 && (logits || gt), "oryon_mask_postproc: bad argument");
  ORYON_REQUIRE(!logits || pred, "oryon_mask_postproc: pred_mask output required with logits");
  ORYON_REQUIRE(!gt || (Hg > 0 && Wg > 0), "oryon_mask_postproc: gt size");
  ORYON_CUDA_CHECK(cudaSetDevice(h->device));
  h->span_begin(KID_ELTWISE, st);
  mask_postproc_kernel<<<B, 1024, 0, st>>>(logits, H, W, th, gt, Hg, Wg, pred, gt_resized, n_pred, n_gt, iou);
  h->span_end(st);
  ORYON_CUDA_CHECK(cudaGetLastError());
  return ORYON_OK;
}

}  // namespace maskpost
}  // namespace oryon
