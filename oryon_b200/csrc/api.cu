// C ABI of liboryon_b200.so (include/oryon_b200.h): argument validation, handle lifecycle and
// dispatch into the kernel translation units.
#include "common.cuh"

namespace oryon {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int DeviceBuffer::reserve(size_t want, cudaStream_t stream) {
  if (want <= bytes) return ORYON_OK;
  // grow geometrically; growth is rare (sizes are monotone for a workload) so a stream sync is fine
  size_t cap = bytes ? bytes : 256;
  while (cap < want) cap += cap / 2 + 256;
  cap = (cap + 255) & ~size_t(255);
  if (ptr) {
    cudaStreamSynchronize(stream);
    cudaFree(ptr);
    ptr = nullptr, bytes = 0;
  }
  cudaError_t e = cudaMalloc(&ptr, cap);
  if (e != cudaSuccess) {
    ptr = nullptr;
    set_error("cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? ORYON_ERR_OUT_OF_MEMORY : ORYON_ERR_CUDA;
  }
  bytes = cap;
  // never expose uninitialised bits to the tensor cores (NaN * 0 in K padding would poison rows)
  e = cudaMemsetAsync(ptr, 0, cap, stream);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return ORYON_ERR_CUDA;
  }
  return ORYON_OK;
}

void DeviceBuffer::release() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr, bytes = 0;
}

namespace match {
int run_match(oryon_handle*, const float*, const float*, int, int, int, int, const int32_t*, const int32_t*, const int32_t*,
              const int32_t*, int, int, int, int32_t*, float*, cudaStream_t);
int run_mask_to_roi(oryon_handle*, const int32_t*, int, int, int, int32_t*, int32_t*, cudaStream_t);
int read_stats(oryon_handle*, int64_t*, cudaStream_t);
int read_hist(oryon_handle*, int64_t*, cudaStream_t);
int plan_debug(const int32_t*, const int32_t*, int, int, int, int32_t*, int, int32_t*, int32_t*);
}  // namespace match
namespace stage {
int run(oryon_handle*, const uint8_t*, const void*, int, const int32_t*, int, int, int, int, int, float*, uint8_t*, cudaStream_t);
}  // namespace stage
namespace eval {
void destroy_state(oryon_handle*);
int set_object(oryon_handle*, int, const double*, int, const double*, int);
int pose_errors(oryon_handle*, int, const int32_t*, const double*, const double*, const double*, double*, cudaStream_t);
int set_object_mesh(oryon_handle*, int, const int32_t*, int);
int vsd(oryon_handle*, int, const int32_t*, const double*, const double*, const double*, const void*, int, int, int, double, const double*, int,
        const double*, double*, cudaStream_t);
}  // namespace eval
namespace lift {
int run_lift(oryon_handle*, const void*, int, int, int, const double*, const int64_t*, const int64_t*, int, float*, cudaStream_t);
int run_corrs_to_pcd(oryon_handle*, const int64_t*, int, int, int, const void*, const void*, int, int, int, int, int, const double*,
                     const double*, float*, float*, int32_t*, cudaStream_t);
int run_select_lift(oryon_handle*, const int32_t*, int, int, const int32_t*, const int32_t*, const int32_t*, int, int, int, int, const void*,
                    const void*, int, int, int, int, int, const double*, const double*, int64_t*, float*, float*, int32_t*, cudaStream_t);
}  // namespace lift
namespace gemm {
int run_gemm_f32(oryon_handle*, const float*, const float*, const float*, const float*, float*, int, int, int, int, int, float, int,
                 cudaStream_t);
}  // namespace gemm
namespace net {
int set_weight(oryon_handle*, const char*, const float*, int64_t);
int finalize(oryon_handle*, const oryon_backbone_config*, cudaStream_t);
int text_forward(oryon_handle*, const int32_t*, int, float*, cudaStream_t);
int backbone_forward(oryon_handle*, const float*, const float*, int, const float*, float*, float*, float*, float*,
                     const oryon_backbone_debug*, cudaStream_t);
void destroy_backbone(oryon_handle*);
}  // namespace net
namespace maskpost {
int run(oryon_handle*, const float*, int, int, int, float, const uint8_t*, int, int, int32_t*, int32_t*, int32_t*, int32_t*, float*,
        cudaStream_t);
}  // namespace maskpost
namespace pdsc {
int load_weights(oryon_handle*, const oryon_pointdsc_config*, const float*, int64_t, cudaStream_t);
int run_pose(oryon_handle*, const float*, const float*, const int32_t*, int, int, float*, const oryon_pointdsc_debug*, cudaStream_t);
void destroy_model(oryon_handle*);
}  // namespace pdsc

}  // namespace oryon

cudaEvent_t oryon_handle::take_event() {
  if (!free_events.empty()) {
    cudaEvent_t e = free_events.back();
    free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void oryon_handle::span_begin(int id, cudaStream_t st) {
  if (!profiling) return;
  oryon::ProfSpan s{span_alias >= 0 ? span_alias : id, take_event(), take_event()};
  cudaEventRecord(s.a, st);
  spans.push_back(s);
}
void oryon_handle::span_end(cudaStream_t st) {
  if (!profiling || spans.empty()) return;
  cudaEventRecord(spans.back().b, st);
}

int64_t oryon_handle::workspace_bytes() const {
  return (int64_t)(rows16_a.bytes + rows16_q.bytes + rows32_a.bytes + rows32_q.bytes + cand.bytes + counters.bytes +
                   overflow_rows.bytes + pair_meta.bytes + match_plan.bytes + lift_scratch.bytes + pdsc_ws.bytes + gemm_scratch.bytes);
}

extern "C" {

int oryon_abi_version(void) { return ORYON_ABI_VERSION; }

const char* oryon_last_error(void) { return oryon::g_last_error.c_str(); }

int oryon_create(int device, oryon_handle** out) {
  ORYON_REQUIRE(out != nullptr, "oryon_create: out is NULL");
  *out = nullptr;
  int count = 0;
  ORYON_CUDA_CHECK(cudaGetDeviceCount(&count));
  ORYON_REQUIRE(device >= 0 && device < count, "oryon_create: device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  ORYON_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    oryon::set_error("oryon_create: device %d is sm_%d%d; this library contains sm_100a code only and has no fallback", device,
                     prop.major, prop.minor);
    return ORYON_ERR_UNSUPPORTED_DEVICE;
  }
  ORYON_CUDA_CHECK(cudaSetDevice(device));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  ORYON_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    oryon::set_error("oryon_create: cuTensorMapEncodeTiled not available from the driver");
    return ORYON_ERR_CUDA;
  }
  oryon_handle* h = new oryon_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major, h->cc_minor = prop.minor;
  h->encode_tiled = reinterpret_cast<oryon::PFN_encodeTiled>(fn);
  *out = h;
  return ORYON_OK;
}

int oryon_destroy(oryon_handle* h) {
  if (!h) return ORYON_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  h->rows16_a.release(), h->rows16_q.release(), h->rows32_a.release(), h->rows32_q.release();
  for (auto& s : h->spans) cudaEventDestroy(s.a), cudaEventDestroy(s.b);
  for (auto e : h->free_events) cudaEventDestroy(e);
  h->cand.release(), h->counters.release(), h->overflow_rows.release(), h->pair_meta.release(), h->match_plan.release(), h->lift_scratch.release();
  oryon::pdsc::destroy_model(h);
  oryon::net::destroy_backbone(h);
  oryon::eval::destroy_state(h);
  h->pdsc_ws.release();
  h->gemm_scratch.release();
  delete h;
  return ORYON_OK;
}

int64_t oryon_workspace_bytes(const oryon_handle* h) { return h ? h->workspace_bytes() : 0; }

int oryon_profile_enable(oryon_handle* h, int enable) {
  ORYON_REQUIRE(h != nullptr, "oryon_profile_enable: null handle");
  h->profiling = enable != 0;
  return ORYON_OK;
}

int oryon_profile_read(oryon_handle* h, double* total_ms, int64_t* launches, int n_ids) {
  ORYON_REQUIRE(h && total_ms && launches && n_ids > 0, "oryon_profile_read: bad argument");
  for (int i = 0; i < n_ids; ++i) total_ms[i] = 0.0, launches[i] = 0;
  for (auto& s : h->spans) {
    ORYON_CUDA_CHECK(cudaEventSynchronize(s.b));
    float ms = 0.f;
    ORYON_CUDA_CHECK(cudaEventElapsedTime(&ms, s.a, s.b));
    if (s.id >= 0 && s.id < n_ids) total_ms[s.id] += ms, launches[s.id] += 1;
    h->free_events.push_back(s.a), h->free_events.push_back(s.b);
  }
  h->spans.clear();
  return ORYON_OK;
}

int oryon_match_nn(oryon_handle* h, const float* feat_a, const float* feat_q, int B, int D, int HW_a, int HW_q, const int32_t* roi_a,
                   const int32_t* roi_q, const int32_t* n_a, const int32_t* n_q, int cap_a, int cap_q, int mode, int32_t* out_idx,
                   float* out_dist, void* stream) {
  return oryon::match::run_match(h, feat_a, feat_q, B, D, HW_a, HW_q, roi_a, roi_q, n_a, n_q, cap_a, cap_q, mode, out_idx, out_dist,
                                 static_cast<cudaStream_t>(stream));
}

int oryon_match_last_stats(oryon_handle* h, int64_t stats[4], void* stream) {
  return oryon::match::read_stats(h, stats, static_cast<cudaStream_t>(stream));
}

int oryon_match_set_hist(oryon_handle* h, int enable) {
  ORYON_REQUIRE(h, "oryon_match_set_hist: null handle");
  h->match_hist = enable != 0;
  return ORYON_OK;
}

int oryon_match_list_hist(oryon_handle* h, int64_t hist[26], void* stream) {
  return oryon::match::read_hist(h, hist, static_cast<cudaStream_t>(stream));
}

int oryon_match_plan(const int32_t* n_a, const int32_t* n_q, int B, int sm_count, int kind, int32_t* segs_out, int seg_cap,
                     int32_t* begin_out, int32_t info_out[3]) {
  return oryon::match::plan_debug(n_a, n_q, B, sm_count, kind, segs_out, seg_cap, begin_out, info_out);
}

int oryon_mask_to_roi(oryon_handle* h, const int32_t* mask, int B, int HW, int value, int32_t* roi_out, int32_t* n_out, void* stream) {
  return oryon::match::run_mask_to_roi(h, mask, B, HW, value, roi_out, n_out, static_cast<cudaStream_t>(stream));
}

int oryon_corrs_to_pcd(oryon_handle* h, const int64_t* corrs, int n, int feat_h, int feat_w, const void* depth_a, const void* depth_q,
                       int depth_dtype, int Ha, int Wa, int Hq, int Wq, const double* cam_a, const double* cam_q, float* pcd_a,
                       float* pcd_q, int32_t* n_valid, void* stream) {
  return oryon::lift::run_corrs_to_pcd(h, corrs, n, feat_h, feat_w, depth_a, depth_q, depth_dtype, Ha, Wa, Hq, Wq, cam_a, cam_q, pcd_a,
                                       pcd_q, n_valid, static_cast<cudaStream_t>(stream));
}

int oryon_select_lift(oryon_handle* h, const int32_t* rows, int B, int n, const int32_t* roi_a, const int32_t* roi_q, const int32_t* nn_idx,
                      int cap_a, int cap_q, int feat_h, int feat_w, const void* depth_a, const void* depth_q, int depth_dtype, int Ha, int Wa,
                      int Hq, int Wq, const double* cams_a, const double* cams_q, int64_t* corrs, float* pcd_a, float* pcd_q,
                      int32_t* n_valid, void* stream) {
  return oryon::lift::run_select_lift(h, rows, B, n, roi_a, roi_q, nn_idx, cap_a, cap_q, feat_h, feat_w, depth_a, depth_q, depth_dtype, Ha,
                                      Wa, Hq, Wq, cams_a, cams_q, corrs, pcd_a, pcd_q, n_valid, static_cast<cudaStream_t>(stream));
}

int oryon_lift_pcd(oryon_handle* h, const void* depth, int depth_dtype, int H, int W, const double* cam, const int64_t* xs,
                   const int64_t* ys, int n, float* out, void* stream) {
  return oryon::lift::run_lift(h, depth, depth_dtype, H, W, cam, xs, ys, n, out, static_cast<cudaStream_t>(stream));
}

int oryon_pointdsc_load(oryon_handle* h, const oryon_pointdsc_config* cfg, const float* weights, int64_t n_floats, void* stream) {
  return oryon::pdsc::load_weights(h, cfg, weights, n_floats, static_cast<cudaStream_t>(stream));
}

int oryon_pointdsc_pose(oryon_handle* h, const float* src, const float* tgt, const int32_t* n, int P, int cap, float* out_T,
                        const oryon_pointdsc_debug* debug, void* stream) {
  return oryon::pdsc::run_pose(h, src, tgt, n, P, cap, out_T, debug, static_cast<cudaStream_t>(stream));
}

int oryon_gemm_f32(oryon_handle* h, const float* A, const float* W, const float* bias, const float* residual, float* out, int M, int N,
                   int K, int batch, int act, float alpha, int precision, void* stream) {
  return oryon::gemm::run_gemm_f32(h, A, W, bias, residual, out, M, N, K, batch, act, alpha, precision, static_cast<cudaStream_t>(stream));
}

int oryon_backbone_set_weight(oryon_handle* h, const char* name, const float* data, int64_t numel) {
  return oryon::net::set_weight(h, name, data, numel);
}
int oryon_backbone_finalize(oryon_handle* h, const oryon_backbone_config* cfg, void* stream) {
  return oryon::net::finalize(h, cfg, static_cast<cudaStream_t>(stream));
}
int oryon_text_forward(oryon_handle* h, const int32_t* tokens, int n, float* out, void* stream) {
  return oryon::net::text_forward(h, tokens, n, out, static_cast<cudaStream_t>(stream));
}
int oryon_backbone_forward(oryon_handle* h, const float* rgb_a, const float* rgb_q, int B, const float* text_emb, float* featmap_a,
                           float* featmap_q, float* mask_a, float* mask_q, const oryon_backbone_debug* debug, void* stream) {
  return oryon::net::backbone_forward(h, rgb_a, rgb_q, B, text_emb, featmap_a, featmap_q, mask_a, mask_q, debug,
                                      static_cast<cudaStream_t>(stream));
}
int oryon_gemm_counters(oryon_handle* h, int64_t* launches, double* flops, double* tensor_flops) {
  ORYON_REQUIRE(h && launches && flops, "oryon_gemm_counters: null argument");
  *launches = h->gemm_launches, *flops = h->gemm_flops;
  if (tensor_flops) *tensor_flops = h->gemm_tensor_flops;
  h->gemm_launches = 0, h->gemm_flops = 0.0, h->gemm_tensor_flops = 0.0;
  return ORYON_OK;
}

int oryon_stage_inputs(oryon_handle* h, const uint8_t* rgb, const void* mask, int mask_is_i32, const int32_t* mask_ids, int B, int H, int W,
                       int out_h, int out_w, float* out_rgb, uint8_t* out_mask, void* stream) {
  return oryon::stage::run(h, rgb, mask, mask_is_i32, mask_ids, B, H, W, out_h, out_w, out_rgb, out_mask, static_cast<cudaStream_t>(stream));
}

int oryon_eval_set_object(oryon_handle* h, int obj_id, const double* pts, int n, const double* syms, int n_sym) {
  return oryon::eval::set_object(h, obj_id, pts, n, syms, n_sym);
}
int oryon_eval_pose_errors(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams,
                           double* out, void* stream) {
  return oryon::eval::pose_errors(h, P, obj_ids, pred, gt, cams, out, static_cast<cudaStream_t>(stream));
}

int oryon_eval_set_object_mesh(oryon_handle* h, int obj_id, const int32_t* faces, int n_faces) {
  return oryon::eval::set_object_mesh(h, obj_id, faces, n_faces);
}
int oryon_eval_vsd(oryon_handle* h, int P, const int32_t* obj_ids, const double* pred, const double* gt, const double* cams,
                   const void* depth_test, int depth_is_f32, int H, int W, double delta, const double* taus, int n_tau,
                   const double* diameters, double* out, void* stream) {
  return oryon::eval::vsd(h, P, obj_ids, pred, gt, cams, depth_test, depth_is_f32, H, W, delta, taus, n_tau, diameters, out,
                          static_cast<cudaStream_t>(stream));
}

int oryon_mask_postproc(oryon_handle* h, const float* logits, int B, int H, int W, float mask_th, const uint8_t* gt, int Hg, int Wg,
                        int32_t* pred_mask, int32_t* gt_resized, int32_t* n_pred, int32_t* n_gt, float* iou, void* stream) {
  return oryon::maskpost::run(h, logits, B, H, W, mask_th, gt, Hg, Wg, pred_mask, gt_resized, n_pred, n_gt, iou,
                              static_cast<cudaStream_t>(stream));
}

}  // extern "C"
