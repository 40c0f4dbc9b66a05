// Shared host-side plumbing for liboryon_b200: status/error reporting, the handle and its
// grow-only device workspace.  No torch, no third-party headers.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/oryon_b200.h"

namespace oryon {

void set_error(const char* fmt, ...);

#define ORYON_CUDA_CHECK(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess) {                                                                     \
      ::oryon::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ORYON_ERR_CUDA;                                                                     \
    }                                                                                            \
  } while (0)

#define ORYON_REQUIRE(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      ::oryon::set_error(__VA_ARGS__);                \
      return ORYON_ERR_INVALID_ARGUMENT;              \
    }                                                 \
  } while (0)

// One grow-only device buffer.  Growth happens on the caller's stream order: the old block is
// released with cudaFreeAsync-like semantics by synchronising the stream first (growth is rare:
// sizes are monotone per workload).
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t want, cudaStream_t stream);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(ptr); }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace oryon

namespace oryon {
// Optional per-kernel timing (oryon_profile_enable): CUDA events recorded on the caller's stream around
// each kernel launch, summed per kernel id when read.  Off by default (no events are recorded).
enum KernelId { KID_PREP = 0, KID_MATCH_TC = 1, KID_REFINE = 2, KID_EXACT = 3, KID_MASK_ROI = 4, KID_LIFT = 5, KID_PDSC_SC = 6, KID_PDSC_NET = 7, KID_PDSC_SEEDS = 8, KID_PDSC_REFINE = 9, KID_GEMM = 10, KID_ATTN = 11, KID_NORM = 12, KID_ELTWISE = 13, KID_IM2COL = 14, KID_ATTN_TC = 15, KID_TRANSPOSE = 16, KID_COUNT = 32 };
struct ProfSpan {
  int id;
  cudaEvent_t a, b;
};
}  // namespace oryon

// The opaque handle of the C ABI.
struct oryon_handle {
  bool profiling = false;
  int span_alias = -1;                      // >= 0: spans opened meanwhile are booked under this kernel id (PointDSC's GEMMs)
  bool gemm_uncounted = false;              // GEMMs launched meanwhile stay out of gemm_launches / gemm_flops (the network's counters)
  std::vector<oryon::ProfSpan> spans;       // recorded, not yet read
  std::vector<cudaEvent_t> free_events;
  cudaEvent_t take_event();
  void span_begin(int id, cudaStream_t st);
  void span_end(cudaStream_t st);

  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  oryon::PFN_encodeTiled encode_tiled = nullptr;

  // ---- matching workspace (see match.cu) ----
  oryon::DeviceBuffer rows16_a, rows16_q;   // fp16 unit rows, K-major, padded   [B][Npad][Dpad]
  oryon::DeviceBuffer rows32_a, rows32_q;   // fp32 unit rows                    [B][Npad][D]
  oryon::DeviceBuffer cand;                 // per (pair,row,split) candidate lists
  oryon::DeviceBuffer counters;             // small counter block (work queue, stats, overflow list size)
  oryon::DeviceBuffer overflow_rows;        // rows that need the exact fallback
  oryon::DeviceBuffer pair_meta;            // per-pair {n_a, n_q} on device
  oryon::DeviceBuffer match_plan;           // segment table of the tensor-core pass (match.cu: Seg)
  int64_t last_launches = 0;
  bool match_hist = false;                  // oryon_match_set_hist: the refine pass also records candidate-list length histograms

  // ---- lift workspace ----
  oryon::DeviceBuffer lift_scratch;

  // ---- PointDSC (see pointdsc.cu) ----
  void* pointdsc_model = nullptr;           // oryon::pdsc::Model
  oryon::DeviceBuffer pdsc_ws;

  // ---- backbone (see gemm.cu, backbone.cu) ----
  oryon::DeviceBuffer gemm_scratch;
  int64_t gemm_launches = 0;
  double gemm_flops = 0.0;
  double gemm_tensor_flops = 0.0;           // issued tensor-pipe work in fp16-equivalent FLOPs: 3x / 2x / 1x the algorithmic ones at precision 3 / 2 / 1
  void* backbone = nullptr;                 // oryon::net::Backbone

  // ---- evaluator (see eval.cu) ----
  void* eval_state = nullptr;               // oryon::eval::State

  int64_t workspace_bytes() const;
};
