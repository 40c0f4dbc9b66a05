// Fused softmax attention for long sequences on tcgen05 (CLIP vision tower, S = 577, d = 64; reference call site models/vlm.py:54 ->
// nn.MultiheadAttention inside clip's ResidualAttentionBlock).  Three kernels, in the order they were written; launch() picks
//   attn_pp_kernel      (default at three products) 256 queries per CTA, two softmax groups out of phase, P over S in place
//   attn_online_kernel  (ORYON_ATTN_LOCKSTEP=1, and precision 1) one pass over the key tiles, lazily renewed reference maximum
//   attn_tc_kernel      (ORYON_ATTN_TWOPASS) two passes, described first:
//
// attn_tc_kernel<NPASS>:
// One CTA = 128 queries of one (sequence, head).  Scores never leave the SM:
//   warp 0      TMA producer   Q tile once; K tiles and V^T tiles through 2-stage rings (both requested two tiles ahead:
//                              a single V^T stage exposed one full TMA latency per key tile, 37k -> cycles per CTA)
//   warp 1      MMA issuer     S = Q K^T (M=128, N=128, K=64) into a double-buffered TMEM accumulator;
//                              O += P V (M=128, N=64, K=128) into a third TMEM region
//   warps 2..17 softmax        thread <-> query row (TMEM lane); four warps per lane quarter, each owning 32 of the 128
//                              key columns of a tile.  Pass 1 over the key tiles: row maximum (a single hi*hi product
//                              is enough: softmax is invariant to the subtracted constant).  Pass 2: S is recomputed
//                              (K = 64: cheap), p = exp2((s - max) * scale * log2 e), the row sum accumulates in a
//                              register, p is held in registers as packed fp16 split pairs until the P V MMA of the
//                              previous tile has released the buffer, then written with tcgen05.st into TENSOR MEMORY:
//                              P is the A operand of the P V MMA and tcgen05.mma takes A from tensor memory (the ".ts"
//                              form: lane = query row, two fp16 per 32-bit column).  Round 1 wrote P into a swizzled
//                              shared-memory tile instead: 16 st.shared.v4 per thread and tile plus a fence.proxy.async,
//                              ~1 000-1 500 of the ~2 900 cycles a key tile took (profiles/r02_attn_tc.md); the 64 KB that
//                              buffer occupied now hold a third K stage.  Finally O / sum is stored as a split pair into
//                              the concatenated-heads activation.
// Two passes instead of an online rescale keep a max-subtracted softmax and need no TMEM read-modify-write of O; the
// extra Q K^T costs little tensor work on a kernel that is bound by the exp / convert work of the softmax warps.
// Operands are fp16 split pairs; NPASS = 3 evaluates hi*hi + lo*hi + hi*lo (see gemm.cuh).
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "gemm.cuh"
#include "ptx_sm100.cuh"

namespace oryon {
namespace attn {

constexpr int kQ = 128;      // queries per CTA
constexpr int kKT = 128;     // keys per tile
constexpr int kD = 64;       // head dim
// Softmax warps: 16 = four per TMEM lane quarter, each owning 32 of the 128 key columns of a tile.  Round 1 ran 8 (64 columns per
// thread): ncu showed the kernel bound by the softmax warps with the issue slots 32 % used -- two warps per scheduler cannot hide
// the tensor-memory loads, MUFU latency and barrier waits of one another.  With four per scheduler (and half the registers per
// thread) the same instruction stream overlaps better; see profiles/r02_attn_tc.md for the measured effect.
constexpr int kSmWarps = 16;
constexpr int kParts = kSmWarps / 4;            // column groups per tile
constexpr int kPartCols = 128 / kParts;         // 32 key columns per thread and tile
constexpr int kThreads = 64 + 32 * kSmWarps;    // TMA, MMA, softmax warps
constexpr int kTileBytes = 128 * 64 * 2;   // one [128][64] fp16 tile, 128-byte rows, SWIZZLE_128B

constexpr int kNK = 3;       // K stages (ring); S accumulators and V stages stay double buffered

template <int NPASS>
struct Cfg {
  static constexpr int kHalves = NPASS == 3 ? 2 : 1;
  static constexpr int kQBytes = kHalves * kTileBytes;
  static constexpr int kKStage = kHalves * kTileBytes;
  static constexpr int kVBytes = kHalves * kTileBytes;        // [64 d][128 keys] = two [64][64] sub-tiles per half
  static constexpr int kOffK = kQBytes;
  static constexpr int kOffV = kOffK + kNK * kKStage;
  static constexpr int kOffBar = kOffV + 2 * kVBytes;
  static constexpr int kOffRed = kOffBar + 256;               // [kParts][128] floats: cross-warp max / sum exchange
  static constexpr int kTotal = 1024 + kOffRed + kParts * 128 * 4;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

// tensor-memory columns (512 allocated): S accumulators [0,256), O [256,320), P hi [320,384), P lo [384,448)
constexpr uint32_t kTmemO = 256, kTmemPhi = 320, kTmemPlo = 384;

struct Args {
  int S, heads, T;            // T = key tiles
  float scale_log2e;          // head_dim^-0.5 * log2(e)
  float tau;                  // online kernel: the reference maximum of a row is renewed when a tile exceeds it by more than 2^tau
  __half* out_hi;
  __half* out_lo;
  int64_t ldh;
  int lo_format;              // gemm::LO_F8X: out_lo receives the 8-bit cross-term blocks of the rows (the out-projection runs at precision 2)
  int qk_f8x;                 // ping-pong kernel: the Q and K columns of the `lo` matrix hold 8-bit cross-term blocks (gemm::LO_QKV): Q K^T = hi*hi
                              // on fp16 + both cross terms as one e5m2 x e5m2 product (8 tensor instructions per score tile instead of 12)
  long long* dbg;             // optional per-phase clock64 stamps of CTA (0,0,0) (ORYON_ATTN_DEBUG)
};

__device__ __forceinline__ void tma_load_2d_(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) { ptx::tma_load_2d(dst, m, bar, c0, c1); }

__device__ __forceinline__ void split_half(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}

// VMN = true: V tiles are fetched straight from the QKV buffer ([128 keys][64 d], the same box as a K tile) and fed to the
// P V MMA as an MN-major B operand; VMN = false: from a pre-transposed V^T buffer as a K-major operand (transpose_v_kernel).
template <int NPASS, bool VMN>
__global__ void __launch_bounds__(kThreads, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi, const __grid_constant__ CUtensorMap tm_qkv_lo,
               const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo, Args a, int width) {
  using L = Cfg<NPASS>;
  constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0);
  constexpr uint32_t kIdescO = ptx::make_idesc_f16(128, 64, 0, VMN ? 1 : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + L::kOffK;
  uint8_t* sV = smem + L::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [kNK]
  uint64_t* k_empty = k_full + kNK;        // [kNK]
  uint64_t* v_full = k_empty + kNK;        // [2]
  uint64_t* v_empty = v_full + 2;          // [2]
  uint64_t* s_full = v_empty + 2;          // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  static_assert(8 * (4 + 2 * kNK + 8) + 4 <= 256, "barrier block");

  const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ, head = blockIdx.y, seq = blockIdx.z;
  const int T = a.T;
  const bool dbg = a.dbg && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 1;
#define STAMP(slot) do { if (dbg && lane == 0) a.dbg[slot] = clock64(); } while (0)
  if (dbg && threadIdx.x == 0) a.dbg[0] = clock64();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv_hi);
    if (!VMN) ptx::prefetch_tensormap(&tm_vt_hi);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kNK; ++i) ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1), ptx::mbar_init(&s_empty[i], kSmWarps);
      ptx::mbar_init(&v_full[i], 1), ptx::mbar_init(&v_empty[i], 1);
    }
    ptx::mbar_init(p_full, kSmWarps), ptx::mbar_init(p_empty, 1), ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + kTmemO;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = seq * a.S;
      ptx::mbar_arrive_expect_tx(q_full, L::kQBytes);
      tma_load_2d_(sQ, &tm_qkv_hi, q_full, head * kD, row0 + q0);
      if (NPASS == 3) tma_load_2d_(sQ + kTileBytes, &tm_qkv_lo, q_full, head * kD, row0 + q0);
      for (int i = 0; i < 2 * T; ++i) {
        const int st = i % kNK, kt = i % T;
        ptx::mbar_wait(&k_empty[st], ((i / kNK) & 1) ^ 1);
        const bool need_lo = NPASS == 3 && i >= T;   // pass 1 (row maximum) multiplies the hi halves only
        ptx::mbar_arrive_expect_tx(&k_full[st], need_lo ? L::kKStage : kTileBytes);
        tma_load_2d_(sK + st * L::kKStage, &tm_qkv_hi, &k_full[st], width + head * kD, row0 + kt * kKT);
        if (need_lo) tma_load_2d_(sK + st * L::kKStage + kTileBytes, &tm_qkv_lo, &k_full[st], width + head * kD, row0 + kt * kKT);
        if (i >= T) {
          const int j = i - T, vs = j & 1;
          ptx::mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&v_full[vs], L::kVBytes);
          uint8_t* dst = sV + vs * L::kVBytes;
          if (VMN) {   // rows past S belong to the next sequence (or are zero-filled past the end): their probabilities are exactly 0
            tma_load_2d_(dst, &tm_qkv_hi, &v_full[vs], 2 * width + head * kD, row0 + j * kKT);
            if (NPASS == 3) tma_load_2d_(dst + kTileBytes, &tm_qkv_lo, &v_full[vs], 2 * width + head * kD, row0 + j * kKT);
          } else {
            const int vrow = (seq * a.heads + head) * kD;
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              tma_load_2d_(dst + sub * (kTileBytes / 2), &tm_vt_hi, &v_full[vs], j * kKT + sub * 64, vrow);
              if (NPASS == 3) tma_load_2d_(dst + kTileBytes + sub * (kTileBytes / 2), &tm_vt_lo, &v_full[vs], j * kKT + sub * 64, vrow);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    auto issue_pv = [&](int j) {
      ptx::mbar_wait(p_full, j & 1);
      STAMP(50 + j);
      ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
      STAMP(60 + j);
      ptx::tc_fence_after();
      {
        // warp-uniform descriptors, one elected lane issues (see gemm.cu): the 8 * NPASS MMAs go out back to back.
        // A = P from tensor memory: [128 query lanes][128 keys] fp16, two keys per 32-bit column -> 8 columns per 16-key MMA.
        const uint32_t v_addr = ptx::smem_u32(sV + (j & 1) * L::kVBytes);
        const bool leader = ptx::elect_one();
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          // pass 0: P_hi V_hi   pass 1: P_lo V_hi   pass 2: P_hi V_lo
          const uint32_t tp = tmem_base + (pass == 1 ? kTmemPlo : kTmemPhi);
          const uint32_t va = v_addr + (pass == 2 ? kTileBytes : 0);
          const uint64_t dv0 = VMN ? ptx::make_smem_desc_mnmajor_sw128(va, 0, 1024) : ptx::make_smem_desc_kmajor(va, 128);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // 16 keys per MMA: MN-major V advances two 8-key groups (2 KB); K-major V^T 32 bytes inside the swizzled row.
              const uint64_t dv = dv0 + (VMN ? (sub * 4 + k) * (2048 >> 4) : sub * (kTileBytes >> 5) + 2 * k);
              if (leader) ptx::umma_f16_ts(tmem_o, tp + (sub * 4 + k) * 8, dv, kIdescO, (j | pass | sub | k) != 0 ? 1u : 0u);
            }
        }
        if (leader) {
          ptx::umma_commit(&v_empty[j & 1]);
          ptx::umma_commit(p_empty);
        }
      }
      __syncwarp();
    };
    ptx::mbar_wait(q_full, 0);
    STAMP(1);
    for (int i = 0; i < 2 * T; ++i) {
      const int st = i & 1, ks = i % kNK;   // S accumulator / K stage
      ptx::mbar_wait(&k_full[ks], (i / kNK) & 1);
      STAMP(10 + i);
      ptx::mbar_wait(&s_empty[st], ((i >> 1) & 1) ^ 1);
      STAMP(30 + i);
      ptx::tc_fence_after();
      {
        const uint32_t q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK + ks * L::kKStage);
        const int npass = i < T ? 1 : NPASS;   // pass 1 only needs an approximate maximum
        const bool leader = ptx::elect_one();
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          if (pass >= npass) break;
          const uint64_t dq0 = ptx::make_smem_desc_kmajor(q_addr + (pass == 1 ? kTileBytes : 0), 128);
          const uint64_t dk0 = ptx::make_smem_desc_kmajor(k_addr + (pass == 2 ? kTileBytes : 0), 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (leader) ptx::umma_f16(tmem_base + st * 128, dq0 + 2 * k, dk0 + 2 * k, kIdescS, (pass | k) != 0 ? 1u : 0u);
        }
        if (leader) {
          ptx::umma_commit(&k_empty[ks]);
          ptx::umma_commit(&s_full[st]);
        }
      }
      __syncwarp();
      if (i > T) issue_pv(i - 1 - T);
    }
    issue_pv(T - 1);
    if (ptx::elect_one()) ptx::umma_commit(o_full);
    __syncwarp();
  } else {
    const int quarter = warp & 3;                       // TMEM lanes 32*quarter .. +31
    const int part = (warp - 2) >> 2;                   // which kPartCols key columns of every tile this warp owns
    const int r = quarter * 32 + lane;                  // row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float* red = reinterpret_cast<float*>(smem + L::kOffRed);   // [kParts][128]
    float m = -INFINITY;
    const bool dbg2 = dbg && warp == 2;
    static_assert(kPartCols == 32, "one 32-column tensor-memory load per thread and tile");
    // ---- pass 1: row maximum ----
    for (int i = 0; i < T; ++i) {
      const int st = i & 1;
      ptx::mbar_wait(&s_full[st], (i >> 1) & 1);
      if (dbg2 && lane == 0) a.dbg[70 + i] = clock64();
      ptx::tc_fence_after();
      const int key0 = i * kKT + part * kPartCols;
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + st * 128 + part * kPartCols, v);
      ptx::tmem_ld_wait();
      if (key0 + 32 <= a.S) {
#pragma unroll
        for (int c = 0; c < 32; ++c) m = fmaxf(m, __uint_as_float(v[c]));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (key0 + c < a.S) m = fmaxf(m, __uint_as_float(v[c]));
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[st]);
    }
    red[part * 128 + r] = m;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
#pragma unroll
    for (int pp = 0; pp < kParts; ++pp) m = fmaxf(m, red[pp * 128 + r]);
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    // ---- pass 2: probabilities -> registers -> shared memory (swizzled K-major), row sum ----
    float l = 0.f;
    const float ms = m * a.scale_log2e;
    // P in tensor memory: this thread's row is lane r; its 32 keys are the 16 packed columns [part * 16, +16) of the tile
    const uint32_t tp_hi = tmem_base + lane_addr + kTmemPhi + part * (kPartCols / 2);
    const uint32_t tp_lo = tmem_base + lane_addr + kTmemPlo + part * (kPartCols / 2);
    for (int j = 0; j < T; ++j) {
      const int i = T + j, st = i & 1;
      ptx::mbar_wait(&s_full[st], (i >> 1) & 1);
      if (dbg2 && lane == 0) a.dbg[70 + i] = clock64();
      ptx::tc_fence_after();
      const int key0 = j * kKT + part * kPartCols;
      uint32_t ph[16], pl[16];   // packed half2: element 2c, 2c+1 of this thread's 32 columns
      {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + st * 128 + part * kPartCols, v);
        ptx::tmem_ld_wait();
        auto convert = [&](auto masked) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float p0, p1;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(__uint_as_float(v[2 * c]), a.scale_log2e, -ms)));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(__uint_as_float(v[2 * c + 1]), a.scale_log2e, -ms)));
            if (decltype(masked)::value) {   // only the tile that straddles the end of the sequence pays for the masking
              if (key0 + 2 * c >= a.S) p0 = 0.f;
              if (key0 + 2 * c + 1 >= a.S) p1 = 0.f;
            }
            l += p0 + p1;
            const __half2 h2 = __floats2half2_rn(p0, p1);
            const float2 back = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(p0 - back.x, p1 - back.y);
            ph[c] = *reinterpret_cast<const uint32_t*>(&h2);
            pl[c] = *reinterpret_cast<const uint32_t*>(&l2);
          }
        };
        if (key0 + 32 <= a.S) convert(std::false_type{});
        else convert(std::true_type{});
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[st]);              // S buffer drained: the next Q K^T may overwrite it
      if (dbg2 && lane == 0) a.dbg[90 + j] = clock64();
      if (j > 0) ptx::mbar_wait(p_empty, (j - 1) & 1);             // P V of the previous tile has consumed the buffer
      if (dbg2 && lane == 0) a.dbg[100 + j] = clock64();
      ptx::tc_fence_after();
      ptx::tmem_st_32x32b_x16(tp_hi, ph);
      if (NPASS == 3) ptx::tmem_st_32x32b_x16(tp_lo, pl);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    red[part * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    l = 0.f;
#pragma unroll
    for (int pp = 0; pp < kParts; ++pp) l += red[pp * 128 + r];
    // ---- output: O / l -> split pair, concatenated heads; this warp stores columns part*16 .. +15 ----
    ptx::mbar_wait(o_full, 0);
    if (dbg2 && lane == 0) a.dbg[110] = clock64();
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    const int q = q0 + r;
    {
      constexpr int kOutCols = kD / kParts;   // 16
      static_assert(kOutCols == 16, "one 16-column tensor-memory load per thread");
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(tmem_o + lane_addr + part * kOutCols, v);
      ptx::tmem_ld_wait();
      if (q < a.S) {
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float x0 = __uint_as_float(v[2 * c]) * inv, x1 = __uint_as_float(v[2 * c + 1]) * inv;
          const __half2 h2 = __floats2half2_rn(x0, x1);
          const float2 back = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(x0 - back.x, x1 - back.y);
          oh[c] = *reinterpret_cast<const uint32_t*>(&h2), ol[c] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const int64_t o = ((int64_t)seq * a.S + q) * a.ldh + head * kD + part * kOutCols;
#pragma unroll
        for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_hi + o)[c] = make_uint4(oh[4 * c], oh[4 * c + 1], oh[4 * c + 2], oh[4 * c + 3]);
        if (a.lo_format == gemm::LO_F8X) {
          // 16 consecutive columns of one 64-column head: 16 bytes in each half of the row's 128-byte cross-term block
          uint32_t f[4], g[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            gemm::f8x_act4(__uint_as_float(v[4 * c]) * inv, __uint_as_float(v[4 * c + 1]) * inv, __uint_as_float(v[4 * c + 2]) * inv,
                           __uint_as_float(v[4 * c + 3]) * inv, f[c], g[c]);
          uint8_t* pb = reinterpret_cast<uint8_t*>(a.out_lo + ((int64_t)seq * a.S + q) * a.ldh) + gemm::f8x_off(head * kD + part * kOutCols);
          *reinterpret_cast<uint4*>(pb) = make_uint4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<uint4*>(pb + 64) = make_uint4(g[0], g[1], g[2], g[3]);
        } else if (a.out_lo) {
#pragma unroll
          for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_lo + o)[c] = make_uint4(ol[4 * c], ol[4 * c + 1], ol[4 * c + 2], ol[4 * c + 3]);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) a.dbg[111] = clock64();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// attn_online_kernel: the same attention in ONE pass over the key tiles (online softmax with a lazily renewed reference maximum).
// ------------------------------------------------------------------------------------------------
// The two-pass kernel above spends ~3 k of its ~19 k cycles per CTA on pass 1 (row maxima from an extra hi*hi Q K^T per tile, an
// extra read-out of S, an extra K stream through L2).  Here every key tile is visited once: S = Q K^T in three products, the four
// column warps of a row exchange the tile's row maximum through shared memory (one named barrier per lane quarter and tile), and
// p = 2^(s*c - m_ref) is taken against a REFERENCE maximum m_ref that is only renewed when a tile exceeds it by more than tau (2^8):
// softmax is invariant to the subtracted constant, p <= 2^tau keeps the fp16 split pair exact to 2^-21, and the accumulated O and
// row sum need a correction (x 2^(m_old - m_new)) only on a renewal -- after the first tile that is rare.  The correction of O
// (tcgen05.ld, multiply, tcgen05.st of this warp's 32 rows x 16 columns) happens between the wait for the previous P V product and
// the store of the new P, when no MMA touches O.  Same operands, same three-product P V through tensor memory as above.
constexpr int kRedFloats = 3 * kParts * 128;   // two alternating row-maximum exchange buffers + one for the row sums

template <int NPASS>
struct CfgOnline {
  static constexpr int kHalves = NPASS == 3 ? 2 : 1;
  static constexpr int kQBytes = kHalves * kTileBytes;
  static constexpr int kKStage = kHalves * kTileBytes;
  static constexpr int kVBytes = kHalves * kTileBytes;
  static constexpr int kOffK = kQBytes;
  static constexpr int kOffV = kOffK + kNK * kKStage;
  static constexpr int kOffBar = kOffV + 2 * kVBytes;
  static constexpr int kOffRed = kOffBar + 256;
  static constexpr int kTotal = 1024 + kOffRed + kRedFloats * 4;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <int NPASS, bool VMN>
__global__ void __launch_bounds__(kThreads, 1)
attn_online_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi, const __grid_constant__ CUtensorMap tm_qkv_lo,
                   const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo, Args a, int width) {
  using L = CfgOnline<NPASS>;
  constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0);
  constexpr uint32_t kIdescO = ptx::make_idesc_f16(128, 64, 0, VMN ? 1 : 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + L::kOffK;
  uint8_t* sV = smem + L::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [kNK]
  uint64_t* k_empty = k_full + kNK;        // [kNK]
  uint64_t* v_full = k_empty + kNK;        // [2]
  uint64_t* v_empty = v_full + 2;          // [2]
  uint64_t* s_full = v_empty + 2;          // [2]
  uint64_t* s_empty = s_full + 2;          // [2]
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 1;
  uint64_t* o_full = p_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQ, head = blockIdx.y, seq = blockIdx.z;
  const int T = a.T;
  const bool dbg = a.dbg && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 1;
  if (dbg && threadIdx.x == 0) a.dbg[0] = clock64();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv_hi);
    if (!VMN) ptx::prefetch_tensormap(&tm_vt_hi);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kNK; ++i) ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1), ptx::mbar_init(&s_empty[i], kSmWarps);
      ptx::mbar_init(&v_full[i], 1), ptx::mbar_init(&v_empty[i], 1);
    }
    ptx::mbar_init(p_full, kSmWarps), ptx::mbar_init(p_empty, 1), ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + kTmemO;

  if (warp == 0) {
    if (lane == 0) {
      const int row0 = seq * a.S;
      ptx::mbar_arrive_expect_tx(q_full, L::kQBytes);
      tma_load_2d_(sQ, &tm_qkv_hi, q_full, head * kD, row0 + q0);
      if (NPASS == 3) tma_load_2d_(sQ + kTileBytes, &tm_qkv_lo, q_full, head * kD, row0 + q0);
      for (int j = 0; j < T; ++j) {
        const int st = j % kNK, vs = j & 1;
        ptx::mbar_wait(&k_empty[st], ((j / kNK) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], L::kKStage);
        tma_load_2d_(sK + st * L::kKStage, &tm_qkv_hi, &k_full[st], width + head * kD, row0 + j * kKT);
        if (NPASS == 3) tma_load_2d_(sK + st * L::kKStage + kTileBytes, &tm_qkv_lo, &k_full[st], width + head * kD, row0 + j * kKT);
        ptx::mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[vs], L::kVBytes);
        uint8_t* dst = sV + vs * L::kVBytes;
        if (VMN) {   // rows past S belong to the next sequence (or are zero-filled past the end): their probabilities are exactly 0
          tma_load_2d_(dst, &tm_qkv_hi, &v_full[vs], 2 * width + head * kD, row0 + j * kKT);
          if (NPASS == 3) tma_load_2d_(dst + kTileBytes, &tm_qkv_lo, &v_full[vs], 2 * width + head * kD, row0 + j * kKT);
        } else {
          const int vrow = (seq * a.heads + head) * kD;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            tma_load_2d_(dst + sub * (kTileBytes / 2), &tm_vt_hi, &v_full[vs], j * kKT + sub * 64, vrow);
            if (NPASS == 3) tma_load_2d_(dst + kTileBytes + sub * (kTileBytes / 2), &tm_vt_lo, &v_full[vs], j * kKT + sub * 64, vrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    auto issue_pv = [&](int j) {
      ptx::mbar_wait(p_full, j & 1);
      if (dbg && lane == 0) a.dbg[50 + j] = clock64();
      ptx::mbar_wait(&v_full[j & 1], (j >> 1) & 1);
      ptx::tc_fence_after();
      {
        const uint32_t v_addr = ptx::smem_u32(sV + (j & 1) * L::kVBytes);
        const bool leader = ptx::elect_one();
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          // pass 0: P_hi V_hi   pass 1: P_lo V_hi   pass 2: P_hi V_lo
          const uint32_t tp = tmem_base + (pass == 1 ? kTmemPlo : kTmemPhi);
          const uint32_t va = v_addr + (pass == 2 ? kTileBytes : 0);
          const uint64_t dv0 = VMN ? ptx::make_smem_desc_mnmajor_sw128(va, 0, 1024) : ptx::make_smem_desc_kmajor(va, 128);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t dv = dv0 + (VMN ? (sub * 4 + k) * (2048 >> 4) : sub * (kTileBytes >> 5) + 2 * k);
              if (leader) ptx::umma_f16_ts(tmem_o, tp + (sub * 4 + k) * 8, dv, kIdescO, (j | pass | sub | k) != 0 ? 1u : 0u);
            }
        }
        if (leader) {
          ptx::umma_commit(&v_empty[j & 1]);
          ptx::umma_commit(p_empty);
        }
      }
      __syncwarp();
    };
    ptx::mbar_wait(q_full, 0);
    if (dbg && lane == 0) a.dbg[1] = clock64();
    for (int j = 0; j < T; ++j) {
      const int sb = j & 1, ks = j % kNK;
      ptx::mbar_wait(&k_full[ks], (j / kNK) & 1);
      if (dbg && lane == 0) a.dbg[10 + j] = clock64();
      ptx::mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      {
        const uint32_t q_addr = ptx::smem_u32(sQ), k_addr = ptx::smem_u32(sK + ks * L::kKStage);
        const bool leader = ptx::elect_one();
#pragma unroll
        for (int pass = 0; pass < NPASS; ++pass) {
          const uint64_t dq0 = ptx::make_smem_desc_kmajor(q_addr + (pass == 1 ? kTileBytes : 0), 128);
          const uint64_t dk0 = ptx::make_smem_desc_kmajor(k_addr + (pass == 2 ? kTileBytes : 0), 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (leader) ptx::umma_f16(tmem_base + sb * 128, dq0 + 2 * k, dk0 + 2 * k, kIdescS, (pass | k) != 0 ? 1u : 0u);
        }
        if (leader) {
          ptx::umma_commit(&k_empty[ks]);
          ptx::umma_commit(&s_full[sb]);
        }
      }
      __syncwarp();
      if (j > 0) issue_pv(j - 1);
    }
    issue_pv(T - 1);
    if (ptx::elect_one()) ptx::umma_commit(o_full);
    __syncwarp();
  } else {
    const int quarter = warp & 3;                       // TMEM lanes 32*quarter .. +31
    const int part = (warp - 2) >> 2;                   // which 32 key columns of every tile this warp owns
    const int r = quarter * 32 + lane;                  // row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float* red = reinterpret_cast<float*>(smem + L::kOffRed);   // [3][kParts][128]
    const uint32_t tp_hi = tmem_base + lane_addr + kTmemPhi + part * (kPartCols / 2);
    const uint32_t tp_lo = tmem_base + lane_addr + kTmemPlo + part * (kPartCols / 2);
    const uint32_t to = tmem_o + lane_addr + part * (kD / kParts);
    const bool dbg2 = dbg && warp == 2;
    float m_ref = -INFINITY;   // reference maximum of this row, in the scaled log2 domain (s * scale_log2e)
    float l = 0.f;             // this thread's share of the row sum, relative to m_ref
    for (int j = 0; j < T; ++j) {
      const int sb = j & 1;
      ptx::mbar_wait(&s_full[sb], (j >> 1) & 1);
      if (dbg2 && lane == 0) a.dbg[70 + j] = clock64();
      ptx::tc_fence_after();
      const int key0 = j * kKT + part * kPartCols;
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + lane_addr + sb * 128 + part * kPartCols, v);
      ptx::tmem_ld_wait();
      const bool full = key0 + 32 <= a.S;
      float m_loc = -INFINITY;
      if (full) {
#pragma unroll
        for (int c = 0; c < 32; ++c) m_loc = fmaxf(m_loc, __uint_as_float(v[c]));
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (key0 + c < a.S) m_loc = fmaxf(m_loc, __uint_as_float(v[c]));
      }
      // the tile's row maximum over the four column warps of this lane quarter (alternating buffers: no second barrier)
      float* ex = red + (j & 1) * (kParts * 128);
      ex[part * 128 + r] = m_loc;
      asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
      float m_tile = ex[r];
#pragma unroll
      for (int pp = 1; pp < kParts; ++pp) m_tile = fmaxf(m_tile, ex[pp * 128 + r]);
      const float mt = m_tile * a.scale_log2e;
      const bool grow = mt > m_ref + a.tau;     // always on the first tile (m_ref = -inf); the same decision in all four warps of the row
      float alpha = 1.f;
      if (grow) {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(m_ref - mt));   // 0 on the first tile
        m_ref = mt;
        l *= alpha;
      }
      uint32_t ph[16], pl[16];   // packed half2: element 2c, 2c+1 of this thread's 32 columns
      {
        // packed fp32x2 arithmetic (FFMA2 / FADD2): one instruction for the scale-and-shift of two scores, one for the two row-sum
        // additions, one for the two fp16 residuals -- 9 instead of 12 issue slots per pair of scores on warps that are issue-bound
        const uint64_t c2 = ptx::pack_f32x2(a.scale_log2e, a.scale_log2e), nm2 = ptx::pack_f32x2(-m_ref, -m_ref);
        const uint64_t neg1 = ptx::pack_f32x2(-1.f, -1.f);
        uint64_t l2 = ptx::pack_f32x2(l, 0.f);
        auto convert = [&](auto masked) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float x0, x1, p0, p1;
            ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), c2, nm2), x0, x1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
            if (decltype(masked)::value) {
              if (key0 + 2 * c >= a.S) p0 = 0.f;
              if (key0 + 2 * c + 1 >= a.S) p1 = 0.f;
            }
            const uint64_t pp = ptx::pack_f32x2(p0, p1);
            l2 = ptx::add_f32x2(l2, pp);
            const __half2 h2 = __floats2half2_rn(p0, p1);
            const float2 back = __half22float2(h2);
            float r0, r1;
            ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(back.x, back.y), neg1, pp), r0, r1);   // p - fp16(p), exact
            const __half2 lo2 = __floats2half2_rn(r0, r1);
            ph[c] = *reinterpret_cast<const uint32_t*>(&h2);
            pl[c] = *reinterpret_cast<const uint32_t*>(&lo2);
          }
        };
        if (full) convert(std::false_type{});
        else convert(std::true_type{});
        float la, lb;
        ptx::unpack_f32x2(l2, la, lb);
        l = la + lb;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_empty[sb]);              // S buffer drained: the Q K^T after next may overwrite it
      if (dbg2 && lane == 0) a.dbg[90 + j] = clock64();
      if (j > 0) {
        ptx::mbar_wait(p_empty, (j - 1) & 1);                      // P V of the previous tile is complete: P is free, O is at rest
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {                       // renew the reference of O: this warp's 32 rows x 16 columns
          uint32_t o[16];
          ptx::tmem_ld_32x32b_x16(to, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          ptx::tmem_st_32x32b_x16(to, o);
        }
      }
      if (dbg2 && lane == 0) a.dbg[100 + j] = clock64();
      ptx::tmem_st_32x32b_x16(tp_hi, ph);
      if (NPASS == 3) ptx::tmem_st_32x32b_x16(tp_lo, pl);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    float* exl = red + 2 * (kParts * 128);
    exl[part * 128 + r] = l;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    l = 0.f;
#pragma unroll
    for (int pp = 0; pp < kParts; ++pp) l += exl[pp * 128 + r];
    // ---- output: O / l -> split pair, concatenated heads; this warp stores columns part*16 .. +15 ----
    ptx::mbar_wait(o_full, 0);
    if (dbg2 && lane == 0) a.dbg[110] = clock64();
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    const int q = q0 + r;
    {
      constexpr int kOutCols = kD / kParts;   // 16
      uint32_t v[16];
      ptx::tmem_ld_32x32b_x16(to, v);
      ptx::tmem_ld_wait();
      if (q < a.S) {
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float x0 = __uint_as_float(v[2 * c]) * inv, x1 = __uint_as_float(v[2 * c + 1]) * inv;
          const __half2 h2 = __floats2half2_rn(x0, x1);
          const float2 back = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(x0 - back.x, x1 - back.y);
          oh[c] = *reinterpret_cast<const uint32_t*>(&h2), ol[c] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const int64_t o = ((int64_t)seq * a.S + q) * a.ldh + head * kD + part * kOutCols;
#pragma unroll
        for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_hi + o)[c] = make_uint4(oh[4 * c], oh[4 * c + 1], oh[4 * c + 2], oh[4 * c + 3]);
        if (a.lo_format == gemm::LO_F8X) {
          // 16 consecutive columns of one 64-column head: 16 bytes in each half of the row's 128-byte cross-term block
          uint32_t f[4], g[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            gemm::f8x_act4(__uint_as_float(v[4 * c]) * inv, __uint_as_float(v[4 * c + 1]) * inv, __uint_as_float(v[4 * c + 2]) * inv,
                           __uint_as_float(v[4 * c + 3]) * inv, f[c], g[c]);
          uint8_t* pb = reinterpret_cast<uint8_t*>(a.out_lo + ((int64_t)seq * a.S + q) * a.ldh) + gemm::f8x_off(head * kD + part * kOutCols);
          *reinterpret_cast<uint4*>(pb) = make_uint4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<uint4*>(pb + 64) = make_uint4(g[0], g[1], g[2], g[3]);
        } else if (a.out_lo) {
#pragma unroll
          for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_lo + o)[c] = make_uint4(ol[4 * c], ol[4 * c + 1], ol[4 * c + 2], ol[4 * c + 3]);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) a.dbg[111] = clock64();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------
// attn_pp_kernel: two softmax warp groups working out of phase ("ping-pong"), 256 queries per CTA, three products
// ------------------------------------------------------------------------------------------------
// What the one-group kernels above are bound by (clock stamps of one CTA, profiles/r02_attn_pp.md): (1) all 16 softmax warps walk
// through read-out -> maximum -> exchange -> exponentials -> store in lockstep, so the MUFU (16 384 ex2 per 128 x 128 tile = 1 024
// cycles at 16 per clock) idles half of every ~2 100-cycle tile; (2) every CTA streams all K and V tiles of its (sequence, head) for
// only 128 queries: 352 KB per CTA, 43 GB per 32-pair step at the ~8.5 TB/s the L2 delivers -- the loads of a tile take ~3 000 cycles
// to land and the first scores of a CTA wait ~3 300 cycles for Q and K.
// Here a CTA owns TWO query tiles (K / V tiles fetched once per 256 queries) and its 16 softmax warps form two groups of 8 (thread <->
// one row x 64 key columns).  The work items are (key tile kt, query tile g) in the order i = 2 kt + g; group g takes the items of its
// own query tile, with its own running reference maximum, row sums and O accumulator, so the groups never wait for each other and one
// group's exponentials run under the other's read-out / exchange / store phases.  The last CTA of a 577-token sequence has a single
// query tile: there the groups alternate over KEY tiles instead (item i = key tile i, group i mod 2, both on the same rows) and the
// two partial results (O_g, l_g, m_g) are merged at the end like a split-KV reduction.
// Tensor memory (512 columns): three rotating S buffers of 128 columns (item i -> buffer i mod 3: the scores of item i + 3 are issued
// right behind P V of item i and a group finds its next scores waiting), O_0 and O_1 at 384 / 448.  P overwrites S IN PLACE: a thread
// reads 32 of its fp32 scores and stores their 16 packed fp16 hi columns + 16 lo columns into the same 32 columns; the `.ts` MMA
// takes each 16-key K step from where it lies (hi at column 32 (k / 2) + 8 (k mod 2), lo 16 further).  No s_empty / p_empty
// barriers: S(i + 3) is ordered behind P V(i) by the in-order tensor pipe.
// Each thread reads its scores twice (maximum pass, then the exponential pass in two halves of 32 columns): tensor-memory read
// bandwidth is abundant (tools/tmem_ld_bw.cu) and the second read keeps the live registers at ~64 + 32.
// Shared memory: [Q_0][Q_1 | third K stage of the single-query-tile mode][K stage 0][K stage 1][V stage 0][V stage 1], 32 KB each.
constexpr uint32_t kPpTmemO = 384;
constexpr int kPpQ = 2 * kQ;   // queries per CTA
struct CfgPp {
  static constexpr int kSlot = 2 * kTileBytes;   // hi + lo tile
  static constexpr int kOffQ1 = kSlot;
  static constexpr int kOffK = 2 * kSlot;
  static constexpr int kOffV = 4 * kSlot;
  static constexpr int kOffBar = 6 * kSlot;
  static constexpr int kOffRed = kOffBar + 256;
  static constexpr int kRedFloats = 2 * 2 * 2 * 128 /* row maxima [group][parity][half][row] */ + 2 * 2 * 128 /* sums */ + 2 * 128 /* maxima */;
  static constexpr int kTotal = 1024 + kOffRed + kRedFloats * 4;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
};

template <bool VMN>
__global__ void __launch_bounds__(kThreads, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap tm_qkv_hi, const __grid_constant__ CUtensorMap tm_qkv_lo,
               const __grid_constant__ CUtensorMap tm_vt_hi, const __grid_constant__ CUtensorMap tm_vt_lo, Args a, int width) {
  using L = CfgPp;
  static_assert(kSmWarps == 16, "two groups of 8 softmax warps");
  constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0);
  constexpr uint32_t kIdescO = ptx::make_idesc_f16(128, 64, 0, VMN ? 1 : 0);
  constexpr uint32_t kIdescS8 = ptx::make_idesc_f8(128, 128, ptx::kF8E5M2, ptx::kF8E5M2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sV = smem + L::kOffV;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;             // [3]
  uint64_t* k_empty = k_full + 3;          // [3]
  uint64_t* v_full = k_empty + 3;          // [2]
  uint64_t* v_empty = v_full + 2;          // [2]
  uint64_t* s_full = v_empty + 2;          // [3] scores of the item in buffer b complete
  uint64_t* p_full = s_full + 3;           // [3] probabilities of the item in buffer b stored (8 warps)
  uint64_t* pv_done = p_full + 3;          // [2] the group's latest P V has retired: O_g is at rest
  uint64_t* o_full = pv_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = ptx::warp_idx_uniform(), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kPpQ, head = blockIdx.y, seq = blockIdx.z;
  const int T = a.T;
  const bool two_q = q0 + kQ < a.S;              // the second query tile has rows: groups by query tile; otherwise by key tile
  const int n_items = two_q ? 2 * T : T;
  const int kring = two_q ? 2 : 3;               // K stages: the Q_1 slot serves as the third one when there is no Q_1
  auto k_slot = [&](int kt) -> uint8_t* {
    const int st = kt % kring;
    return st == 2 ? smem + L::kOffQ1 : smem + L::kOffK + st * L::kSlot;
  };
  const bool dbg = a.dbg && blockIdx.x == 1 && blockIdx.y == 3 && blockIdx.z == 1;
  if (dbg && threadIdx.x == 0) a.dbg[0] = clock64();

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tm_qkv_hi);
    ptx::prefetch_tensormap(&tm_qkv_lo);
    if (!VMN) ptx::prefetch_tensormap(&tm_vt_hi), ptx::prefetch_tensormap(&tm_vt_lo);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 3; ++i) {
      ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1), ptx::mbar_init(&p_full[i], kSmWarps / 2);
    }
    for (int i = 0; i < 2; ++i) ptx::mbar_init(&v_full[i], 1), ptx::mbar_init(&v_empty[i], 1), ptx::mbar_init(&pv_done[i], 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================ TMA producer: K runs `kring` tiles ahead of V ============================
    if (lane == 0) {
      const int row0 = seq * a.S;
      ptx::mbar_arrive_expect_tx(q_full, two_q ? 2 * L::kSlot : L::kSlot);
      for (int t = 0; t < (two_q ? 2 : 1); ++t) {
        tma_load_2d_(sQ + t * L::kSlot, &tm_qkv_hi, q_full, head * kD, row0 + q0 + t * kQ);
        tma_load_2d_(sQ + t * L::kSlot + kTileBytes, &tm_qkv_lo, q_full, head * kD, row0 + q0 + t * kQ);
      }
      auto load_k = [&](int kt) {
        const int st = kt % kring;
        ptx::mbar_wait(&k_empty[st], ((kt / kring) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], L::kSlot);
        uint8_t* dst = k_slot(kt);
        tma_load_2d_(dst, &tm_qkv_hi, &k_full[st], width + head * kD, row0 + kt * kKT);
        tma_load_2d_(dst + kTileBytes, &tm_qkv_lo, &k_full[st], width + head * kD, row0 + kt * kKT);
      };
      for (int kt = 0; kt < kring && kt < T; ++kt) load_k(kt);
      for (int kt = 0; kt < T; ++kt) {
        const int vs = kt & 1;
        ptx::mbar_wait(&v_empty[vs], ((kt >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[vs], L::kSlot);
        uint8_t* dst = sV + vs * L::kSlot;
        if (VMN) {   // rows past S belong to the next sequence (or are zero-filled past the end): their probabilities are exactly 0
          tma_load_2d_(dst, &tm_qkv_hi, &v_full[vs], 2 * width + head * kD, row0 + kt * kKT);
          tma_load_2d_(dst + kTileBytes, &tm_qkv_lo, &v_full[vs], 2 * width + head * kD, row0 + kt * kKT);
        } else {
          const int vrow = (seq * a.heads + head) * kD;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            tma_load_2d_(dst + sub * (kTileBytes / 2), &tm_vt_hi, &v_full[vs], kt * kKT + sub * 64, vrow);
            tma_load_2d_(dst + kTileBytes + sub * (kTileBytes / 2), &tm_vt_lo, &v_full[vs], kt * kKT + sub * 64, vrow);
          }
        }
        if (kt + kring < T) load_k(kt + kring);
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    auto issue_s = [&](int i) {
      const int b = i % 3, kt = two_q ? i >> 1 : i, g = two_q ? i & 1 : 0, st = kt % kring;
      ptx::mbar_wait(&k_full[st], (kt / kring) & 1);
      ptx::tc_fence_after();
      const uint32_t q_addr = ptx::smem_u32(sQ + g * L::kSlot), k_addr = ptx::smem_u32(k_slot(kt));
      const bool leader = ptx::elect_one();
      if (a.qk_f8x) {
        const uint64_t dq_hi = ptx::make_smem_desc_kmajor(q_addr, 128), dq_8 = ptx::make_smem_desc_kmajor(q_addr + kTileBytes, 128);
        const uint64_t dk_hi = ptx::make_smem_desc_kmajor(k_addr, 128), dk_8 = ptx::make_smem_desc_kmajor(k_addr + kTileBytes, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (leader) ptx::umma_f16(tmem_base + b * 128, dq_hi + 2 * k, dk_hi + 2 * k, kIdescS, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // the 128-byte blocks [q 2^-4 | (q - hi) 2^7] x [(k - hi) 2^4 | k 2^-7]: both cross terms
          if (leader) ptx::umma_f8(tmem_base + b * 128, dq_8 + 2 * k, dk_8 + 2 * k, kIdescS8, 1u);
      } else {
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint64_t dq0 = ptx::make_smem_desc_kmajor(q_addr + (pass == 1 ? kTileBytes : 0), 128);
          const uint64_t dk0 = ptx::make_smem_desc_kmajor(k_addr + (pass == 2 ? kTileBytes : 0), 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (leader) ptx::umma_f16(tmem_base + b * 128, dq0 + 2 * k, dk0 + 2 * k, kIdescS, (pass | k) != 0 ? 1u : 0u);
        }
      }
      if (leader) {
        if (!two_q || (i & 1)) ptx::umma_commit(&k_empty[st]);   // the tile's last reader
        ptx::umma_commit(&s_full[b]);
      }
      __syncwarp();
      if (dbg && lane == 0) a.dbg[20 + i] = clock64();
    };
    ptx::mbar_wait(q_full, 0);
    if (dbg && lane == 0) a.dbg[1] = clock64();
    for (int i = 0; i < 3 && i < n_items; ++i) issue_s(i);
    for (int i = 0; i < n_items; ++i) {
      const int b = i % 3, g = i & 1, kt = two_q ? i >> 1 : i;
      ptx::mbar_wait(&p_full[b], (i / 3) & 1);
      if (dbg && lane == 0) a.dbg[10 + i] = clock64();
      ptx::mbar_wait(&v_full[kt & 1], (kt >> 1) & 1);
      ptx::tc_fence_after();
      {
        const uint32_t v_addr = ptx::smem_u32(sV + (kt & 1) * L::kSlot);
        const uint32_t t_o = tmem_base + kPpTmemO + g * 64, t_p = tmem_base + b * 128;
        const bool leader = ptx::elect_one();
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          // pass 0: P_hi V_hi   pass 1: P_lo V_hi   pass 2: P_hi V_lo
          const uint32_t va = v_addr + (pass == 2 ? kTileBytes : 0);
          const uint64_t dv0 = VMN ? ptx::make_smem_desc_mnmajor_sw128(va, 0, 1024) : ptx::make_smem_desc_kmajor(va, 128);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {   // 16 keys per step
            const uint64_t dv = dv0 + (VMN ? kk * (2048 >> 4) : (kk >> 2) * (kTileBytes >> 5) + 2 * (kk & 3));
            const uint32_t tp = t_p + 32 * (kk >> 1) + 8 * (kk & 1) + (pass == 1 ? 16 : 0);
            if (leader) ptx::umma_f16_ts(t_o, tp, dv, kIdescO, (i >= 2 || pass != 0 || kk != 0) ? 1u : 0u);   // items 0 and 1 start O_0 and O_1
          }
        }
        if (leader) {
          if (!two_q || (i & 1)) ptx::umma_commit(&v_empty[kt & 1]);
          ptx::umma_commit(&pv_done[g]);
        }
      }
      __syncwarp();
      if (i + 3 < n_items) issue_s(i + 3);   // into the buffer P V(i) has just read: ordered behind it by the tensor pipe
    }
    if (ptx::elect_one()) ptx::umma_commit(o_full);
    __syncwarp();
  } else {
    // ============================ softmax: group g = items i = g (mod 2); thread <-> row r x 64 key columns (half h) ============================
    const int quarter = warp & 3;                       // TMEM lanes 32*quarter .. +31
    const int idx = (warp - 2) >> 2;                    // 0..3
    const int g = idx & 1, h = idx >> 1;
    const int r = quarter * 32 + lane;                  // row of the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    float* red = reinterpret_cast<float*>(smem + L::kOffRed);
    float* ex_base = red + g * (2 * 2 * 128);           // [parity][half][row]
    float* sums = red + 2 * 2 * 2 * 128;                // [group][half][row]
    float* maxs = sums + 2 * 2 * 128;                   // [group][row]
    const uint32_t t_og = tmem_base + kPpTmemO + g * 64 + lane_addr + h * 32;   // this warp's 32 rows x 32 columns of O_g
    const int bar_id = 2 + g * 4 + quarter;             // the two warps (h = 0, 1) that share these 32 rows in this group
    float m_ref = -INFINITY;   // reference maximum of this row over THIS group's items, in the scaled log2 domain
    float l = 0.f;             // this thread's share of the group's row sum, relative to m_ref
    int n = 0;
    const bool dbg2 = dbg && h == 0 && quarter == 2 && lane == 0;   // warps 2 (group 0) and 6 (group 1)
    for (int i = g; i < n_items; i += 2, ++n) {
      const int b = i % 3, kt = two_q ? i >> 1 : i;
      ptx::mbar_wait(&s_full[b], (i / 3) & 1);
      if (dbg2) a.dbg[30 + i] = clock64();
      ptx::tc_fence_after();
      const uint32_t t_s = tmem_base + lane_addr + b * 128 + h * 64;
      const int key0 = kt * kKT + h * 64;
      const bool full = key0 + 64 <= a.S;
      // ---- pass 1: maximum of this thread's 64 scores ----
      float m_loc = -INFINITY;
#pragma unroll
      for (int sblk = 0; sblk < 2; ++sblk) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(t_s + sblk * 32, v);
        ptx::tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int c = 0; c < 32; ++c) m_loc = fmaxf(m_loc, __uint_as_float(v[c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (key0 + sblk * 32 + c < a.S) m_loc = fmaxf(m_loc, __uint_as_float(v[c]));
        }
      }
      float* ex = ex_base + (n & 1) * (2 * 128);
      ex[h * 128 + r] = m_loc;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      const float m_tile = fmaxf(ex[r], ex[128 + r]);
      if (dbg2) a.dbg[40 + i] = clock64();
      const float mt = m_tile * a.scale_log2e;
      const bool grow = mt > m_ref + a.tau;     // always on the group's first item (m_ref = -inf); the same decision in both warps of the row
      float alpha = 1.f;
      if (grow) {
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(alpha) : "f"(m_ref - mt));   // 0 on the first item
        m_ref = mt;
        l *= alpha;
      }
      if (n > 0 && __any_sync(0xffffffffu, grow)) {   // renew the reference of O_g: this warp's 32 rows x 32 columns
        ptx::mbar_wait(&pv_done[g], (n - 1) & 1);     // the group's previous P V has retired
        ptx::tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t o[16];
          ptx::tmem_ld_32x32b_x16(t_og + cb * 16, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
          ptx::tmem_st_32x32b_x16(t_og + cb * 16, o);
        }
      }
      // ---- pass 2: exponentials, row sums, P = hi + lo stored over the scores just read ----
      {
        const uint64_t c2 = ptx::pack_f32x2(a.scale_log2e, a.scale_log2e), nm2 = ptx::pack_f32x2(-m_ref, -m_ref);
        const uint64_t neg1 = ptx::pack_f32x2(-1.f, -1.f);
        uint64_t l2 = ptx::pack_f32x2(l, 0.f);
#pragma unroll
        for (int sblk = 0; sblk < 2; ++sblk) {
          const int kb0 = key0 + sblk * 32;
          uint32_t ph[16], pl[16];   // packed half2: element 2c, 2c+1 of these 32 columns
          if (kb0 >= a.S) {          // nothing but padding here (the last key tile of a 577-token sequence holds 65 keys): P = 0, no exponentials
#pragma unroll
            for (int c = 0; c < 16; ++c) ph[c] = 0u, pl[c] = 0u;
          } else {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(t_s + sblk * 32, v);
            ptx::tmem_ld_wait();
            auto convert = [&](auto masked) {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                float x0, x1, p0, p1;
                ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1])), c2, nm2), x0, x1);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x0));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x1));
                if (decltype(masked)::value) {
                  if (kb0 + 2 * c >= a.S) p0 = 0.f;
                  if (kb0 + 2 * c + 1 >= a.S) p1 = 0.f;
                }
                const uint64_t pp = ptx::pack_f32x2(p0, p1);
                l2 = ptx::add_f32x2(l2, pp);
                const __half2 h2 = __floats2half2_rn(p0, p1);
                const float2 back = __half22float2(h2);
                float r0, r1;
                ptx::unpack_f32x2(ptx::fma_f32x2(ptx::pack_f32x2(back.x, back.y), neg1, pp), r0, r1);   // p - fp16(p), exact
                const __half2 lo2 = __floats2half2_rn(r0, r1);
                ph[c] = *reinterpret_cast<const uint32_t*>(&h2);
                pl[c] = *reinterpret_cast<const uint32_t*>(&lo2);
              }
            };
            if (full) convert(std::false_type{});
            else convert(std::true_type{});
          }
          ptx::tmem_st_32x32b_x16(t_s + sblk * 32, ph);        // hi halves: the first 16 of the 32 columns just consumed
          ptx::tmem_st_32x32b_x16(t_s + sblk * 32 + 16, pl);   // lo halves: the other 16
        }
        float la, lb;
        ptx::unpack_f32x2(l2, la, lb);
        l = la + lb;
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[b]);
      if (dbg2) a.dbg[50 + i] = clock64();
    }
    // ---- output.  Two query tiles: group g's rows are O_g / l_g.  One query tile: merge the two groups,
    //      O = (O_0 w_0 + O_1 w_1) / (l_0 w_0 + l_1 w_1), w_g = 2^(m_g - max(m_0, m_1)). ----
    sums[(g * 2 + h) * 128 + r] = l;
    if (h == 0) maxs[g * 128 + r] = m_ref;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kSmWarps) : "memory");
    float w0, w1 = 0.f;
    bool use1;                 // this thread's output takes O_1 into account
    if (two_q) {
      const float lg = sums[(g * 2) * 128 + r] + sums[(g * 2 + 1) * 128 + r];
      w0 = g == 0 ? 1.f / lg : 0.f, w1 = g == 1 ? 1.f / lg : 0.f;
      use1 = g == 1;
    } else {
      const float m0 = maxs[r], m1 = maxs[128 + r];
      const float l0 = sums[r] + sums[128 + r], l1 = sums[256 + r] + sums[384 + r];
      use1 = T > 1;                                 // group 1 saw at least one key tile: O_1 is defined
      const float mm = use1 ? fmaxf(m0, m1) : m0;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(m0 - mm));
      if (use1) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(m1 - mm));
      const float inv = 1.f / (l0 * w0 + l1 * w1);
      w0 *= inv, w1 *= inv;
    }
    ptx::mbar_wait(o_full, 0);
    if (dbg2 && g == 0) a.dbg[110] = clock64();
    ptx::tc_fence_after();
    // columns: two query tiles -> this warp (group g, half h) writes 32 columns of its group's rows in two steps of 16;
    //          one query tile  -> the four warps of a lane quarter (idx 0..3) write 16 columns each of the merged rows
    const int q = q0 + (two_q ? g * kQ : 0) + r;
    const int n_steps = two_q ? 2 : 1;
    for (int step = 0; step < n_steps; ++step) {
      const int col0 = two_q ? h * 32 + step * 16 : idx * 16;
      const uint32_t t_o0 = tmem_base + kPpTmemO + lane_addr + col0;
      uint32_t v0[16], v1[16];
      const bool use0 = !two_q || g == 0;
      if (use0) ptx::tmem_ld_32x32b_x16(t_o0, v0);
      if (use1) ptx::tmem_ld_32x32b_x16(t_o0 + 64, v1);
      ptx::tmem_ld_wait();
      if (q < a.S) {
        float x[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float acc = 0.f;
          if (use0) acc = __uint_as_float(v0[c]) * w0;
          if (use1) acc = fmaf(__uint_as_float(v1[c]), w1, acc);
          x[c] = acc;
        }
        uint32_t oh[8], ol[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const __half2 h2 = __floats2half2_rn(x[2 * c], x[2 * c + 1]);
          const float2 back = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(x[2 * c] - back.x, x[2 * c + 1] - back.y);
          oh[c] = *reinterpret_cast<const uint32_t*>(&h2), ol[c] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        const int64_t o = ((int64_t)seq * a.S + q) * a.ldh + head * kD + col0;
#pragma unroll
        for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_hi + o)[c] = make_uint4(oh[4 * c], oh[4 * c + 1], oh[4 * c + 2], oh[4 * c + 3]);
        if (a.lo_format == gemm::LO_F8X) {
          uint32_t f[4], gg[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) gemm::f8x_act4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3], f[c], gg[c]);
          uint8_t* pb = reinterpret_cast<uint8_t*>(a.out_lo + ((int64_t)seq * a.S + q) * a.ldh) + gemm::f8x_off(head * kD + col0);
          *reinterpret_cast<uint4*>(pb) = make_uint4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<uint4*>(pb + 64) = make_uint4(gg[0], gg[1], gg[2], gg[3]);
        } else if (a.out_lo) {
#pragma unroll
          for (int c = 0; c < 2; ++c) reinterpret_cast<uint4*>(a.out_lo + o)[c] = make_uint4(ol[4 * c], ol[4 * c + 1], ol[4 * c + 2], ol[4 * c + 3]);
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) a.dbg[111] = clock64();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

static int make_map_2d(oryon_handle* h, CUtensorMap* tm, const __half* base, int64_t cols, int64_t rows, int64_t ld, int box_cols, int box_rows) {
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = h->encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attn_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ORYON_ERR_CUDA;
  }
  return ORYON_OK;
}

// true when launch() will run the ping-pong kernel at three products (the only kernel that reads gemm::LO_QKV operands)
bool pp_active() {
  static const bool two_pass = getenv("ORYON_ATTN_TWOPASS") != nullptr;
  const char* e = getenv("ORYON_ATTN_LOCKSTEP");   // read per call
  static const bool qk16 = getenv("ORYON_ATTN_QK16") != nullptr;   // A/B switch: three fp16 products for Q K^T in the ping-pong kernel
  return !two_pass && !(e && e[0] == '1') && !qk16;
}

bool v_from_qkv() {
  static const bool vt = getenv("ORYON_ATTN_VT") != nullptr;   // A/B switch: pre-transposed V^T operand
  return !vt;
}

// qkv split pair [n_seq*S][3*width] (q | k | v column blocks, heads of 64) -> out split pair [n_seq*S][ldh] (heads concatenated).
// vt split pair [n_seq*heads*64][ld_vt] (V^T, zero padded for keys >= S up to a multiple of 128) is only read when
// v_from_qkv() is false.
int launch(oryon_handle* h, const __half* qkv_hi, const __half* qkv_lo, const __half* vt_hi, const __half* vt_lo, int ld_vt, int n_seq, int S,
           int heads, int width, int precision, __half* out_hi, __half* out_lo, int64_t ldh, cudaStream_t st, int out_lo_format, int qk_f8x) {
  ORYON_REQUIRE(width == heads * kD, "attn_tc: head dim must be 64");
  ORYON_REQUIRE(out_lo_format == gemm::LO_F16 || (precision == 3 && out_lo && ldh == width), "attn_tc: the 8-bit cross-term output needs ldh == width");
  const int T = (S + kKT - 1) / kKT;
  const bool vmn = v_from_qkv();
  ORYON_REQUIRE(vmn || ld_vt >= T * kKT, "attn_tc: V^T rows must be padded to %d keys", T * kKT);
  CUtensorMap tq_hi, tq_lo, tv_hi, tv_lo;
  int rc;
  const int64_t rows = (int64_t)n_seq * S;
  if ((rc = make_map_2d(h, &tq_hi, qkv_hi, 3 * width, rows, 3 * width, 64, 128))) return rc;
  tv_hi = tq_hi;
  if (!vmn && (rc = make_map_2d(h, &tv_hi, vt_hi, ld_vt, (int64_t)n_seq * heads * kD, ld_vt, 64, 64))) return rc;
  tq_lo = tq_hi, tv_lo = tv_hi;
  if (precision == 3) {
    if ((rc = make_map_2d(h, &tq_lo, qkv_lo, 3 * width, rows, 3 * width, 64, 128))) return rc;
    tv_lo = tq_lo;
    if (!vmn && (rc = make_map_2d(h, &tv_lo, vt_lo, ld_vt, (int64_t)n_seq * heads * kD, ld_vt, 64, 64))) return rc;
  }
  Args a;
  a.S = S, a.heads = heads, a.T = T;
  a.scale_log2e = (1.f / sqrtf((float)kD)) * 1.4426950408889634f;
  static const bool two_pass = getenv("ORYON_ATTN_TWOPASS") != nullptr;     // A/B switch: the two-pass kernel
  const char* tau_env = getenv("ORYON_ATTN_TAU");                           // test switch: 0 renews the reference maximum at every increase
  a.tau = tau_env ? (float)atof(tau_env) : 8.f;
  a.out_hi = out_hi, a.out_lo = precision == 3 ? out_lo : nullptr, a.ldh = ldh, a.lo_format = out_lo_format, a.qk_f8x = qk_f8x;
  a.dbg = nullptr;
  static const bool want_dbg = getenv("ORYON_ATTN_DEBUG") != nullptr;
  static long long* dbg_dev = nullptr;
  static int dbg_calls = 0;
  if (want_dbg) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, 128 * sizeof(long long));
    cudaMemsetAsync(dbg_dev, 0, 128 * sizeof(long long), st);
    a.dbg = dbg_dev;
  }
  const dim3 grid((S + kQ - 1) / kQ, heads, n_seq);
  h->span_begin(KID_ATTN_TC, st);
  auto run = [&](auto kernel, int smem) -> int {
    ORYON_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kernel<<<grid, kThreads, smem, st>>>(tq_hi, tq_lo, tv_hi, tv_lo, a, width);
    return ORYON_OK;
  };
  // Default at three products: the ping-pong kernel (two softmax groups on alternating key tiles); ORYON_ATTN_LOCKSTEP=1 (read per
  // call) selects the one-group online kernel, ORYON_ATTN_TWOPASS the two-pass one.
  const char* lockstep_env = getenv("ORYON_ATTN_LOCKSTEP");
  const bool lockstep = lockstep_env && lockstep_env[0] == '1';
  ORYON_REQUIRE(!qk_f8x || (precision == 3 && !two_pass && !lockstep), "attn_tc: 8-bit Q / K cross-term blocks are read by the ping-pong kernel only");
  if (precision == 3 && !two_pass && !lockstep) {
    auto run_pp = [&](auto kernel) -> int {
      ORYON_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CfgPp::kTotal));
      kernel<<<dim3((S + kPpQ - 1) / kPpQ, heads, n_seq), kThreads, CfgPp::kTotal, st>>>(tq_hi, tq_lo, tv_hi, tv_lo, a, width);
      return ORYON_OK;
    };
    rc = vmn ? run_pp(attn_pp_kernel<true>) : run_pp(attn_pp_kernel<false>);
  } else if (two_pass) {
    if (precision == 3) rc = vmn ? run(attn_tc_kernel<3, true>, Cfg<3>::kTotal) : run(attn_tc_kernel<3, false>, Cfg<3>::kTotal);
    else rc = vmn ? run(attn_tc_kernel<1, true>, Cfg<1>::kTotal) : run(attn_tc_kernel<1, false>, Cfg<1>::kTotal);
  } else {
    if (precision == 3) rc = vmn ? run(attn_online_kernel<3, true>, CfgOnline<3>::kTotal) : run(attn_online_kernel<3, false>, CfgOnline<3>::kTotal);
    else rc = vmn ? run(attn_online_kernel<1, true>, CfgOnline<1>::kTotal) : run(attn_online_kernel<1, false>, CfgOnline<1>::kTotal);
  }
  h->span_end(st);
  if (rc) return rc;
  ORYON_CUDA_CHECK(cudaGetLastError());
  if (want_dbg && ++dbg_calls == 10) {
    long long t[128];
    cudaMemcpyAsync(t, dbg_dev, sizeof(t), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    auto rel = [&](int i) { return t[i] ? (long long)(t[i] - t[0]) : -1; };
    fprintf(stderr, "attn_tc dbg: q_full %lld end %lld o_full %lld\n", rel(1), rel(111), rel(110));
    ORYON_REQUIRE(!qk_f8x || (precision == 3 && !two_pass && !lockstep), "attn_tc: 8-bit Q / K cross-term blocks are read by the ping-pong kernel only");
  if (precision == 3 && !two_pass && !lockstep) {
      for (int j = 0; j < 2 * T; ++j)   // the stamped CTA (1, 3, 1) has two query tiles when S > 384: item j = (key tile j / 2, query tile j % 2)
        fprintf(stderr, "  pp item=%d (group %d): scores issued %lld | softmax s_full %lld max_done %lld p_stored %lld | mma p_full %lld\n", j, j & 1,
                rel(20 + j), rel(30 + j), rel(40 + j), rel(50 + j), rel(10 + j));
      return ORYON_OK;
    }
    for (int i = 0; i < 2 * T; ++i)
      fprintf(stderr, "  i=%d k_full %lld s_empty %lld | softmax s_full %lld\n", i, rel(10 + i), rel(30 + i), rel(70 + i));
    for (int j = 0; j < T; ++j)
      fprintf(stderr, "  j=%d mma p_full %lld v_full %lld | softmax exp_done %lld p_empty %lld\n", j, rel(50 + j), rel(60 + j), rel(90 + j), rel(100 + j));
  }
  return ORYON_OK;
}

}  // namespace attn
}  // namespace oryon
