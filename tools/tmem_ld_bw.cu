// Micro-benchmark: how fast can warps read accumulators out of tensor memory (tcgen05.ld), alone and while the tensor pipe is
// filling other columns -- the quantity that bounds match_tc_kernel at D = 128 (4 bytes of tensor memory per
// score against 1 024 tensor cycles per 128 KB tile).  One CTA per SM, all 512 columns allocated.
//   mode 0: W reader warps, one tcgen05.ld.32x32b.x32 (4 KB) in flight per warp
//   mode 1: W reader warps, two loads in flight per warp
//   mode 2: as 1, plus the matcher's scan of each 32-column group (four FMNMX chains + one test)
//   mode 3: as 2, while one more warp issues 128x128x16 fp16 UMMAs back to back into the other half of the columns
//   mode 4: the UMMAs alone (their rate without readers), one tcgen05.commit per 64 UMMAs
//   mode 5 / 6 / 7: the UMMAs alone with one tcgen05.commit (to a barrier nobody waits on) per 8 / 4 / 2 UMMAs: what a commit costs
//   mode 8: one commit per 8 UMMAs, issued AFTER the first UMMA of the next accumulator instead of at the boundary
//   mode 9: one commit per 8 UMMAs in the MIDDLE of the accumulation (after the 4th UMMA)
//   mode 10 / 11: UMMAs alone, M = 128, N = 64 (the P V shape of d = 64 attention): both operands in shared memory / A in tensor memory
//   mode 12: M = 128, N = 128 with A in tensor memory          (tensor_duty of modes 10-12 is relative to N / 256 * 128 cycles per UMMA)
// Prints bytes per clock and SM for the readers and the tensor-pipe duty of the UMMAs.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I oryon_b200/csrc tools/tmem_ld_bw.cu -o tools/bin/tmem_ld_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "ptx_sm100.cuh"

using namespace oryon;

constexpr int kIters = 2048;        // 32-column groups per reader warp
constexpr int kMmaBatches = 1024;   // batches of 8 UMMAs (one 128 x 128 x 128 accumulation)

struct Out {
  long long reader_cycles;   // slowest reader warp of the CTA
  long long mma_cycles;
  float sink;
};

template <int mode>
__global__ void __launch_bounds__(32 * 17) tmem_bw_kernel(int n_readers, int iters, Out* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ long long cyc[17];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mma_warp = n_readers;   // the warp after the readers
  for (int i = threadIdx.x; i < 2 * 128 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < 17) cyc[threadIdx.x] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  float sink = 0.f;

  if (warp < n_readers && mode < 4) {
    const int quarter = warp & 3;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint32_t va[32], vb[32];
    float thr = 3.0e38f, m_run = -3.0e38f;
    int cnt = 0;
    auto scan = [&](uint32_t (&v)[32]) {
      float cm[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float m = fmaxf(fmaxf(__uint_as_float(v[c * 8]), __uint_as_float(v[c * 8 + 1])), __uint_as_float(v[c * 8 + 2]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[c * 8 + 3])), __uint_as_float(v[c * 8 + 4]));
        m = fmaxf(fmaxf(m, __uint_as_float(v[c * 8 + 5])), __uint_as_float(v[c * 8 + 6]));
        cm[c] = fmaxf(m, __uint_as_float(v[c * 8 + 7]));
      }
      const float gm = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
      if (gm >= thr) {
        ++cnt;
        m_run = fmaxf(m_run, gm);
      }
    };
    __syncwarp();
    const long long t0 = clock64();
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
        ptx::tmem_ld_32x32b_x32(taddr + ((it * 32) & 255), va);   // the half of the columns the UMMAs of mode 3 do not write
        ptx::tmem_ld_wait();
        sink += __uint_as_float(va[0]) + __uint_as_float(va[31]);   // static indices: the arrays must stay in registers
      }
    } else {
      ptx::tmem_ld_32x32b_x32(taddr, va);
      for (int it = 0; it < iters; it += 2) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + (((it + 1) * 32) & 255), vb);
        if (mode >= 2) scan(va); else sink += __uint_as_float(va[0]) + __uint_as_float(va[31]);
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32b_x32(taddr + (((it + 2) * 32) & 255), va);
        if (mode >= 2) scan(vb); else sink += __uint_as_float(vb[0]) + __uint_as_float(vb[31]);
      }
      ptx::tmem_ld_wait();
    }
    const long long t1 = clock64();
    sink += m_run + cnt;
    if (lane == 0) cyc[warp] = t1 - t0;
  } else if (warp == mma_warp && mode >= 3) {
    constexpr int kN = (mode == 10 || mode == 11) ? 64 : 128;
    constexpr bool kTs = mode == 11 || mode == 12;
    constexpr uint32_t idesc = ptx::make_idesc_f16(128, kN, 0);
    const uint64_t da = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem), 128);
    const uint64_t db = ptx::make_smem_desc_kmajor(ptx::smem_u32(smem + 128 * 128), 128);
    __syncwarp();
    const long long t0 = clock64();
    uint32_t phase = 0;
    for (int b = 0; b < kMmaBatches; ++b) {
      const bool leader = ptx::elect_one();
      const uint32_t d = tmem_base + 256 + (b & 1) * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kTs) {
          if (leader) ptx::umma_f16_ts(d, tmem_base + 8 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);   // A: columns 0 .. 31 of tensor memory
        } else {
          if (leader) ptx::umma_f16(d, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
        }
        if (mode == 7 && (k & 1) && leader) ptx::umma_commit(&bar[1]);
        if (mode == 8 && k == 0 && b > 0 && leader) ptx::umma_commit(&bar[1]);
      }
      if ((mode == 6 || mode == 9) && leader) ptx::umma_commit(&bar[1]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kTs) {
          if (leader) ptx::umma_f16_ts(d, tmem_base + 8 * k, db + 2 * k, idesc, 1u);
        } else {
          if (leader) ptx::umma_f16(d, da + 2 * k, db + 2 * k, idesc, 1u);
        }
        if (mode == 7 && (k & 1) && leader) ptx::umma_commit(&bar[1]);
      }
      if ((mode == 5 || mode == 6) && leader) ptx::umma_commit(&bar[1]);
      if ((b & 7) == 7) {   // bound the queue: wait for every eighth batch
        if (leader) ptx::umma_commit(&bar[0]);
        __syncwarp();
        ptx::mbar_wait(&bar[0], phase);
        phase ^= 1;
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (lane == 0) cyc[16] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long r = 0;
    for (int w = 0; w < n_readers; ++w) r = cyc[w] > r ? cyc[w] : r;
    out[blockIdx.x].reader_cycles = r;
    out[blockIdx.x].mma_cycles = cyc[16];
  }
  if (sink == 123.456f) out[blockIdx.x].sink = sink;
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  Out* d_out = nullptr;
  cudaMalloc(&d_out, sizeof(Out) * sms);
  const int smem_bytes = 1024 + 2 * 128 * 128;
  using Kern = void (*)(int, int, Out*);
  const Kern kerns[13] = {tmem_bw_kernel<0>, tmem_bw_kernel<1>, tmem_bw_kernel<2>, tmem_bw_kernel<3>, tmem_bw_kernel<4>,
                          tmem_bw_kernel<5>, tmem_bw_kernel<6>, tmem_bw_kernel<7>, tmem_bw_kernel<8>, tmem_bw_kernel<9>,
                          tmem_bw_kernel<10>, tmem_bw_kernel<11>, tmem_bw_kernel<12>};
  for (Kern k : kerns) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<Out> h(sms);
  printf("{\"sms\": %d, \"rows\": [\n", sms);
  bool first = true;
  for (int mode = 0; mode <= 12; ++mode) {
    for (int readers : {4, 8, 16}) {
      if (mode >= 4 && readers != 4) continue;
      // mode 3: the readers outlast the UMMAs, so the tensor duty is measured under read load throughout
      const int iters = mode == 3 ? 6 * kIters : kIters;
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d_out, 0, sizeof(Out) * sms);
        kerns[mode]<<<sms, 32 * 17, smem_bytes>>>(readers, iters, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          fprintf(stderr, "mode %d readers %d: %s\n", mode, readers, cudaGetErrorString(e));
          return 1;
        }
      }
      cudaMemcpy(h.data(), d_out, sizeof(Out) * sms, cudaMemcpyDeviceToHost);
      std::vector<long long> rc, mc;
      for (auto& o : h) rc.push_back(o.reader_cycles), mc.push_back(o.mma_cycles);
      std::sort(rc.begin(), rc.end());
      std::sort(mc.begin(), mc.end());
      const double r_med = (double)rc[sms / 2], m_med = (double)mc[sms / 2];
      const double bytes = (double)readers * iters * 4096.0;
      const double mma_ideal = (double)kMmaBatches * 8 * ((mode == 10 || mode == 11) ? 32 : 64);   // 128 x N x 16 fp16 = N / 2 tensor cycles
      printf("%s  {\"mode\": %d, \"reader_warps\": %d, \"reader_cycles\": %.0f, \"ld_bytes_per_clk_sm\": %.1f, \"cycles_per_4KB_load_per_warp\": %.1f, "
             "\"mma_cycles\": %.0f, \"tensor_duty\": %.3f}",
             first ? "" : ",\n", mode, readers, r_med, r_med > 0 ? bytes / r_med : 0.0, r_med > 0 ? r_med / iters : 0.0, m_med,
             m_med > 0 ? mma_ideal / m_med : 0.0);
      first = false;
    }
  }
  printf("\n]}\n");
  cudaFree(d_out);
  return 0;
}
