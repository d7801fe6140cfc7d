// tcgen05.mma issue-rate microbenchmark (sm_100a): what does ONE MMA of the shapes the convolution kernels use cost when
// nothing else is in its way?  Operands sit in shared memory (no-swizzle K-major, as tc_conv.cu lays them out), one thread
// per CTA (per CTA pair for cta_group::2) issues REPS x PATTERN instructions back to back and waits for the last commit.
// All 148 SMs run the same loop at the same time, so the clocks are the loaded ones.  Prints cycles per instruction and
// the FLOP/clk/SM that follows; these are the denominators of the per-stage tensor floors in DESIGN.md 4.1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I dict_tts_b200/csrc -o tools/_build/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace dtts;

struct Case {
  int pair;       // 0: cta_group::1 (M = 128), 1: cta_group::2 (M = 256)
  int N;          // accumulator columns of one instruction
  int n16, n8;    // per (tap, m): kind::f16 K = 16 instructions, then kind::f8f6f4 K = 32 instructions
  int nacc;       // 128-row sub-tiles per CTA that share one B operand
  int run;        // taps issued back to back into ONE accumulator before moving to the next sub-tile (1: tap-major order)
  const char* what;
};

template <bool PAIR, int N16, int N8>
__global__ void __launch_bounds__(128, 1) rate_kernel(Case c, int reps, unsigned long long* cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(smem), tptr = smem_u32(smem + 16);
  const uint32_t a_base = smem_u32(smem + 1024), b_base = a_base + 72 * 1024;   // A: <= 2 slabs x 2 x (256 + 16) rows x 16 B x 4
  for (int i = threadIdx.x; i < (200 * 1024 - 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 1024)[i] = 0;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tptr), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 16);
  if (warp == 1 && rank == 0 && elect_one()) {
    const uint32_t hiw = (128u >> 4) | (1u << 14);
    const uint32_t RA = 128u * (uint32_t)c.nacc + 16u;
    const uint32_t NB = (uint32_t)(PAIR ? c.N / 2 : c.N);               // rows of B in this CTA
    const uint32_t a_low = ((a_base >> 4) & 0x3FFFu) | (RA << 16);
    const uint32_t b_low = ((b_base >> 4) & 0x3FFFu) | (NB << 16);
    const uint32_t mfield = (PAIR ? 256u >> 4 : 128u >> 4) << 24;
    const uint32_t idesc16 = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | mfield;                          // f16 x f16 -> f32
    const uint32_t idesc8 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | mfield;  // e5m2
    auto mma16 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
      if (PAIR) umma_bf16_2cta(d, a, b, id, acc); else umma_bf16(d, a, b, id, acc);
    };
    auto mma8 = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
      if (PAIR) umma_f8_2cta(d, a, b, id, acc); else umma_f8(d, a, b, id, acc);
    };
    const long long t0 = clock64();
    // lean issue loop (the instruction counts are compile-time, no division): the single thread must not be the limiter
    const int run = c.run, nacc = c.nacc;
    const uint32_t Ncols = (uint32_t)c.N, a8 = 4u * RA, b8 = b_low + 4u * NB;
    uint32_t tap0 = 0;
    for (int r = 0; r < reps; r += run) {
      uint32_t d = tmem + ((r & 64) && nacc * c.N <= 256 ? 256u : 0u);
      uint32_t am0 = a_low;
      for (int m = 0; m < nacc; ++m, d += Ncols, am0 += 128u) {
        uint32_t tap = tap0;
        for (int j = 0; j < run; ++j) {
          const uint32_t am = am0 + tap;                                // the A window slides by one row per tap
#pragma unroll
          for (int k = 0; k < N16; ++k)
            mma16(d, desc64(am + (uint32_t)k * 2u * RA, hiw), desc64(b_low + (uint32_t)k * 2u * NB, hiw), idesc16, 1u);
#pragma unroll
          for (int k = 0; k < N8; ++k) mma8(d, desc64(am + a8, hiw), desc64(b8, hiw), idesc8, 1u);
          if (++tap == 11u) tap = 0u;
        }
      }
      tap0 += (uint32_t)run;
      while (tap0 >= 11u) tap0 -= 11u;
    }
    if (PAIR) umma_commit_2cta(bar); else umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  } else if (PAIR && rank == 1 && threadIdx.x == 0) {
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <bool PAIR, int N16, int N8>
static cudaError_t launch_one(cudaLaunchConfig_t& cfg, const Case& c, int reps, unsigned long long* cyc) {
  cudaError_t e = cudaFuncSetAttribute(rate_kernel<PAIR, N16, N8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
  if (e != cudaSuccess) return e;
  return cudaLaunchKernelEx(&cfg, rate_kernel<PAIR, N16, N8>, c, reps, cyc);
}
static cudaError_t launch(cudaLaunchConfig_t& cfg, const Case& c, int reps, unsigned long long* cyc) {
  const int key = (c.pair ? 100 : 0) + c.n16 * 10 + c.n8;
  switch (key) {
    case 10: return launch_one<false, 1, 0>(cfg, c, reps, cyc);
    case 20: return launch_one<false, 2, 0>(cfg, c, reps, cyc);
    case 1: return launch_one<false, 0, 1>(cfg, c, reps, cyc);
    case 21: return launch_one<false, 2, 1>(cfg, c, reps, cyc);
    case 110: return launch_one<true, 1, 0>(cfg, c, reps, cyc);
    case 120: return launch_one<true, 2, 0>(cfg, c, reps, cyc);
    case 140: return launch_one<true, 4, 0>(cfg, c, reps, cyc);
    case 101: return launch_one<true, 0, 1>(cfg, c, reps, cyc);
    case 121: return launch_one<true, 2, 1>(cfg, c, reps, cyc);
    default: return cudaErrorInvalidValue;
  }
}

int main() {
  const Case cases[] = {
      // instructions per accumulator visit = run * (n16 + n8)
      {0, 128, 1, 0, 1, 1, "cg1 f16 N=128, ONE accumulator"},
      {0, 128, 1, 0, 2, 1, "cg1 f16 N=128, 2 accumulators, 1 instr per visit"},
      {0, 128, 2, 0, 2, 1, "cg1 f16 N=128, 2 accumulators, 2 per visit (rb_pair64 today)"},
      {0, 128, 2, 0, 2, 2, "cg1 f16 N=128, 2 accumulators, 4 per visit"},
      {0, 128, 2, 0, 2, 4, "cg1 f16 N=128, 2 accumulators, 8 per visit (rb_pair64, sub-tile-major inside a weight stage)"},
      {0, 128, 2, 0, 2, 16, "cg1 f16 N=128, 2 accumulators, 32 per visit"},
      {0, 64, 2, 0, 2, 1, "cg1 f16 N=64, 2 accumulators, 2 per visit (rb_pair32 today)"},
      {0, 64, 2, 0, 2, 4, "cg1 f16 N=64, 2 accumulators, 8 per visit"},
      {0, 64, 2, 0, 2, 11, "cg1 f16 N=64, 2 accumulators, 22 per visit"},
      {0, 256, 2, 0, 1, 1, "cg1 f16 N=256, ONE accumulator"},
      {0, 128, 0, 1, 1, 1, "cg1 f8 N=128 K=32, ONE accumulator"},
      {0, 128, 0, 1, 2, 4, "cg1 f8 N=128 K=32, 2 accumulators, 4 per visit"},
      {0, 256, 0, 1, 1, 1, "cg1 f8 N=256 K=32, ONE accumulator"},
      {1, 128, 1, 0, 1, 1, "cg2 f16 N=128 M=256, ONE accumulator"},
      {1, 128, 2, 0, 2, 1, "cg2 f16 N=128 M=256, 2 accumulators, 2 per visit"},
      {1, 128, 2, 0, 2, 4, "cg2 f16 N=128 M=256, 2 accumulators, 8 per visit"},
      {1, 128, 4, 0, 2, 1, "cg2 N=128 stage-2 k=3 today: 4 x f16 per visit"},
      {1, 128, 4, 0, 2, 3, "cg2 N=128 stage-2 k=3, sub-tile-major: 12 x f16 per visit"},
      {1, 128, 2, 1, 2, 1, "cg2 N=128 stage-2 lo8 today: 2 x f16 + 1 x f8 per visit"},
      {1, 128, 2, 1, 2, 4, "cg2 N=128 stage-2 lo8, sub-tile-major: 4 x (2 x f16 + f8) per visit"},
      {1, 128, 0, 1, 1, 1, "cg2 f8 N=128 M=256 K=32, ONE accumulator"},
      {1, 256, 2, 0, 1, 1, "cg2 f16 N=256 M=256, ONE accumulator"},
      {1, 256, 2, 1, 1, 1, "cg2 N=256 stage-1 lo8: 2 x f16 + 1 x f8, ONE accumulator"},
      {1, 256, 0, 1, 1, 1, "cg2 f8 N=256 M=256 K=32, ONE accumulator"},
  };

  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  unsigned long long* cyc;
  CK(cudaMallocManaged(&cyc, 256 * sizeof(unsigned long long)));
  const int smem = 200 * 1024, reps = 4224;
  printf("%d SMs, %d repetitions of the pattern per CTA\n", dev_sms, reps);
  for (const Case& c : cases) {
    for (int pass = 0; pass < 2; ++pass) {          // pass 0 warms up
      for (int i = 0; i < 256; ++i) cyc[i] = 0;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(dev_sms & ~1);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c.pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      CK(launch(cfg, c, reps, cyc));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (pass == 0) continue;
      unsigned long long mx = 0, mn = ~0ull;
      for (int i = 0; i < (dev_sms & ~1); ++i) if (cyc[i]) { mx = cyc[i] > mx ? cyc[i] : mx; mn = cyc[i] < mn ? cyc[i] : mn; }
      const double n_ins = (double)(reps / c.run * c.run) * c.nacc * (c.n16 + c.n8);
      // dense-equivalent flop per SM: every instruction is 128 rows (per CTA) x N x K x 2
      const double flop = (double)(reps / c.run * c.run) * c.nacc * 128.0 * c.N * 2.0 * (16.0 * c.n16 + 32.0 * c.n8);
      printf("%-100s  %7.1f cyc/instr (min %7.1f)  %7.0f flop/clk/SM  %.3f ms -> %.2f GHz\n", c.what, mx / n_ins, mn / n_ins,
             flop / mx, ms, mx / (ms * 1e6));
    }
  }
  return 0;
}
