// Micro-benchmark (developer tool, not part of libdtts): cycles per tcgen05.mma (kind::f16, M=128, K=16, SS mode) as a
// function of N and of the shared-memory operand layout (no swizzle vs 128-byte swizzle), with all operands already in
// shared memory.  Answers "what is the tensor-pipe / smem-read ceiling of the conv kernel's operand layout?".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__global__ void __launch_bounds__(128, 1) bench(int N, int layout, int iters, int a_step16, int b_step16, int nslots,
                                                long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 96 * 1024;
    uint32_t a_lo, b_lo, hiw;
    if (layout == 0) {           // no swizzle: core matrix 8 rows x 16 B; LBO = 600 rows * 16 B (as in the conv kernel)
      a_lo = ((a_base >> 4) & 0x3FFF) | (600u << 16);
      b_lo = ((b_base >> 4) & 0x3FFF) | ((uint32_t)N << 16);
      hiw = (128u >> 4) | (1u << 14);
    } else {                     // 128-byte swizzle, K-major: rows of 128 B, SBO = 1024 B, layout type 2 @ bits 61-63
      a_lo = ((a_base >> 4) & 0x3FFF) | (1u << 16);
      b_lo = ((b_base >> 4) & 0x3FFF) | (1u << 16);
      hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
    }
    long long t0 = clock64();
    int slot = 0;
    for (int i = 0; i < iters; ++i) {
      const uint64_t a = desc64(a_lo + slot * a_step16, hiw), b = desc64(b_lo + slot * b_step16, hiw);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(a), "l"(b), "r"(idesc), "r"(i)
          : "memory");
      if (++slot == nslots) slot = 0;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar)), "r"(0u)
          : "memory");
    }
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("%-10s %-5s %-22s %s\n", "layout", "N", "operand walk", "cycles per MMA (SM 0 / max over 148 SMs)  ideal N/2");
  for (int layout = 0; layout < 2; ++layout)
    for (int N : {32, 64, 128, 256})
      for (int walk = 0; walk < 3; ++walk) {
        // walk 0: same operands every MMA; 1: A shifts by 1 row (16 B / 128 B) per MMA, 8 slots; 2: A and B move by whole tiles
        int a_step = 0, b_step = 0, nslots = 1;
        if (walk == 1) { a_step = layout == 0 ? 1 : 8; nslots = 8; }
        if (walk == 2) { a_step = layout == 0 ? 128 : 1024; b_step = layout == 0 ? 2 * N : (N * 128 / 16); nslots = 4; }
        bench<<<148, 128, 200 * 1024>>>(N, layout, iters, a_step, b_step, nslots, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-10s %-5d %-22s %.1f / %.1f   %d\n", layout ? "swizzle128" : "none", N,
               walk == 0 ? "fixed" : (walk == 1 ? "A row shift" : "A,B tile step"), (double)h[0] / iters,
               (double)mx / iters, N / 2);
      }
  return 0;
}
