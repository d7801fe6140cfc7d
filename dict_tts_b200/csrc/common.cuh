// Shared helpers for the libdtts kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DTTS_WARP 32

namespace dtts {

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// 128-bit streaming load that does not pollute L1 (read-once data: dictionary keys/values).
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// start while its predecessor in the stream is still running -- after the predecessor's CTAs have all executed
// griddep_launch() (or exited) -- and must execute griddep_wait() before it touches anything the predecessor wrote (or
// writes anything the predecessor reads): the wait returns when the predecessor grid has completed and its memory
// operations are visible.  Used to hide the prologue (barrier init, TMEM allocation, first weight stages) of the ~250
// short launches of a step behind the tail of the kernel in front.  Kernels launched without the attribute are ordered
// as usual; griddep_launch() in a kernel nobody depends on is a no-op.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Early trigger for short element-wise kernels whose whole grid is resident at once (a multi-wave grid would lose SM
// capacity to the dependent's CTAs waiting in griddep_wait()).
__device__ __forceinline__ void griddep_launch_if_resident() {
  if ((unsigned long long)gridDim.x * gridDim.y * gridDim.z <= 1184ull) griddep_launch();
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: set it once per (kernel, device
// ordinal) -- one process may drive several GPUs (the reference-plugin path).  `done` is the caller's static bit mask;
// a lost race only repeats the idempotent call.
template <typename K>
static inline cudaError_t ensure_max_dyn_smem(K kernel, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) __atomic_fetch_or(done, bit, __ATOMIC_RELEASE);
  return e;
}

}  // namespace dtts
