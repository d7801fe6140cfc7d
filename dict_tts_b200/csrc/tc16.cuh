// 16-bit tensor-core operand helpers shared by the convolution kernel and by the element-wise kernels that write
// operand planes directly (LayerNorm, attention, WaveNet gate): conversions and the slab-plane store.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_conv.cuh"

namespace dtts {

// fp32 -> 16-bit operand (fmt 0: fp16, saturated to the finite range; 1: bf16), round to nearest even, and back.
__device__ __forceinline__ uint32_t cvt16(float v, int fmt) {
  if (fmt) return __bfloat16_as_ushort(__float2bfloat16_rn(v));
  return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
}
__device__ __forceinline__ float back16(uint32_t h, int fmt) {
  if (fmt) return __bfloat162float(__ushort_as_bfloat16((unsigned short)h));
  return __half2float(__ushort_as_half((unsigned short)h));
}
// two consecutive channels -> one packed 32-bit word with a single F2FP instruction (fp16 saturates to the finite range)
__device__ __forceinline__ uint32_t pack2(float a0, float a1, int fmt) {
  if (fmt) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  uint32_t r;                                       // one F2FP.SATFINITE.F16.F32.PACK_AB: a0 -> low half, a1 -> high half
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a1), "f"(a0));
  return r;
}
// two consecutive channels -> packed hi word and (residual) lo word
__device__ __forceinline__ void split2(float a0, float a1, int fmt, uint32_t& hw, uint32_t& lw) {
  hw = pack2(a0, a1, fmt);
  lw = pack2(a0 - back16(hw & 0xFFFFu, fmt), a1 - back16(hw >> 16, fmt), fmt);
}

// tanh(t) * sigmoid(s) (the WaveNet gate, wavenet.py:64-70) without the slow paths of tanhf / expf: tanh(t) =
// 1 - 2 / (1 + e^{2t}), both exponentials on ex2.approx (2 ulp), the divisions as approximate reciprocals; absolute error
// ~2e-7 per gate, the saturated ends are exact (e^{2t} = inf -> 1, 0 -> -1).
__device__ __forceinline__ float gate_fast(float t, float s) {
  const float th = 1.f - __fdividef(2.f, 1.f + __expf(2.f * t));
  const float sg = __fdividef(1.f, 1.f + __expf(-s));
  return th * sg;
}

// FP8 lo-plane correction (tc_conv.cuh, TcMode::lo8): the weights of such a layer are packed x 2^10 -- fp16(w * 2^10) for the
// hi plane, e5m2((w - fp16(w)) * 2^10) for the lo plane, which puts w_lo (2^-12 of w) into the e5m2 range -- and the epilogue
// multiplies the accumulator by 2^-10.  The activations need no scale: their e5m2 copy is the high byte of the fp16 value.
constexpr float kLo8WScale = 1024.f;

// Destination operand planes of an element-wise producer: [B][C/8][rows][8], row = pad + t (tc_conv.cuh).
struct PlaneOut {
  tc16* hi = nullptr;
  tc16* lo = nullptr;      // null: single plane
  int rows = 0, pad = 0, fmt = 1;
  int zero_halo = 0;       // the producer also zero-fills rows outside [pad, pad + T) (needed by k > 1 consumers)
};

// 8 consecutive channels (one slab) of time step t
__device__ __forceinline__ void store_slab(const PlaneOut& o, int b, int C, int slab, int t, const float (&v)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (o.lo) split2(v[2 * e], v[2 * e + 1], o.fmt, hw[e], lw[e]);
    else hw[e] = pack2(v[2 * e], v[2 * e + 1], o.fmt);
  }
  const size_t off = (((size_t)b * (C / 8) + slab) * o.rows + o.pad + t) * 8;
  *reinterpret_cast<uint4*>(o.hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if (o.lo) *reinterpret_cast<uint4*>(o.lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// zero the halo rows of batch item b (all threads of the block take part)
__device__ __forceinline__ void zero_halo_rows(const PlaneOut& o, int b, int C, int T) {
  const int nz = o.rows - T, slabs = C / 8;
  for (int i = threadIdx.x; i < nz * slabs; i += blockDim.x) {
    const int sl = i / nz, k = i - sl * nz;
    const int row = k < o.pad ? k : T + k;
    const size_t off = (((size_t)b * slabs + sl) * o.rows + row) * 8;
    *reinterpret_cast<uint4*>(o.hi + off) = make_uint4(0, 0, 0, 0);
    if (o.lo) *reinterpret_cast<uint4*>(o.lo + off) = make_uint4(0, 0, 0, 0);
  }
}

}  // namespace dtts
