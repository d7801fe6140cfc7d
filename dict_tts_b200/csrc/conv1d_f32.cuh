// Generic fp32 1-D convolution as a register-tiled implicit GEMM on the FP32 FMA pipe.
//
// One kernel covers every convolution on the Dict-TTS path: plain / dilated / strided Conv1d, 1x1 projections and
// ConvTranspose1d (polyphase: one launch plane per output phase).  It is the exact-fp32 path; the HiFi-GAN stack
// additionally has a tcgen05 tensor-core path (hifigan_tc.cu).
//
//   out[b, co, t(q)] = post * ( act( sum_{ci,j} Wp[ci][j][co] * pre(x[b, ci, q*xs + j*xd + x0]) + bias[co] ) * alpha
//                               * mask[b, t] + res[b, co, t] )   (+ out[b, co, t] if accumulate)
//   t(q) = q*ot_mul + ot_add (+ phase for transposed launches)
//
// Reference ops replaced: torch.nn.Conv1d / ConvTranspose1d call sites in modules/hifigan/hifigan.py:27-58,101-142,
// modules/commons/wavenet.py:54-78, modules/commons/rel_transformer_encoder.py:117-126,250-258,
// modules/portaspeech/model.py:58-66, modules/dict_tts/fvae_semantics.py:53-58,93.
#pragma once
#include "common.cuh"

namespace dtts {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2, ACT_GELU = 3 };   // GELU: exact (erf) form, F.gelu default

struct ConvParams {
  // input, element (c, t) of batch b at x[b*x_bs + c*x_cs + t*x_ts]
  const float* x;
  long x_bs;
  int x_cs, x_ts;
  int C_in, T_in;
  // packed weights [phase][C_in][ktaps][w_ld]; this launch uses columns [0, C_out) of the (pre-offset) pointer
  const float* w;
  int w_ld;
  long w_phase_stride;
  const float* bias;  // may be null; already offset to this launch's first output channel
  // output, element (co, t) at out[b*o_bs + co*o_cs + t*o_ts]
  float* out;
  long o_bs;
  int o_cs, o_ts;
  int C_out, T_out;
  const float* res;  // may be null
  long r_bs;
  int r_cs, r_ts;
  const float* mask;  // may be null; mask[b*m_bs + t]
  int m_bs;
  int ktaps, xs, xd, x0;
  int ot_mul, ot_add;
  int nq;
  int phases;  // grid.z = B*phases; transposed conv: ot_add += phase, weights += phase*w_phase_stride
  float pre_slope;  // leaky-relu slope applied to x on load (1 = identity)
  int act;
  float alpha, post;
  int accumulate;
  int ci_chunk;
};

template <int WCO, int WT, int TCO, int TT>
__global__ void __launch_bounds__(256, 2) conv1d_f32_kernel(const ConvParams p) {
  constexpr int CO_TILE = WCO * TCO;
  constexpr int Q_TILE = WT * 32 * TT;
  extern __shared__ float smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wco = warp % WCO, wt = warp / WCO;
  const int b = blockIdx.z / p.phases, phase = blockIdx.z % p.phases;
  const int q0 = blockIdx.x * Q_TILE;
  const int co0 = blockIdx.y * CO_TILE;

  const int span_t = (p.ktaps - 1) * p.xd;                  // signed tap span
  const int min_off = span_t < 0 ? span_t : 0;
  const int XW = (Q_TILE - 1) * p.xs + (span_t < 0 ? -span_t : span_t) + 1;
  const int x_base = q0 * p.xs + p.x0 + min_off;            // global x index of smem column 0
  float* xs_s = smem;                                       // [ci_chunk][XW]
  float* ws_s = smem + (((size_t)p.ci_chunk * XW + 3) & ~(size_t)3);   // [ci_chunk][ktaps][CO_TILE], 16B aligned

  const float* xg = p.x + (size_t)b * p.x_bs;
  const float* wg = p.w + (size_t)phase * p.w_phase_stride + co0;

  float acc[TCO][TT];
#pragma unroll
  for (int c = 0; c < TCO; ++c)
#pragma unroll
    for (int i = 0; i < TT; ++i) acc[c][i] = 0.f;

  const int ql = wt * 32 * TT + lane;                       // local q of element i: ql + 32*i
  const int wrow = p.ktaps * CO_TILE;

  for (int c0 = 0; c0 < p.C_in; c0 += p.ci_chunk) {
    const int nci = min(p.ci_chunk, p.C_in - c0);
    // ---- stage x tile (pre-activation fused) ----
    for (int idx = tid; idx < nci * XW; idx += 256) {
      const int ci = idx / XW, xx = idx - ci * XW;
      const int gx = x_base + xx;
      float v = 0.f;
      if (gx >= 0 && gx < p.T_in) v = __ldg(xg + (size_t)(c0 + ci) * p.x_cs + (size_t)gx * p.x_ts);
      xs_s[idx] = leaky(v, p.pre_slope);
    }
    // ---- stage weight tile ----
    for (int idx = tid; idx < nci * wrow; idx += 256) {
      const int r = idx / CO_TILE, c = idx - r * CO_TILE;   // r = ci*ktaps + j
      float v = 0.f;
      if (co0 + c < p.C_out) v = __ldg(wg + (size_t)(c0 * p.ktaps + r) * p.w_ld + c);
      ws_s[idx] = v;
    }
    __syncthreads();
    for (int ci = 0; ci < nci; ++ci) {
      const float* xrow = xs_s + ci * XW + ql * p.xs - min_off;
      const float* wrow_p = ws_s + ci * wrow + wco * TCO;
      for (int j = 0; j < p.ktaps; ++j) {
        float wv[TCO], xv[TT];
        if constexpr (TCO % 4 == 0) {
#pragma unroll
          for (int c = 0; c < TCO; c += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(wrow_p + j * CO_TILE + c);
            wv[c] = t4.x; wv[c + 1] = t4.y; wv[c + 2] = t4.z; wv[c + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < TCO; ++c) wv[c] = wrow_p[j * CO_TILE + c];
        }
#pragma unroll
        for (int i = 0; i < TT; ++i) xv[i] = xrow[(32 * i) * p.xs + j * p.xd];
#pragma unroll
        for (int c = 0; c < TCO; ++c)
#pragma unroll
          for (int i = 0; i < TT; ++i) acc[c][i] = fmaf(wv[c], xv[i], acc[c][i]);
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int ot_add = p.ot_add + phase;
#pragma unroll
  for (int i = 0; i < TT; ++i) {
    const int q = q0 + ql + 32 * i;
    if (q >= p.nq) continue;
    const int t = q * p.ot_mul + ot_add;
    if (t < 0 || t >= p.T_out) continue;
    const float mk = p.mask ? p.mask[(size_t)b * p.m_bs + t] : 1.f;
#pragma unroll
    for (int c = 0; c < TCO; ++c) {
      const int co = co0 + wco * TCO + c;
      if (co >= p.C_out) continue;
      float v = acc[c][i];
      if (p.bias) v += __ldg(p.bias + co);
      if (p.act == ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == ACT_TANH) v = tanhf(v);
      else if (p.act == ACT_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
      v *= p.alpha * mk;
      if (p.res) v += p.res[(size_t)b * p.r_bs + (size_t)co * p.r_cs + (size_t)t * p.r_ts];
      v *= p.post;
      float* o = p.out + (size_t)b * p.o_bs + (size_t)co * p.o_cs + (size_t)t * p.o_ts;
      if (p.accumulate) v += *o;
      *o = v;
    }
  }
}

// Host-side launch: picks the tile shape from (C_out, nq) and sizes dynamic shared memory.
cudaError_t launch_conv1d_f32(ConvParams p, int B, cudaStream_t stream);
// One-time: raise the dynamic shared memory limit of every instantiation.
cudaError_t conv1d_f32_init();

}  // namespace dtts
