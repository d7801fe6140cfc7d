#include "conv1d_f32.cuh"

namespace dtts {

namespace {

constexpr int kMaxSmem = 100 * 1024;  // two CTAs per SM

template <int WCO, int WT, int TCO, int TT>
cudaError_t launch_t(ConvParams p, int B, cudaStream_t stream) {
  constexpr int CO_TILE = WCO * TCO;
  constexpr int Q_TILE = WT * 32 * TT;
  const int span = (p.ktaps - 1) * (p.xd < 0 ? -p.xd : p.xd);
  const int XW = (Q_TILE - 1) * p.xs + span + 1;
  // K-chunk: aim for ~96 (ci, tap) rows per stage, bounded by shared memory
  int chunk = 96 / p.ktaps;
  if (chunk < 4) chunk = 4;
  if (chunk > p.C_in) chunk = p.C_in;
  auto bytes = [&](int c) { return (((size_t)c * XW + 3) / 4 * 4 + (size_t)c * p.ktaps * CO_TILE) * sizeof(float); };
  while (chunk > 1 && bytes(chunk) > (size_t)kMaxSmem) chunk--;
  if (bytes(chunk) > 200 * 1024) return cudaErrorInvalidConfiguration;
  p.ci_chunk = chunk;
  dim3 grid(cdiv(p.nq, Q_TILE), cdiv(p.C_out, CO_TILE), B * p.phases);
  conv1d_f32_kernel<WCO, WT, TCO, TT><<<grid, 256, bytes(chunk), stream>>>(p);
  return cudaGetLastError();
}

template <int WCO, int WT, int TCO, int TT>
cudaError_t init_t() {
  return cudaFuncSetAttribute(conv1d_f32_kernel<WCO, WT, TCO, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              200 * 1024);
}

}  // namespace

cudaError_t conv1d_f32_init() {
  cudaError_t e;
  if ((e = init_t<8, 1, 8, 8>()) != cudaSuccess) return e;
  if ((e = init_t<8, 1, 8, 4>()) != cudaSuccess) return e;
  if ((e = init_t<8, 1, 8, 2>()) != cudaSuccess) return e;
  if ((e = init_t<8, 1, 8, 1>()) != cudaSuccess) return e;
  if ((e = init_t<4, 2, 8, 8>()) != cudaSuccess) return e;
  if ((e = init_t<4, 2, 8, 2>()) != cudaSuccess) return e;
  if ((e = init_t<1, 8, 8, 4>()) != cudaSuccess) return e;
  if ((e = init_t<1, 8, 8, 1>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_conv1d_f32(ConvParams p, int B, cudaStream_t stream) {
  if (p.phases < 1) p.phases = 1;
  if (B <= 0 || p.nq <= 0 || p.C_out <= 0) return cudaSuccess;
  const int nq = p.nq;
  if (p.C_out <= 16) {             // flow post (8), latent (16), conv_post (1)
    return nq > 2048 ? launch_t<1, 8, 8, 4>(p, B, stream) : launch_t<1, 8, 8, 1>(p, B, stream);
  }
  if (p.C_out <= 32) {             // last HiFi-GAN stage
    return nq > 256 ? launch_t<4, 2, 8, 8>(p, B, stream) : launch_t<4, 2, 8, 2>(p, B, stream);
  }
  if (nq > 1024) return launch_t<8, 1, 8, 8>(p, B, stream);   // Q tile 256
  if (nq > 64) {
    // pick the Q tile (128 or 64) wasting the fewest padded columns
    const int w128 = cdiv(nq, 128) * 128, w64 = cdiv(nq, 64) * 64;
    return (w128 <= w64) ? launch_t<8, 1, 8, 4>(p, B, stream) : launch_t<8, 1, 8, 2>(p, B, stream);
  }
  return nq > 32 ? launch_t<8, 1, 8, 2>(p, B, stream) : launch_t<8, 1, 8, 1>(p, B, stream);
}

}  // namespace dtts
