// Create-time weight re-layout kernels (run once per handle).
#include "kernels.cuh"

namespace dtts {

__global__ void repack_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int C_out, int C_in, int K,
                                   int rci, int rco) {
  const int n = C_out * C_in * K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int co = i % C_out, r = i / C_out, j = r % K, ci = r / K;     // destination index [ci][j][co]
  const int sci = rci ? C_in - 1 - ci : ci, sco = rco ? C_out - 1 - co : co;
  out[i] = w[((size_t)sco * C_in + sci) * K + j];
}
cudaError_t repack_conv(const float* w, float* out, int C_out, int C_in, int K, int reverse_ci, int reverse_co,
                        cudaStream_t s) {
  const int n = C_out * C_in * K;
  repack_conv_kernel<<<cdiv(n, 256), 256, 0, s>>>(w, out, C_out, C_in, K, reverse_ci, reverse_co);
  return cudaGetLastError();
}

__global__ void repack_convT_kernel(const float* __restrict__ w, float* __restrict__ out, int C_in, int C_out, int K,
                                    int S) {
  const int n = C_in * C_out * K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int M = K / S;
  const int co = i % C_out;
  int r = i / C_out;
  const int m = r % M; r /= M;
  const int ci = r % C_in;
  const int ph = r / C_in;                                             // destination [ph][ci][m][co]
  out[i] = w[((size_t)ci * C_out + co) * K + m * S + ph];
}
cudaError_t repack_convT(const float* w, float* out, int C_in, int C_out, int K, int S, cudaStream_t s) {
  if (K % S) return cudaErrorInvalidValue;
  const int n = C_in * C_out * K;
  repack_convT_kernel<<<cdiv(n, 256), 256, 0, s>>>(w, out, C_in, C_out, K, S);
  return cudaGetLastError();
}

__global__ void transpose2d_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const int r = i % R, c = i / R;                                       // out[c][r]
  out[i] = in[(size_t)r * C + c];
}
cudaError_t transpose2d(const float* in, float* out, int R, int C, cudaStream_t s) {
  transpose2d_kernel<<<cdiv((long)R * C, 256), 256, 0, s>>>(in, out, R, C);
  return cudaGetLastError();
}

__global__ void reverse_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[n - 1 - i];
}
cudaError_t reverse_vec(const float* in, float* out, int n, cudaStream_t s) {
  reverse_kernel<<<cdiv(n, 256), 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace dtts
