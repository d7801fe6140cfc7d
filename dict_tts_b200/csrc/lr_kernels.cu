// Length regulator: integer durations -> mel2word -> frame-level gather.  Integer / index work: bit-exact.
// Reference: modules/fastspeech/tts_modules.py:215-251 (LengthRegulator), modules/dict_tts/model.py:98-110
// (pad T to frames_multiple by repeating the last column, zero-row pad + torch.gather).
#include "kernels.cuh"

namespace dtts {

// One warp per utterance: chunked inclusive scan with shuffles.
__global__ void lr_scan_kernel(const int64_t* __restrict__ dur, const int64_t* __restrict__ ilens, int Tw,
                               int* __restrict__ cum, int* __restrict__ totals, int* __restrict__ t_max) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int n = (int)ilens[b];
  n = n < 0 ? 0 : (n > Tw ? Tw : n);
  const int64_t* d = dur + (size_t)b * Tw;
  // first pass: is every duration zero?  (then LengthRegulator fills them with 1, tts_modules.py:248-250)
  int any = 0;
  for (int w = lane; w < n; w += 32) any |= d[w] != 0;
  any = __any_sync(0xffffffffu, any);
  int carry = 0;
  for (int w0 = 0; w0 < Tw; w0 += 32) {
    const int w = w0 + lane;
    int v = 0;
    if (w < n) {
      const int64_t dv = any ? d[w] : 1;
      v = dv > 0 ? (int)dv : 0;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += up;
    }
    v += carry;
    if (w < Tw) cum[(size_t)b * Tw + w] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) {
    totals[b] = carry;
    atomicMax(t_max, carry);
  }
}

cudaError_t lr_scan(const int64_t* dur, const int64_t* ilens, int B, int Tw, int* cum, int* totals, int* t_max,
                    cudaStream_t s) {
  lr_scan_kernel<<<B, 32, 0, s>>>(dur, ilens, Tw, cum, totals, t_max);
  return cudaGetLastError();
}

// mel2word[b,t] = 1 + #(w : cum[b,w] <= t)  for t < total_b, else 0  (binary search over the prefix sums).
__global__ void lr_fill_kernel(const int* __restrict__ cum, const int64_t* __restrict__ ilens, int Tw, int T_raw,
                               int T, int64_t* __restrict__ mel2word) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int tt = t < T_raw ? t : T_raw - 1;                    // padded columns repeat the last real column
  const int* c = cum + (size_t)b * Tw;
  int n = (int)ilens[b];
  n = n < 0 ? 0 : (n > Tw ? Tw : n);
  int64_t m = 0;
  if (tt >= 0 && n > 0 && tt < c[n - 1]) {
    int lo = 0, hi = n - 1;                                    // first w with cum[w] > tt
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (c[mid] > tt) hi = mid; else lo = mid + 1;
    }
    m = lo + 1;
  }
  mel2word[(size_t)b * T + t] = m;
}

cudaError_t lr_fill(const int* cum, const int64_t* ilens, int B, int Tw, int T_raw, int T, int64_t* mel2word,
                    cudaStream_t s) {
  dim3 grid(cdiv(T, 128), B);
  lr_fill_kernel<<<grid, 128, 0, s>>>(cum, ilens, Tw, T_raw, T, mel2word);
  return cudaGetLastError();
}

// Gather 32 frames per block.  Each warp copies whole rows with 128-bit accesses into the [B,T,H] output and a
// padded shared tile; the tile is then written transposed so the [B,H,T] output is coalesced as well.
// Algorithmic traffic: H*4 B read + 2*H*4 B written per frame.
template <int HMAX>
__global__ void __launch_bounds__(256) lr_gather_kernel(const float* __restrict__ enc, const int64_t* __restrict__ m2w,
                                                         int Tw, int T, int H, float* __restrict__ out_btc,
                                                         float* __restrict__ out_bct, float* __restrict__ nonpad) {
  __shared__ float tile[32][HMAX + 1];
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < 32; r += 8) {
    const int t = t0 + r;
    if (t >= T) break;
    long m = m2w[(size_t)b * T + t];
    if (m < 0 || m > Tw) m = 0;
    if (lane == 0) nonpad[(size_t)b * T + t] = m > 0 ? 1.f : 0.f;
    const float* src = enc + ((size_t)b * Tw + (m > 0 ? m - 1 : 0)) * H;
    float* dst = out_btc + ((size_t)b * T + t) * H;
    for (int h = lane; h < H; h += 32) {
      const float v = m > 0 ? src[h] : 0.f;
      dst[h] = v;
      tile[r][h] = v;
    }
  }
  __syncthreads();
  const int t = t0 + lane;
  if (t < T)
    for (int h = warp; h < H; h += 8) out_bct[((size_t)b * H + h) * T + t] = tile[lane][h];
}

cudaError_t lr_gather(const float* enc_btc, const int64_t* mel2word, int B, int Tw, int T, int H, float* out_btc,
                      float* out_bct, float* nonpad, cudaStream_t s) {
  if (H > 256) return cudaErrorInvalidValue;
  dim3 grid(cdiv(T, 32), B);
  lr_gather_kernel<256><<<grid, 256, 0, s>>>(enc_btc, mel2word, Tw, T, H, out_btc, out_bct, nonpad);
  return cudaGetLastError();
}

}  // namespace dtts
