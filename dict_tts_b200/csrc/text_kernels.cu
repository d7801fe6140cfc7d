// Text-encoder side kernels: embedding, channel LayerNorm, tiny-sequence self attention, the S2PA dictionary
// attention (streaming, folded form), pronunciation mixing, duration head.
//
// Reference semantics: modules/dict_tts/layers/dict_encoder.py:32-66,130-144, modules/dict_tts/layers/utils.py:40-58,
// 109-115, modules/commons/rel_transformer_encoder.py:55-79,117-158,261-279, modules/portaspeech/model.py:58-66,
// modules/dict_tts/model.py:64-82.
#include "kernels.cuh"
#include "tc16.cuh"

namespace dtts {

// ------------------------------------------------------------------------------------------------------------
__global__ void embed_kernel(const int64_t* __restrict__ tok, const float* __restrict__ emb, float scale, int Tw,
                             int H, int vocab, float* __restrict__ x, float* __restrict__ seq_mask,
                             float* __restrict__ tok_mask, int* __restrict__ lens) {
  const int b = blockIdx.x;
  __shared__ int s_len;
  if (threadIdx.x == 0) s_len = 0;
  __syncthreads();
  int cnt = 0;
  for (int t = threadIdx.x; t < Tw; t += blockDim.x) cnt += tok[(size_t)b * Tw + t] > 0;
  cnt = (int)warp_sum((float)cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_len, cnt);
  __syncthreads();
  const int len = s_len;
  if (threadIdx.x == 0) lens[b] = len;
  for (int t = threadIdx.x; t < Tw; t += blockDim.x) {
    seq_mask[(size_t)b * Tw + t] = t < len ? 1.f : 0.f;      // sequence_mask(x_lengths) -- prefix mask by COUNT
    tok_mask[(size_t)b * Tw + t] = tok[(size_t)b * Tw + t] > 0 ? 1.f : 0.f;
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < H * Tw; i += blockDim.x) {
    const int c = i / Tw, t = i - c * Tw;
    long id = tok[(size_t)b * Tw + t];
    if (id < 0 || id >= vocab) id = 0;
    x[(size_t)b * H * Tw + i] = emb[(size_t)id * H + c] * scale;
  }
}

cudaError_t embed_tokens(const int64_t* tok, const float* emb, float scale, int B, int Tw, int H, int vocab, float* x,
                         float* seq_mask, float* tok_mask, int* lens, cudaStream_t s) {
  embed_kernel<<<B, 256, 0, s>>>(tok, emb, scale, Tw, H, vocab, x, seq_mask, tok_mask, lens);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// One thread per (b, t) column; lanes run along t so every pass over C is coalesced.
__global__ void channel_ln_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, const float* __restrict__ in_mask,
                                  const float* __restrict__ out_mask, int C, int T, int relu) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* xp = x + (size_t)b * C * T + t;
  float* yp = y + (size_t)b * C * T + t;
  const float im = in_mask ? in_mask[(size_t)b * T + t] : 1.f;
  const float om = out_mask ? out_mask[(size_t)b * T + t] : 1.f;
  float mean = 0.f;
  for (int c = 0; c < C; ++c) mean += xp[(size_t)c * T] * im;
  mean /= (float)C;
  float var = 0.f;
  for (int c = 0; c < C; ++c) {
    const float d = xp[(size_t)c * T] * im - mean;
    var = fmaf(d, d, var);
  }
  const float rstd = rsqrtf(var / (float)C + eps);
  for (int c = 0; c < C; ++c) {
    float v = (xp[(size_t)c * T] * im - mean) * rstd * gamma[c] + beta[c];
    if (relu) v = fmaxf(v, 0.f);
    yp[(size_t)c * T] = v * om;
  }
}

__global__ void channel_ln_tiled_kernel(const float* __restrict__ x, float* __restrict__ xw, float* __restrict__ y,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        const float* __restrict__ in_mask, const float* __restrict__ out_mask, int C, int T,
                                        PlaneOut po, int relu);

cudaError_t channel_layernorm(const float* x, float* y, const float* gamma, const float* beta, float eps,
                              const float* in_mask, const float* out_mask, int B, int C, int T, cudaStream_t s,
                              int relu) {
  dim3 grid(cdiv(T, 32), B);
  const size_t smem = ((size_t)C * 33 + 256 + 64) * sizeof(float);
  if (smem <= 48 * 1024) {          // the tiled kernel (8 x 32 threads per 32 time steps): 30 -> 10 us at [60, 192, 64]
    channel_ln_tiled_kernel<<<grid, 256, smem, s>>>(x, nullptr, y, gamma, beta, eps, in_mask, out_mask, C, T, PlaneOut{}, relu);
    return cudaGetLastError();
  }
  channel_ln_kernel<<<grid, 32, 0, s>>>(x, y, gamma, beta, eps, in_mask, out_mask, C, T, relu);
  return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------------------
// Tiled channel LayerNorm: block = (32 time steps, batch item), 256 threads = 8 channel groups x 32 lanes along t.
// The tile is staged in shared memory (coalesced along t), statistics are two-pass like the reference, and the result
// goes to y [B,C,T] and/or straight into tensor-core operand planes (saves a conversion launch per convolution).
// xw (optional): x * in_mask is written back (Encoder.forward's `x = x * x_mask`, rel_transformer_encoder.py:58).
__global__ void __launch_bounds__(256) channel_ln_tiled_kernel(const float* __restrict__ x, float* __restrict__ xw,
                                                               float* __restrict__ y, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps,
                                                               const float* __restrict__ in_mask,
                                                               const float* __restrict__ out_mask, int C, int T,
                                                               PlaneOut po, int relu) {
  extern __shared__ float sm[];                      // xs[C][33], red[8][32], mean[32], rstd[32]
  griddep_launch_if_resident();                      // the consumer (a tcgen05 launch) may set itself up meanwhile
  float* xs = sm;
  float* red = sm + (size_t)C * 33;
  float* mean_s = red + 256;
  float* rstd_s = mean_s + 32;
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int t = t0 + lane;
  const bool tv = t < T;
  const float im = (in_mask && tv) ? in_mask[(size_t)b * T + t] : 1.f;
  const float om = (out_mask && tv) ? out_mask[(size_t)b * T + t] : 1.f;
  const float* xb = x + (size_t)b * C * T;
  float part = 0.f;
  // eight loads in flight per thread (one load per iteration left this kernel waiting ~0.5 us of global latency per
  // channel: 16 us for 22 tokens x 192 channels); same summation order as the plain loop
  for (int c0 = grp; c0 < C; c0 += 64) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + 8 * i;
      v[i] = (tv && c < C) ? xb[(size_t)c * T + t] * im : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + 8 * i;
      if (c < C) {
        xs[c * 33 + lane] = v[i];
        if (xw && tv) xw[(size_t)b * C * T + (size_t)c * T + t] = v[i];
        part += v[i];
      }
    }
  }
  red[grp * 32 + lane] = part;
  __syncthreads();
  if (grp == 0) {
    float m = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) m += red[g * 32 + lane];
    mean_s[lane] = m / (float)C;
  }
  __syncthreads();
  const float mean = mean_s[lane];
  part = 0.f;
  for (int c = grp; c < C; c += 8) {
    const float d = xs[c * 33 + lane] - mean;
    part = fmaf(d, d, part);
  }
  red[grp * 32 + lane] = part;
  __syncthreads();
  if (grp == 0) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) v += red[g * 32 + lane];
    rstd_s[lane] = rsqrtf(v / (float)C + eps);
  }
  __syncthreads();
  const float rstd = rstd_s[lane];
#pragma unroll 4
  for (int c = grp; c < C; c += 8) {
    float v = (xs[c * 33 + lane] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    if (relu) v = fmaxf(v, 0.f);
    v *= om;
    xs[c * 33 + lane] = v;
    if (y && tv) y[(size_t)b * C * T + (size_t)c * T + t] = v;
  }
  if (po.hi) {
    __syncthreads();
    for (int sl = grp; sl < C / 8; sl += 8) {
      if (!tv) continue;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = xs[(sl * 8 + e) * 33 + lane];
      store_slab(po, b, C, sl, t, v);
    }
    if (po.zero_halo && blockIdx.x == 0) zero_halo_rows(po, b, C, T);
  }
}

cudaError_t channel_layernorm_planes(const float* x, float* xw, float* y, const float* gamma, const float* beta, float eps,
                                     const float* in_mask, const float* out_mask, int B, int C, int T, const PlaneOut& po,
                                     cudaStream_t s) {
  if (po.hi && C % 8) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)C * 33 + 256 + 64) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidConfiguration;
  dim3 grid(cdiv(T, 32), B);
  channel_ln_tiled_kernel<<<grid, 256, smem, s>>>(x, xw, y, gamma, beta, eps, in_mask, out_mask, C, T, po, 0);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Self attention for short sequences (T <= 64) with q, k, v of one (batch item, head) resident in shared memory:
// scores = q.k/sqrt(dk), masked_fill(mask_t*mask_s == 0, -1e4), softmax over s, out = p.v
// (rel_transformer_encoder.py:128-158, window_size None).  Output to [B,C,T] and/or operand planes.
__global__ void __launch_bounds__(256) self_attn_small_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                              const float* __restrict__ v,
                                                              const float* __restrict__ mask, float* __restrict__ out,
                                                              int C, int T, int heads, long bs, PlaneOut po) {
  extern __shared__ float sm[];                      // qs[dk][T], ks[dk][T], vs[dk][T], p[T][T+1]
  griddep_launch_if_resident();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dk = C / heads;
  float* qs = sm;
  float* ks = qs + (size_t)dk * T;
  float* vs = ks + (size_t)dk * T;
  float* ps = vs + (size_t)dk * T;
  const size_t base = (size_t)b * bs + (size_t)h * dk * T;
#pragma unroll 4
  for (int i = threadIdx.x; i < dk * T; i += blockDim.x) {
    const float a = __ldg(q + base + i), c = __ldg(k + base + i), d = __ldg(v + base + i);
    qs[i] = a;
    ks[i] = c;
    vs[i] = d;
  }
  __syncthreads();
  const float* mb = mask + (size_t)b * T;
  const float inv = rsqrtf((float)dk);
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int t = i / T, s2 = i - t * T;
    float acc = 0.f;
    for (int d = 0; d < dk; ++d) acc = fmaf(qs[d * T + t], ks[d * T + s2], acc);
    acc *= inv;
    if (mb[t] * mb[s2] == 0.f) acc = -1e4f;
    ps[t * (T + 1) + s2] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = warp; t < T; t += nw) {               // softmax of row t by one warp
    float* pr = ps + t * (T + 1);
    float mx = -INFINITY;
    for (int s2 = lane; s2 < T; s2 += 32) mx = fmaxf(mx, pr[s2]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s2 = lane; s2 < T; s2 += 32) {
      const float e = expf(pr[s2] - mx);
      pr[s2] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float rs = 1.f / sum;
    for (int s2 = lane; s2 < T; s2 += 32) pr[s2] *= rs;
  }
  __syncthreads();
  // out[d, t] = sum_s p[t, s] v[d, s]; results overwrite qs (q is dead) so the plane store can read 8 channels per thread
  for (int i = threadIdx.x; i < dk * T; i += blockDim.x) {
    const int d = i / T, t = i - d * T;
    float acc = 0.f;
    const float* pr = ps + t * (T + 1);
    for (int s2 = 0; s2 < T; ++s2) acc = fmaf(pr[s2], vs[d * T + s2], acc);
    qs[i] = acc;
    if (out) out[(size_t)b * C * T + (size_t)h * dk * T + i] = acc;
  }
  if (po.hi) {
    __syncthreads();
    for (int i = threadIdx.x; i < (dk / 8) * T; i += blockDim.x) {
      const int sl = i / T, t = i - sl * T;
      float v8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v8[e] = qs[(sl * 8 + e) * T + t];
      store_slab(po, b, C, h * (dk / 8) + sl, t, v8);
    }
  }
}

// q, k, v of one (item, head) and the T x T scores must fit in shared memory: T <= 128 at d_k = 96 (213 KB)
bool self_attention_planes_fits(int C, int T, int heads) {
  const size_t dk = (size_t)(C / (heads > 0 ? heads : 1));
  return T > 0 && T <= 128 && (3 * dk * T + (size_t)T * (T + 1)) * sizeof(float) <= (size_t)227 * 1024;
}
cudaError_t self_attention_planes(const float* q, const float* k, const float* v, const float* mask, float* out, int B,
                                  int C, int T, int heads, const PlaneOut& po, cudaStream_t s) {
  const int dk = C / heads;
  if (!self_attention_planes_fits(C, T, heads) || (po.hi && dk % 8)) return cudaErrorInvalidValue;
  const size_t smem = ((size_t)3 * dk * T + (size_t)T * (T + 1)) * sizeof(float);
  static unsigned long long attr_done = 0;
  {
    const cudaError_t e = ensure_max_dyn_smem(self_attn_small_kernel, 227 * 1024, &attr_done);
    if (e != cudaSuccess) return e;
  }
  self_attn_small_kernel<<<B * heads, 256, smem, s>>>(q, k, v, mask, out, C, T, heads, (long)3 * C * T, po);
  return cudaGetLastError();
}

__global__ void apply_mask_kernel(float* x, const float* __restrict__ mask, int C, int T, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t b = i / ((size_t)C * T);
  const int t = (int)(i % T);
  x[i] *= mask[b * T + t];
}
cudaError_t apply_mask(float* x, const float* mask, int B, int C, int T, cudaStream_t s) {
  const size_t n = (size_t)B * C * T;
  apply_mask_kernel<<<cdiv(n, 256), 256, 0, s>>>(x, mask, C, T, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Self attention for short sequences: one warp per query row, block = (b, head).  scores = q.k/sqrt(dk),
// masked_fill(mask_t*mask_s == 0, -1e4), softmax over s, out = p.v  (rel_transformer_encoder.py:128-158, window None).
__global__ void self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                 const float* __restrict__ mask, float* __restrict__ out, int C, int T, int heads,
                                 long bs) {
  extern __shared__ float sm[];             // [warps][T] probabilities
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dk = C / heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* qb = q + (size_t)b * bs + (size_t)h * dk * T;
  const float* kb = k + (size_t)b * bs + (size_t)h * dk * T;
  const float* vb = v + (size_t)b * bs + (size_t)h * dk * T;
  float* ob = out + (size_t)b * C * T + (size_t)h * dk * T;
  const float* mb = mask + (size_t)b * T;
  float* pr = sm + (size_t)warp * T;
  const float inv = rsqrtf((float)dk);
  for (int t = warp; t < T; t += nw) {
    const float mt = mb[t];
    float mx = -INFINITY;
    for (int s = lane; s < T; s += 32) {
      float acc = 0.f;
      for (int d = 0; d < dk; ++d) acc = fmaf(qb[(size_t)d * T + t], kb[(size_t)d * T + s], acc);
      acc *= inv;
      if (mt * mb[s] == 0.f) acc = -1e4f;
      pr[s] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < T; s += 32) {
      const float e = expf(pr[s] - mx);
      pr[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float rs = 1.f / sum;
    __syncwarp();
    for (int d = lane; d < dk; d += 32) {
      float acc = 0.f;
      for (int s = 0; s < T; ++s) acc = fmaf(pr[s], vb[(size_t)d * T + s], acc);
      ob[(size_t)d * T + t] = acc * rs;
    }
    __syncwarp();
  }
}

cudaError_t self_attention(const float* q, const float* k, const float* v, const float* mask, float* out, int B, int C,
                           int T, int heads, cudaStream_t s) {
  const int threads = 256;
  const size_t smem = (size_t)(threads / 32) * T * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  // q, k, v are slices of one [B, 3C, T] buffer: batch stride 3*C*T
  self_attn_kernel<<<B * heads, threads, smem, s>>>(q, k, v, mask, out, C, T, heads, (long)3 * C * T);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// S2PA, folded form: logits[l] = keys[b,t,l,:] . qk[b,t,:]  (qk = W_k^T (W_q x) * D^-1/2), masked softmax over the
// gloss tokens of this character, ctx = sum_l w[l] * values[b,t,l,:].  HBM-bound, and ONE pass over the gloss rows:
// every warp keeps a running (max, sum, weighted row sum) over the rows it owns (online softmax) and the eight partial
// results are merged at the end, so a key row is used for its logit and -- when `values` is the same tensor, as in the
// binarized data -- for the weighted sum from the same registers: 3 072 B per valid gloss token instead of 6 144 B.
// Rows with key_map == 0 (logit forced to -1e9 -> weight exactly 0 unless the whole row is masked) are not read at all;
// 128-bit streaming loads, two rows in flight per warp.  The attention weights themselves are the exact softmax of the
// stored logits.  (The two-pass predecessor of this kernel read every row twice and idled at the softmax barrier in
// between: 79 us alone, 122 us inside the step at cfg 2.)
template <bool ALIASED>
__global__ void __launch_bounds__(256, 2) s2pa_stream_kernel(const float* __restrict__ keys,
                                                           const float* __restrict__ values,
                                                           const float* __restrict__ key_map,
                                                           const float* __restrict__ qk, int Tw, int Lk, int D,
                                                           float* __restrict__ weights, float* __restrict__ align,
                                                           float* __restrict__ ctx,
                                                           const int64_t* __restrict__ row_off,
                                                           const int32_t* __restrict__ row_len) {
  // row_off != null: keys / values are a dictionary BANK [rows][D]; character (b,t) owns rows
  // [row_off[bt], row_off[bt] + row_len[bt]) of it and every further gloss position is an all-zero row (never read).
  extern __shared__ float sm[];
  float* s_q = sm;                  // [D]
  float* s_w = s_q + D;             // [Lk] logits, then weights
  float* s_acc = s_w + ((Lk + 3) & ~3);   // [8][D] per-warp weighted row sums (16-byte aligned: float4 stores)
  __shared__ float s_m[8], s_s[8];
  constexpr int R = 6;              // float4 per lane and row: D <= 768 (checked by the launcher)
  const int bt = blockIdx.x;
  const int b = bt / Tw, t = bt - b * Tw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  griddep_launch_if_resident();
  for (int d = tid; d < D; d += 256) s_q[d] = qk[((size_t)b * D + d) * Tw + t];
  __syncthreads();
  const float* km = key_map + (size_t)bt * Lk;
  const int nrow = row_off ? row_len[bt] : Lk;                 // rows that exist in memory
  const size_t row0 = row_off ? (size_t)(nrow > 0 ? row_off[bt] : 0) : (size_t)bt * Lk;
  const int D4 = D >> 2;
  const float4* kp = reinterpret_cast<const float4*>(keys + row0 * D);
  const float4* vp = reinterpret_cast<const float4*>(values + row0 * D);
  float m = -INFINITY, ssum = 0.f;
  float4 acc[R];
#pragma unroll
  for (int u = 0; u < R; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_row = [&](const float4* base, int l, float4 (&dst)[R]) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = lane + 32 * u;
      dst[u] = (l < nrow && i < D4) ? ld_stream_f4(base + (size_t)l * D4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto consume = [&](int l, const float4 (&kr)[R], const float4 (&vr)[R]) {
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = lane + 32 * u;
      if (i < D4) {                                               // q stays in shared memory: registers hold rows in flight
        const float4 q4 = *reinterpret_cast<const float4*>(s_q + 4 * i);
        dot = fmaf(kr[u].x, q4.x, dot); dot = fmaf(kr[u].y, q4.y, dot);
        dot = fmaf(kr[u].z, q4.z, dot); dot = fmaf(kr[u].w, q4.w, dot);
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) s_w[l] = dot;
    const float mn = fmaxf(m, dot);
    const float sc = expf(m - mn), pw = expf(dot - mn);          // first row: exp(-inf) = 0
    ssum = fmaf(ssum, sc, pw);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      acc[u].x = fmaf(pw, vr[u].x, acc[u].x * sc); acc[u].y = fmaf(pw, vr[u].y, acc[u].y * sc);
      acc[u].z = fmaf(pw, vr[u].z, acc[u].z * sc); acc[u].w = fmaf(pw, vr[u].w, acc[u].w * sc);
    }
    m = mn;
  };
  // rows l = warp, warp + 8, ...; two unmasked rows of this warp are in flight at a time
  int any = 0;
  for (int l0 = warp; l0 < Lk; l0 += 16) {
    const int l1 = l0 + 8;
    const bool u0 = km[l0] != 0.f, u1 = l1 < Lk && km[l1] != 0.f;
    float4 k0[R], k1[R];
    if (u0) load_row(kp, l0, k0);
    if (u1) load_row(kp, l1, k1);
    if (ALIASED) {
      if (u0) consume(l0, k0, k0); else if (lane == 0) s_w[l0] = -1e9f;
      if (u1) consume(l1, k1, k1); else if (lane == 0 && l1 < Lk) s_w[l1] = -1e9f;
    } else {
      float4 v0[R];
      if (u0) { load_row(vp, l0, v0); consume(l0, k0, v0); } else if (lane == 0) s_w[l0] = -1e9f;
      if (u1) { load_row(vp, l1, v0); consume(l1, k1, v0); } else if (lane == 0 && l1 < Lk) s_w[l1] = -1e9f;
    }
    any |= (u0 || u1);
  }
  if (lane == 0) { s_m[warp] = m; s_s[warp] = ssum; }
  __syncthreads();
  // merge the eight partial softmaxes
  float M = s_m[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) M = fmaxf(M, s_m[i]);
  const bool all_masked = M == -INFINITY;                        // no unmasked gloss token at all
  if (!all_masked) {
    float S = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) S += s_s[i] * expf(s_m[i] - M);   // exp(-inf) = 0 for a warp without rows
    const float rs = 1.f / S;
    const float mine = any ? expf(m - M) : 0.f;
    float* dst = s_acc + (size_t)warp * D;
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int i = lane + 32 * u;
      if (i < D4)
        *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(acc[u].x * mine, acc[u].y * mine, acc[u].z * mine, acc[u].w * mine);
    }
    for (int l = tid; l < Lk; l += 256) {
      const float w = expf(s_w[l] - M) * rs;                     // masked: exp(-1e9 - M) = 0
      weights[(size_t)bt * Lk + l] = w;
      align[((size_t)b * Lk + l) * Tw + t] = w;                  // [B,1,Lk,Tw]
    }
    __syncthreads();
    for (int d = tid; d < D; d += 256) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) v += s_acc[(size_t)i * D + d];
      ctx[((size_t)b * D + d) * Tw + t] = v * rs;
    }
    return;
  }
  // Fully masked character (padding): softmax of Lk equal logits = 1/Lk each, ctx = mean over ALL Lk positions of the
  // value rows (zero rows where nothing exists) -- still read, like the reference does.
  const float w = 1.f / (float)Lk;
  for (int l = tid; l < Lk; l += 256) {
    weights[(size_t)bt * Lk + l] = w;
    align[((size_t)b * Lk + l) * Tw + t] = w;
  }
  for (int i = tid; i < D4; i += 256) {
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < nrow; ++l) {
      const float4 vv = ld_stream_f4(vp + (size_t)l * D4 + i);
      a4.x = fmaf(w, vv.x, a4.x); a4.y = fmaf(w, vv.y, a4.y); a4.z = fmaf(w, vv.z, a4.z); a4.w = fmaf(w, vv.w, a4.w);
    }
    float* c = ctx + ((size_t)b * D + 4 * i) * Tw + t;
    c[0] = a4.x; c[Tw] = a4.y; c[2 * (size_t)Tw] = a4.z; c[3 * (size_t)Tw] = a4.w;
  }
}

cudaError_t s2pa_stream(const float* keys, const float* values, const float* key_map, const float* qk, int B, int Tw,
                        int Lk, int D, float* weights, float* align, float* ctx, cudaStream_t s,
                        const int64_t* row_off, const int32_t* row_len) {
  if (D % 4 || D > 768) return cudaErrorInvalidValue;           // 6 x 128 bit per lane and row
  const size_t smem = (size_t)(9 * D + ((Lk + 3) & ~3)) * sizeof(float);
  const bool aliased = keys == values;
  auto kern = aliased ? s2pa_stream_kernel<true> : s2pa_stream_kernel<false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<B * Tw, 256, smem, s>>>(keys, values, key_map, qk, Tw, Lk, D, weights, align, ctx, row_off, row_len);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// S2PA as the reference computes it (s2pa_route = 1, dict_encoder.py:40-58): kv [B][2H][Tw*Lk] holds k = W_k keys
// (channels [0,H)) and v = W_v values (channels [H,2H)) of EVERY gloss token, written by the tcgen05 projection GEMM;
// q [B][H][Tw] is W_q x * dict_dim^-1/2.  One block per character: logits[l] = k[:, l] . q, masked_fill(key_map == 0,
// -1e9), softmax over l, ctx[c] = sum_l w[l] * v[c, l].  Reads are coalesced along l (the fastest axis of kv).
__global__ void __launch_bounds__(256) s2pa_attend_kernel(const float* __restrict__ kv, const float* __restrict__ q,
                                                           const float* __restrict__ key_map, int Tw, int Lk, int H,
                                                           float* __restrict__ weights, float* __restrict__ align,
                                                           float* __restrict__ ctx) {
  extern __shared__ float sm[];
  float* s_q = sm;            // [H]
  float* s_w = sm + H;        // [Lk]
  __shared__ float s_red[8];
  const int bt = blockIdx.x;
  const int b = bt / Tw, t = bt - b * Tw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t TL = (size_t)Tw * Lk;
  const float* kb = kv + (size_t)b * 2 * H * TL + (size_t)t * Lk;      // k[c][l] = kb[c*TL + l]
  const float* vb = kb + (size_t)H * TL;
  for (int c = tid; c < H; c += 256) s_q[c] = q[((size_t)b * H + c) * Tw + t];
  __syncthreads();
  const float* km = key_map + (size_t)bt * Lk;
  for (int l = tid; l < Lk; l += 256) {
    float acc = 0.f;
    for (int c = 0; c < H; ++c) acc = fmaf(__ldg(kb + (size_t)c * TL + l), s_q[c], acc);
    s_w[l] = km[l] != 0.f ? acc : -1e9f;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int l = tid; l < Lk; l += 256) mx = fmaxf(mx, s_w[l]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int l = tid; l < Lk; l += 256) {
    const float e = expf(s_w[l] - mx);
    s_w[l] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += s_red[i];
  const float rs = 1.f / sum;
  for (int l = tid; l < Lk; l += 256) {
    const float w = s_w[l] * rs;
    s_w[l] = w;
    weights[(size_t)bt * Lk + l] = w;
    align[((size_t)b * Lk + l) * Tw + t] = w;                 // [B,1,Lk,Tw]
  }
  __syncthreads();
  // ctx[c] = sum_l w[l] * v[c][l]: one warp per channel, lanes along l
  for (int c = warp; c < H; c += 8) {
    float acc = 0.f;
    for (int l = lane; l < Lk; l += 32) acc = fmaf(s_w[l], __ldg(vb + (size_t)c * TL + l), acc);
    acc = warp_sum(acc);
    if (lane == 0) ctx[((size_t)b * H + c) * Tw + t] = acc;
  }
}

cudaError_t s2pa_attend(const float* kv, const float* q, const float* key_map, int B, int Tw, int Lk, int H,
                        float* weights, float* align, float* ctx, cudaStream_t s) {
  const size_t smem = (size_t)(H + Lk) * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;
  s2pa_attend_kernel<<<B * Tw, 256, smem, s>>>(kv, q, key_map, Tw, Lk, H, weights, align, ctx);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Dictionary bank -> the small per-batch tensors of dict_msg, exactly as DictTTSDataset.collater pads them
// (tasks/tts/dataset_utils.py:264-302): id >= 0: the entry's key_map / pinyin / pinyin_map, zero padded; id == -1
// (BOS / EOS row): key_map 1, pinyin 0, pinyin_map 1 over the whole row; id == -2 (padding): zeros.  Also the row
// window of the big keys / values bank for s2pa_stream.  One block per (b, t).
__global__ void dict_bank_gather_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tok_off,
                                        const int64_t* __restrict__ pin_off, const float* __restrict__ bank_key_map,
                                        const int64_t* __restrict__ bank_pinyin,
                                        const int64_t* __restrict__ bank_pinyin_map, int n_entries, int Lk, int Lp,
                                        float* __restrict__ key_map, int64_t* __restrict__ pinyin,
                                        int64_t* __restrict__ pinyin_map, int64_t* __restrict__ row_off,
                                        int32_t* __restrict__ row_len, int* __restrict__ err) {
  const int bt = blockIdx.x;
  const int64_t id = ids[bt];
  int64_t t0 = 0, p0 = 0;
  int nl = 0, np = 0;
  if (id >= 0) {
    if (id >= n_entries) {
      if (threadIdx.x == 0) atomicOr(err, 1);
    } else {
      t0 = tok_off[id]; nl = (int)(tok_off[id + 1] - t0);
      p0 = pin_off[id]; np = (int)(pin_off[id + 1] - p0);
      if (nl > Lk || np > Lp) {
        if (threadIdx.x == 0) atomicOr(err, 2);
        nl = min(nl, Lk); np = min(np, Lp);
      }
    }
  }
  const float edge = id == -1 ? 1.f : 0.f;
  for (int l = threadIdx.x; l < Lk; l += blockDim.x) key_map[(size_t)bt * Lk + l] = l < nl ? bank_key_map[t0 + l] : edge;
  for (int p = threadIdx.x; p < Lp; p += blockDim.x) {
    pinyin[(size_t)bt * Lp + p] = p < np ? bank_pinyin[p0 + p] : 0;
    pinyin_map[(size_t)bt * Lp + p] = p < np ? bank_pinyin_map[p0 + p] : (int64_t)edge;
  }
  if (threadIdx.x == 0) { row_off[bt] = t0; row_len[bt] = nl; }
}

cudaError_t dict_bank_gather(const int64_t* ids, const int64_t* tok_off, const int64_t* pin_off,
                             const float* bank_key_map, const int64_t* bank_pinyin, const int64_t* bank_pinyin_map,
                             int n_entries, int B, int Tw, int Lk, int Lp, float* key_map, int64_t* pinyin,
                             int64_t* pinyin_map, int64_t* row_off, int32_t* row_len, int* err, cudaStream_t s) {
  dict_bank_gather_kernel<<<B * Tw, 128, 0, s>>>(ids, tok_off, pin_off, bank_key_map, bank_pinyin, bank_pinyin_map,
                                                 n_entries, Lk, Lp, key_map, pinyin, pinyin_map, row_off, row_len, err);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
__global__ void dict_maxes_kernel(const float* __restrict__ key_map, size_t n_key, const int64_t* __restrict__ pm,
                                  size_t n_pin, int* maxes) {
  int m0 = 0, m1 = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_key; i += stride) m0 = max(m0, (int)key_map[i]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pin; i += stride) m1 = max(m1, (int)pm[i]);
  m0 = warp_max_i(m0);
  m1 = warp_max_i(m1);
  if ((threadIdx.x & 31) == 0) {
    atomicMax(maxes, m0);
    atomicMax(maxes + 1, m1);
  }
}
cudaError_t dict_maxes(const float* key_map, size_t n_key, const int64_t* pinyin_map, size_t n_pin, int* maxes,
                       cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(maxes, 0, 2 * sizeof(int), s);
  if (e != cudaSuccess) return e;
  dict_maxes_kernel<<<148, 256, 0, s>>>(key_map, n_key, pinyin_map, n_pin, maxes);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// mask_weights_attn + add_pron_rule + pinyin mixing (layers/utils.py:49-58,109-115; dict_encoder.py:60-64).
__global__ void s2pa_pron_kernel(const float* __restrict__ weights, const float* __restrict__ key_map,
                                 const int64_t* __restrict__ pinyin, const int64_t* __restrict__ pinyin_map,
                                 const int64_t* __restrict__ pron_modified, const int* __restrict__ maxes,
                                 const float* __restrict__ pinyin_emb, int pinyin_vocab,
                                 const float* __restrict__ context, const float* __restrict__ seq_mask, int Tw, int Lk,
                                 int Lp, int H, int apply_rule, float* __restrict__ pron_attn, float* __restrict__ x2) {
  extern __shared__ float s_pw[];            // [Lp]
  const int bt = blockIdx.x;
  const int b = bt / Tw, t = bt - b * Tw;
  const int kmax = maxes[0], pmax = maxes[1];
  const float* w = weights + (size_t)bt * Lk;
  const float* km = key_map + (size_t)bt * Lk;
  const int64_t* pm = pinyin_map + (size_t)bt * Lp;
  const long forced = (apply_rule && pron_modified) ? (long)pron_modified[bt] : 0;
  for (int p = threadIdx.x; p < Lp; p += blockDim.x) {
    const long id = pm[p];
    float acc = 0.f;
    if (id >= 1 && id <= kmax) {
      const float idf = (float)id;
      for (int l = 0; l < Lk; ++l) acc += (km[l] == idf) ? w[l] : 0.f;
    }
    float r = acc;
    if (forced >= 1 && forced <= pmax) {
      const float oh = (id == forced) ? 1.f : 0.f;
      r = (oh - acc) + acc;                                    // weights_ - weights.detach() + weights
    } else if (apply_rule && pron_modified) {
      r = (acc - acc) + acc;
    }
    s_pw[p] = r;
    pron_attn[(size_t)bt * Lp + p] = r;
  }
  __syncthreads();
  const float mk = seq_mask[bt];
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < Lp; ++p) {
      long id = pinyin[(size_t)bt * Lp + p];
      if (id < 0 || id >= pinyin_vocab) id = 0;
      acc = fmaf(s_pw[p], pinyin_emb[(size_t)id * H + h], acc);
    }
    const size_t o = ((size_t)b * H + h) * Tw + t;
    x2[o] = context[o] * mk + acc;
  }
}

cudaError_t s2pa_pron(const float* weights, const float* key_map, const int64_t* pinyin, const int64_t* pinyin_map,
                      const int64_t* pron_modified, const int* maxes, const float* pinyin_emb, int pinyin_vocab,
                      const float* context, const float* seq_mask, int B, int Tw, int Lk, int Lp, int H, int apply_rule,
                      float* pron_attn, float* x2, cudaStream_t s) {
  s2pa_pron_kernel<<<B * Tw, 192, Lp * sizeof(float), s>>>(weights, key_map, pinyin, pinyin_map, pron_modified, maxes,
                                                          pinyin_emb, pinyin_vocab, context, seq_mask, Tw, Lk, Lp, H,
                                                          apply_rule, pron_attn, x2);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
__global__ void finish_text_kernel(const float* __restrict__ x, const float* __restrict__ tok_mask, int Tw, int H,
                                   float* __restrict__ enc_btc, float* __restrict__ dur_in, float* __restrict__ keep) {
  const int bt = blockIdx.x;
  const int b = bt / Tw, t = bt - b * Tw;
  const float mk = tok_mask[bt];
  __shared__ float s_red[8];
  float a = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const size_t i = ((size_t)b * H + h) * Tw + t;
    const float v = x[i] * mk;
    enc_btc[(size_t)bt * H + h] = v;
    dur_in[i] = v;
    a += fabsf(v);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_red[i];
    keep[bt] = tot == 0.f ? 0.f : 1.f;                         // src_padding = (abs().sum(-1) == 0)
  }
}
cudaError_t finish_text(const float* x, const float* tok_mask, int B, int Tw, int H, float* enc_btc, float* dur_in,
                        float* keep, cudaStream_t s) {
  finish_text_kernel<<<B * Tw, 64, 0, s>>>(x, tok_mask, Tw, H, enc_btc, dur_in, keep);
  return cudaGetLastError();
}

__global__ void count_keep_kernel(const float* __restrict__ keep, int Tw, int64_t* ilens) {
  const int b = blockIdx.x;
  float c = 0.f;
  for (int t = threadIdx.x; t < Tw; t += 32) c += keep[(size_t)b * Tw + t];
  c = warp_sum(c);
  if (threadIdx.x == 0) ilens[b] = (int64_t)(c + 0.5f);
}
cudaError_t count_keep(const float* keep, int B, int Tw, int64_t* ilens, cudaStream_t s) {
  count_keep_kernel<<<B, 32, 0, s>>>(keep, Tw, ilens);
  return cudaGetLastError();
}

__global__ void dur_head_kernel(const float* __restrict__ xs, const float* __restrict__ w,
                                const float* __restrict__ bias, const float* __restrict__ keep, int C, int Tw,
                                float* __restrict__ dur, int64_t* __restrict__ dur_int, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = i / Tw, t = i - b * Tw;
  float acc = 0.f;
#pragma unroll 16                                             // sixteen loads in flight; the fmaf chain keeps its order
  for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w + c), __ldg(xs + ((size_t)b * C + c) * Tw + t), acc);
  acc += bias[0];
  const float sp = acc > 20.f ? acc : log1pf(expf(acc));        // nn.Softplus(beta=1, threshold=20)
  const float d = sp * keep[i];
  dur[i] = d;
  float r = rintf(expf(d) - 1.f);                               // torch.round = half-to-even
  r = fmaxf(r, 0.f);
  dur_int[i] = (int64_t)r;
}
cudaError_t dur_head(const float* xs, const float* w, const float* bias, const float* keep, int B, int C, int Tw,
                     float* dur, int64_t* dur_int, cudaStream_t s) {
  const int n = B * Tw;
  dur_head_kernel<<<cdiv(n, 128), 128, 0, s>>>(xs, w, bias, keep, C, Tw, dur, dur_int, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
__global__ void wn_gate_kernel(const float* __restrict__ a, float* __restrict__ acts, int H, int T, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t per = (size_t)H * T;
  const size_t b = i / per, r = i - b * per;
  const float ta = a[b * 2 * per + r];
  const float sa = a[b * 2 * per + per + r];
  acts[i] = tanhf(ta) * (1.f / (1.f + expf(-sa)));
}
cudaError_t wn_gate(const float* a, float* acts, int B, int H, int T, cudaStream_t s) {
  const size_t n = (size_t)B * H * T;
  wn_gate_kernel<<<cdiv(n, 256), 256, 0, s>>>(a, acts, H, T, n);
  return cudaGetLastError();
}

// Same gate, written straight into tensor-core operand planes: thread = (8-channel slab, t).
__global__ void wn_gate_planes_kernel(const float* __restrict__ a, int H, int T, PlaneOut po) {
  griddep_launch_if_resident();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int sl = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float* ab = a + (size_t)b * 2 * H * T;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float ta = ab[(size_t)(sl * 8 + e) * T + t];
    const float sa = ab[(size_t)(H + sl * 8 + e) * T + t];
    v[e] = tanhf(ta) * (1.f / (1.f + expf(-sa)));
  }
  store_slab(po, b, H, sl, t, v);
}
cudaError_t wn_gate_planes(const float* a, int B, int H, int T, const PlaneOut& po, cudaStream_t s) {
  if (H % 8 || !po.hi) return cudaErrorInvalidValue;
  dim3 grid(cdiv(T, 128), H / 8, B);
  wn_gate_planes_kernel<<<grid, 128, 0, s>>>(a, H, T, po);
  return cudaGetLastError();
}

// FVAE decoder pre_net: block = (32 latent positions = 128 output frames, batch item), 256 threads = 8 warps, lane = latent
// position, warp w takes the 8-channel slabs w, w + 8, ...  A thread computes ALL FOUR phases of its position, so that
// what it stores are whole 32-byte sectors -- four consecutive fp32 samples of a channel, two consecutive 16-byte plane
// rows -- (a phase per warp wrote quarter sectors: L2 read every sector back from HBM to merge, 41 us; a phase per lane
// with a rolled loop was a chain of dependent cache misses, 64 us); weight reads are warp-uniform broadcasts.
__device__ __forceinline__ void st_rows2(tc16* p, const uint32_t (&a)[4], const uint32_t (&b)[4]) {   // 32 B, 32-B aligned
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]),
               "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3])
               : "memory");
}
template <int CI>
__global__ void __launch_bounds__(256) fvae_pre_net_planes_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                                  const float* __restrict__ bias, int C_out, int Tq,
                                                                  float* __restrict__ out, PlaneOut po) {
  extern __shared__ float4 wsm4[];                   // the whole weight [4][CI][C_out]: fetched once, all loads in flight
  griddep_launch_if_resident();
  const int b = blockIdx.y, q = blockIdx.x * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5;
  const int T = 4 * Tq;
  {
    const float4* wg = reinterpret_cast<const float4*>(w);
    for (int i = threadIdx.x; i < CI * C_out; i += 256) wsm4[i] = __ldg(wg + i);      // 4 * CI * C_out floats
  }
  float zv[CI];
#pragma unroll
  for (int ci = 0; ci < CI; ++ci) zv[ci] = q < Tq ? z[((size_t)b * CI + ci) * Tq + q] : 0.f;
  __syncthreads();
  const float* ws = reinterpret_cast<const float*>(wsm4);
  const int slabs = C_out / 8;
  for (int sl = warp; sl < slabs; sl += 8) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + sl * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + sl * 8 + 4));
    float v[4][8];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      const float* wp = ws + (size_t)ph * CI * C_out + sl * 8;
      float4 w0[CI], w1[CI];
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        w0[ci] = *reinterpret_cast<const float4*>(wp + (size_t)ci * C_out);
        w1[ci] = *reinterpret_cast<const float4*>(wp + (size_t)ci * C_out + 4);
      }
      v[ph][0] = b0.x; v[ph][1] = b0.y; v[ph][2] = b0.z; v[ph][3] = b0.w;
      v[ph][4] = b1.x; v[ph][5] = b1.y; v[ph][6] = b1.z; v[ph][7] = b1.w;
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        v[ph][0] = fmaf(w0[ci].x, zv[ci], v[ph][0]); v[ph][1] = fmaf(w0[ci].y, zv[ci], v[ph][1]);
        v[ph][2] = fmaf(w0[ci].z, zv[ci], v[ph][2]); v[ph][3] = fmaf(w0[ci].w, zv[ci], v[ph][3]);
        v[ph][4] = fmaf(w1[ci].x, zv[ci], v[ph][4]); v[ph][5] = fmaf(w1[ci].y, zv[ci], v[ph][5]);
        v[ph][6] = fmaf(w1[ci].z, zv[ci], v[ph][6]); v[ph][7] = fmaf(w1[ci].w, zv[ci], v[ph][7]);
      }
    }
    if (q < Tq) {
      float4* op = reinterpret_cast<float4*>(out + ((size_t)b * C_out + sl * 8) * T + 4 * q);   // T % 4 == 0: 16-B aligned
#pragma unroll
      for (int e = 0; e < 8; ++e) op[(size_t)e * (T / 4)] = make_float4(v[0][e], v[1][e], v[2][e], v[3][e]);
      if (po.hi) {
        uint32_t hw[4][4], lw[4][4];
#pragma unroll
        for (int ph = 0; ph < 4; ++ph)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (po.lo) split2(v[ph][2 * e], v[ph][2 * e + 1], po.fmt, hw[ph][e], lw[ph][e]);
            else hw[ph][e] = pack2(v[ph][2 * e], v[ph][2 * e + 1], po.fmt);
          }
        const size_t off = (((size_t)b * slabs + sl) * po.rows + po.pad + 4 * q) * 8;     // pad % 2 == 0: 32-B aligned rows
        st_rows2(po.hi + off, hw[0], hw[1]);
        st_rows2(po.hi + off + 16, hw[2], hw[3]);
        if (po.lo) {
          st_rows2(po.lo + off, lw[0], lw[1]);
          st_rows2(po.lo + off + 16, lw[2], lw[3]);
        }
      }
    }
  }
  if (po.hi && po.zero_halo && blockIdx.x == 0) zero_halo_rows(po, b, C_out, T);
}
cudaError_t fvae_pre_net_planes(const float* z, const float* w, const float* bias, int B, int C_in, int C_out, int Tq,
                                float* out, const PlaneOut& po, cudaStream_t s) {
  if ((C_in != 16 && C_in != 8) || C_out % 8 || B <= 0 || Tq <= 0 || (po.hi && ((po.pad & 1) || (po.rows & 1))))
    return cudaErrorInvalidValue;
  dim3 grid(cdiv(Tq, 32), B);
  const size_t smem = (size_t)4 * C_in * C_out * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidConfiguration;
  if (C_in == 16) fvae_pre_net_planes_kernel<16><<<grid, 256, smem, s>>>(z, w, bias, C_out, Tq, out, po);
  else fvae_pre_net_planes_kernel<8><<<grid, 256, smem, s>>>(z, w, bias, C_out, Tq, out, po);
  return cudaGetLastError();
}

// 1x1 convolution with a handful of channels on one side (the 8 <-> 64 channel pre / post projections of a coupling
// layer, glow_modules.py:113-127): out[b,co,t] = alpha * (sum_ci w[ci][co] * x[b,ci,t] + bias[co]) + res[b,co,t].
// One thread per (b, t) walks the output channels eight at a time; w is the packed [C_in][C_out] layout of ConvW.
// (The generic register-tiled kernel spent 13 / 49 us on these 3-MFLOP layers.)
__global__ void __launch_bounds__(128) pointwise_small_kernel(const float* __restrict__ x, long x_bs,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ bias, int C_in, int C_out,
                                                              int T, float alpha, const float* __restrict__ res,
                                                              long r_bs, float* __restrict__ out, long o_bs) {
  griddep_launch_if_resident();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* xb = x + (size_t)b * x_bs + t;
  for (int c0 = 0; c0 < C_out; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int ci0 = 0; ci0 < C_in; ci0 += 16) {                   // sixteen independent loads in flight per thread: the
      float xv[16];                                              // layer is pure load latency (6 000 positions in all)
#pragma unroll
      for (int u = 0; u < 16; ++u) xv[u] = ci0 + u < C_in ? xb[(size_t)(ci0 + u) * T] : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (ci0 + u < C_in) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ci0 + u) * C_out + c0));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(ci0 + u) * C_out + c0 + 4));
          acc[0] = fmaf(w0.x, xv[u], acc[0]); acc[1] = fmaf(w0.y, xv[u], acc[1]); acc[2] = fmaf(w0.z, xv[u], acc[2]);
          acc[3] = fmaf(w0.w, xv[u], acc[3]); acc[4] = fmaf(w1.x, xv[u], acc[4]); acc[5] = fmaf(w1.y, xv[u], acc[5]);
          acc[6] = fmaf(w1.z, xv[u], acc[6]); acc[7] = fmaf(w1.w, xv[u], acc[7]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = acc[k];
      if (bias) v += __ldg(bias + c0 + k);
      v *= alpha;
      const size_t o = (size_t)(c0 + k) * T + t;
      if (res) v += res[(size_t)b * r_bs + o];
      out[(size_t)b * o_bs + o] = v;
    }
  }
}
cudaError_t pointwise_small(const float* x, long x_bs, const float* w, const float* bias, int C_in, int C_out, int B, int T,
                            float alpha, const float* res, long r_bs, float* out, long o_bs, cudaStream_t s) {
  if (C_out % 8 || (((uintptr_t)w) & 15)) return cudaErrorInvalidValue;
  dim3 grid(cdiv(T, 128), B);
  pointwise_small_kernel<<<grid, 128, 0, s>>>(x, x_bs, w, bias, C_in, C_out, T, alpha, res, r_bs, out, o_bs);
  return cudaGetLastError();
}

__global__ void copy_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
cudaError_t copy_f32(const float* in, float* out, size_t n, cudaStream_t s) {
  if (!n) return cudaSuccess;
  copy_kernel<<<cdiv(n, 256), 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

__global__ void fill_kernel(float* p, float v, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
cudaError_t fill_f32(float* p, float v, size_t n, cudaStream_t s) {
  if (!n) return cudaSuccess;
  fill_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, v, n);
  return cudaGetLastError();
}

}  // namespace dtts
