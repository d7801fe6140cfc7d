// Kernels of the PortaSpeech (non-dict) sibling's text side: relative-position self attention,
// segment mean, FFT-block preparation, per-word durations, in-word positions and the word-to-phoneme attention.
//
// Reference semantics: modules/commons/rel_transformer_encoder.py:117-233 (MultiHeadAttention with window_size),
// modules/portaspeech/utils.py:3-16 (group_hidden_by_segs), modules/fastspeech/tts_modules.py:493-518 and
// modules/commons/common_layers.py:93-148 (FFTBlocks positions), modules/portaspeech/model.py:22-33,304-363
// (SinusoidalPosEmb, attention, add_dur, build_pos_embed / build_word_mask).
#include "kernels.cuh"
#include "tc16.cuh"

namespace dtts {

// ------------------------------------------------------------------------------------------------------------
// Multi-head self attention with optional relative-position terms, any sequence length.  One warp per query row,
// block = 8 rows of one (batch item, head).  q, k, v are channel-first slices of one [B, 3C, T] tensor.
//   score(t, s) = q_t . k_s / sqrt(dk)  +  [|s - t| <= w] q_t . emb_rel_k[s - t + w] / sqrt(dk)
//   masked_fill(mask_t * mask_s == 0, -1e4); softmax over s
//   out_t = sum_s p(t, s) v_s  +  sum_{|s - t| <= w} p(t, s) emb_rel_v[s - t + w]
// (the reference reaches the same terms through its pad-and-reshape relative <-> absolute index trick,
// rel_transformer_encoder.py:160-233; the embeddings are shared by the heads, :108-111).
__global__ void __launch_bounds__(256) rel_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                             const float* __restrict__ v,
                                                             const float* __restrict__ mask,
                                                             const float* __restrict__ rel_k,
                                                             const float* __restrict__ rel_v, int window,
                                                             float* __restrict__ out, int C, int T, int heads, long bs,
                                                             PlaneOut po) {
  extern __shared__ float sm[];                 // per warp: q row [dk], probabilities [T], output row [dk]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dk = C / heads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int t = blockIdx.y * nw + warp;
  float* qs = sm + (size_t)warp * (2 * dk + T);
  float* pr = qs + dk;
  float* os = pr + T;
  const float* qb = q + (size_t)b * bs + (size_t)h * dk * T;
  const float* kb = k + (size_t)b * bs + (size_t)h * dk * T;
  const float* vb = v + (size_t)b * bs + (size_t)h * dk * T;
  const float* mb = mask + (size_t)b * T;
  const float inv = rsqrtf((float)dk);
  if (t < T) {
    for (int d = lane; d < dk; d += 32) qs[d] = qb[(size_t)d * T + t];
    __syncwarp();
    const float mt = mb[t];
    float mx = -INFINITY;
    for (int s = lane; s < T; s += 32) {
      float acc = 0.f;
      for (int d = 0; d < dk; ++d) acc = fmaf(qs[d], kb[(size_t)d * T + s], acc);
      acc *= inv;
      const int r = s - t + window;
      if (rel_k && r >= 0 && r <= 2 * window) {
        const float* e = rel_k + (size_t)r * dk;
        float a2 = 0.f;
        for (int d = 0; d < dk; ++d) a2 = fmaf(qs[d], e[d], a2);
        acc += a2 * inv;
      }
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
    for (int s = lane; s < T; s += 32) pr[s] *= rs;
    __syncwarp();
    const int s_lo = max(0, t - window), s_hi = min(T - 1, t + window);
    for (int d = 0; d < dk; ++d) {                       // lanes along s: coalesced reads of v, one reduction per channel
      float acc = 0.f;
      for (int s = lane; s < T; s += 32) acc = fmaf(pr[s], vb[(size_t)d * T + s], acc);
      if (rel_v) {
        const int s = s_lo + lane;
        if (s <= s_hi) acc = fmaf(pr[s], rel_v[(size_t)(s - t + window) * dk + d], acc);   // 2w + 1 <= 32 (launcher)
      }
      acc = warp_sum(acc);
      if (lane == 0) {
        os[d] = acc;
        if (out) out[(size_t)b * C * T + (size_t)(h * dk + d) * T + t] = acc;
      }
    }
  }
  if (po.hi) {
    __syncwarp();
    if (t < T) {
      for (int sl = lane; sl < dk / 8; sl += 32) {
        float v8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v8[e] = os[sl * 8 + e];
        store_slab(po, b, C, h * (dk / 8) + sl, t, v8);
      }
    }
  }
}

// The same attention for T <= 64 with q, k, v of one (batch item, head) -- and the relative-position tables -- resident in
// shared memory, one block per (b, head): the per-row kernel above re-reads K and V from L2 for every query row and pays a
// warp reduction per output channel (255 us at [60, 2 heads, T = 64]; this one: one staging pass + T*T*dk FMAs from shared
// memory).  Same formulas, scores summed over d in the same order.
__global__ void __launch_bounds__(256) rel_self_attn_small_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                                   const float* __restrict__ v,
                                                                   const float* __restrict__ mask,
                                                                   const float* __restrict__ rel_k,
                                                                   const float* __restrict__ rel_v, int window,
                                                                   float* __restrict__ out, int C, int T, int heads,
                                                                   long bs, PlaneOut po) {
  extern __shared__ float sm[];   // qs[dk][T], ks[dk][T], vs[dk][T], ps[T][T+1], rk[dk][2w+1] (transposed), rv[2w+1][dk]
  griddep_launch_if_resident();
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int dk = C / heads, nrel = 2 * window + 1;
  float* qs = sm;
  float* ks = qs + (size_t)dk * T;
  float* vs = ks + (size_t)dk * T;
  float* ps = vs + (size_t)dk * T;
  float* rk = ps + (size_t)T * (T + 1);
  float* rv = rk + (size_t)dk * nrel;
  const size_t base = (size_t)b * bs + (size_t)h * dk * T;
#pragma unroll 4
  for (int i = threadIdx.x; i < dk * T; i += blockDim.x) {
    const float a = __ldg(q + base + i), c = __ldg(k + base + i), d = __ldg(v + base + i);
    qs[i] = a;
    ks[i] = c;
    vs[i] = d;
  }
  if (rel_k) {
    for (int i = threadIdx.x; i < nrel * dk; i += blockDim.x) {
      const int r = i / dk, d = i - r * dk;
      rk[d * nrel + r] = __ldg(rel_k + i);
      rv[i] = __ldg(rel_v + i);
    }
  }
  __syncthreads();
  const float* mb = mask + (size_t)b * T;
  const float inv = rsqrtf((float)dk);
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    const int t = i / T, s2 = i - t * T;
    float acc = 0.f;
    for (int d = 0; d < dk; ++d) acc = fmaf(qs[d * T + t], ks[d * T + s2], acc);
    acc *= inv;
    const int r = s2 - t + window;
    if (rel_k && r >= 0 && r < nrel) {
      float a2 = 0.f;
      for (int d = 0; d < dk; ++d) a2 = fmaf(qs[d * T + t], rk[d * nrel + r], a2);
      acc += a2 * inv;
    }
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
  // out[d, t] = sum_s p[t, s] v[d, s] + sum_{|s - t| <= w} p[t, s] rel_v[s - t + w][d]; results overwrite qs (q is dead)
  for (int i = threadIdx.x; i < dk * T; i += blockDim.x) {
    const int d = i / T, t = i - d * T;
    const float* pr = ps + t * (T + 1);
    float acc = 0.f;
    for (int s2 = 0; s2 < T; ++s2) acc = fmaf(pr[s2], vs[d * T + s2], acc);
    if (rel_v) {
      for (int j = 0; j < nrel; ++j) {
        const int s2 = t - window + j;
        if (s2 >= 0 && s2 < T) acc = fmaf(pr[s2], rv[j * dk + d], acc);
      }
    }
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

cudaError_t rel_self_attention(const float* q, const float* k, const float* v, const float* mask, const float* rel_k,
                               const float* rel_v, int window, float* out, int B, int C, int T, int heads,
                               const PlaneOut& po, cudaStream_t s) {
  const int dk = C / heads;
  if (C % heads || (po.hi && dk % 8) || (rel_k && 2 * window + 1 > 32) || (!rel_k != !rel_v)) return cudaErrorInvalidValue;
  if (T <= 64) {
    const int nrel = rel_k ? 2 * window + 1 : 0;
    const size_t small = ((size_t)3 * dk * T + (size_t)T * (T + 1) + (size_t)2 * nrel * dk) * sizeof(float);
    if (small <= 160 * 1024) {
      static unsigned long long attr_done = 0;
      {
        const cudaError_t e = ensure_max_dyn_smem(rel_self_attn_small_kernel, 160 * 1024, &attr_done);
        if (e != cudaSuccess) return e;
      }
      rel_self_attn_small_kernel<<<B * heads, 256, small, s>>>(q, k, v, mask, rel_k, rel_v, rel_k ? window : 0, out, C, T, heads,
                                                               (long)3 * C * T, po);
      return cudaGetLastError();
    }
  }
  const int nw = 8;
  const size_t smem = (size_t)nw * (2 * dk + T) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rel_self_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(B * heads, cdiv(T, nw));
  rel_self_attn_kernel<<<grid, nw * 32, smem, s>>>(q, k, v, mask, rel_k, rel_v, window, out, C, T, heads, (long)3 * C * T,
                                                   po);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// ph_out[b,t,:] = x[b,:,t] * (tok[b,t] > 0)  ([B,Tp,H], ret['ph_encoder_out']); keep[b,t] = (sum_h |.| != 0), the
// src_padding of add_dur (model.py:325); ilens[b] = sum_t keep.  One block per (b, t) + a tail block per b.
__global__ void ps_finish_ph_kernel(const float* __restrict__ x, const int64_t* __restrict__ tok, int Tp, int H,
                                    float* __restrict__ ph_btc, float* __restrict__ ph_bct, float* __restrict__ keep) {
  const int bt = blockIdx.x;
  const int b = bt / Tp, t = bt - b * Tp;
  const float mk = tok[bt] > 0 ? 1.f : 0.f;
  __shared__ float s_red[8];
  float a = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const size_t i = ((size_t)b * H + h) * Tp + t;
    const float v = x[i] * mk;
    ph_btc[(size_t)bt * H + h] = v;
    ph_bct[i] = v;
    a += fabsf(v);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += s_red[i];
    keep[bt] = tot == 0.f ? 0.f : 1.f;
  }
}
cudaError_t ps_finish_ph(const float* x, const int64_t* tok, int B, int Tp, int H, float* ph_btc, float* ph_bct,
                         float* keep, cudaStream_t s) {
  ps_finish_ph_kernel<<<B * Tp, 64, 0, s>>>(x, tok, Tp, H, ph_btc, ph_bct, keep);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// group_hidden_by_segs (portaspeech/utils.py:3-16): word[b, :, w] = mean of the phoneme vectors with ph2word == w + 1
// (summed in phoneme order like scatter_add; a word without phonemes is a zero row).  ph [B,Tp,H] -> out [B,H,Tw].
__global__ void ps_group_by_segs_kernel(const float* __restrict__ ph, const int64_t* __restrict__ ph2word, int Tp, int Tw,
                                        int H, float* __restrict__ out) {
  const int w = blockIdx.x, b = blockIdx.y;
  const int64_t* seg = ph2word + (size_t)b * Tp;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float sum = 0.f, cnt = 0.f;
    for (int t = 0; t < Tp; ++t) {
      if (seg[t] == w + 1) {
        sum += ph[((size_t)b * Tp + t) * H + h];
        cnt += 1.f;
      }
    }
    out[((size_t)b * H + h) * Tw + w] = sum / fmaxf(cnt, 1.f);
  }
}
cudaError_t ps_group_by_segs(const float* ph, const int64_t* ph2word, int B, int Tp, int Tw, int H, float* out,
                             cudaStream_t s) {
  dim3 grid(Tw, B);
  ps_group_by_segs_kernel<<<grid, 64, 0, s>>>(ph, ph2word, Tp, Tw, H, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// FFTBlocks.forward prologue (tts_modules.py:493-507): keep = 1 - (x.abs().sum(-1) == 0); positions =
// make_positions(x[..., 0], padding_idx 0) = cumsum(x0 != 0) * (x0 != 0); x = (x + alpha * sin_table[positions]) * keep.
// x [B,C,T] in place; one block per utterance.
__global__ void ps_fft_prepare_kernel(float* __restrict__ x, const float* __restrict__ table, int table_rows,
                                      const float* __restrict__ alpha, int C, int T, float* __restrict__ keep) {
  extern __shared__ int s_pos[];                // [T] positions, then [T] keep flags
  int* s_keep = s_pos + T;
  const int b = blockIdx.x;
  float* xb = x + (size_t)b * C * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a += fabsf(xb[(size_t)c * T + t]);
    s_keep[t] = a != 0.f;
    keep[(size_t)b * T + t] = a != 0.f ? 1.f : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < T; ++t) {
      const int nz = xb[t] != 0.f;                    // channel 0
      run += nz;
      int p = nz ? run : 0;
      s_pos[t] = p < table_rows ? p : table_rows - 1;
    }
  }
  __syncthreads();
  const float al = alpha[0];
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const int c = i / T, t = i - c * T;
    const float v = xb[i] + al * table[(size_t)s_pos[t] * C + c];
    xb[i] = s_keep[t] ? v : 0.f;
  }
}
cudaError_t ps_fft_prepare(float* x, const float* table, int table_rows, const float* alpha, int B, int C, int T,
                           float* keep, cudaStream_t s) {
  ps_fft_prepare_kernel<<<B, 256, 2 * T * sizeof(int), s>>>(x, table, table_rows, alpha, C, T, keep);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// add_dur at word level (model.py:328-339): dur[b,w] = sum of the phoneme-level predictions of word w + 1 (scatter_add
// order = phoneme order), dur_int = clamp(round(exp(dur) - 1), 0); ilens[b] = number of non-padding PHONEMES.
__global__ void ps_word_durations_kernel(const float* __restrict__ dur_ph, const float* __restrict__ keep_ph,
                                         const int64_t* __restrict__ ph2word, int Tp, int Tw, float* __restrict__ dur,
                                         int64_t* __restrict__ dur_int, int64_t* __restrict__ ilens) {
  const int b = blockIdx.x;
  const int64_t* seg = ph2word + (size_t)b * Tp;
  for (int w = threadIdx.x; w < Tw; w += blockDim.x) {
    float sum = 0.f;
    for (int t = 0; t < Tp; ++t)
      if (seg[t] == w + 1) sum += dur_ph[(size_t)b * Tp + t];
    dur[(size_t)b * Tw + w] = sum;
    dur_int[(size_t)b * Tw + w] = (int64_t)fmaxf(rintf(expf(sum) - 1.f), 0.f);
  }
  if (threadIdx.x == 0) {
    float c = 0.f;
    for (int t = 0; t < Tp; ++t) c += keep_ph[(size_t)b * Tp + t];
    ilens[b] = (int64_t)(c + 0.5f);
  }
}
cudaError_t ps_word_durations(const float* dur_ph, const float* keep_ph, const int64_t* ph2word, int B, int Tp, int Tw,
                              float* dur, int64_t* dur_int, int64_t* ilens, cudaStream_t s) {
  ps_word_durations_kernel<<<B, 128, 0, s>>>(dur_ph, keep_ph, ph2word, Tp, Tw, dur, dur_int, ilens);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Concatenated attention inputs, channel-first [B, 2H, N] (model.py:279-283,359-363):
//   channels [0, H)  : feat[b, n, :]            (gather == 0: the phoneme encoder output [B,N,H])
//                      or feat[b, x2word - 1, :] (gather != 0: word encoder output [B,Tw,H] by mel2word, zero row for 0)
//   channels [H, 2H) : SinusoidalPosEmb(pos)[c], pos = (rank of n among the positions of its word) / (size of the word),
//                      0 for x2word == 0 or a word id above Tw; emb = [sin(pos * e_i) | cos(pos * e_i)], i < H/2
// x2word need not be sorted: rank and size are counted over the whole row, as the reference's cumsum over the
// [B, T_word, N] mask does.
__global__ void ps_build_cat_kernel(const float* __restrict__ feat_btc, int gather,
                                    const int64_t* __restrict__ x2word, const float* __restrict__ freqs, int N, int Tw,
                                    int H, float* __restrict__ out) {
  extern __shared__ float s_posv[];             // [N]
  const int b = blockIdx.x;
  const int64_t* seg = x2word + (size_t)b * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const int64_t w = seg[n];
    float pos = 0.f;
    if (w >= 1 && w <= Tw) {
      int rank = 0, cnt = 0;
      for (int m = 0; m < N; ++m) {
        const int same = seg[m] == w;
        cnt += same;
        rank += same && m <= n;
      }
      pos = (float)rank / (float)(cnt > 0 ? cnt : 1);
    }
    s_posv[n] = pos;
  }
  __syncthreads();
  const int half = H / 2;
  float* ob = out + (size_t)b * 2 * H * N;
  for (int i = threadIdx.x; i < H * N; i += blockDim.x) {
    const int c = i / N, n = i - c * N;
    float f;
    if (!gather) {
      f = feat_btc[((size_t)b * N + n) * H + c];
    } else {
      const int64_t w = seg[n];
      f = (w >= 1 && w <= Tw) ? feat_btc[((size_t)b * Tw + (w - 1)) * H + c] : 0.f;
    }
    ob[i] = f;
    const float a = s_posv[n] * freqs[c < half ? c : c - half];
    ob[(size_t)H * N + i] = c < half ? sinf(a) : cosf(a);
  }
}
cudaError_t ps_build_cat(const float* feat_btc, int gather, const int64_t* x2word, const float* freqs, int B, int N,
                         int Tw, int H, float* out, cudaStream_t s) {
  if (H % 2 || !feat_btc) return cudaErrorInvalidValue;
  const size_t smem = (size_t)N * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ps_build_cat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ps_build_cat_kernel<<<B, 256, smem, s>>>(feat_btc, gather, x2word, freqs, N, Tw, H, out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Word-to-phoneme attention (model.py:304-315; one head, no biases): q [B,H,T] already scaled by H^-1/2, kv [B,2H,Tp]
// (k | v).  scores[t,s] = q_t . k_s + (mel2word[t] == ph2word[s] ? 0 : -1e9), softmax over s, ctx_t = sum_s w v_s.
// One warp per frame.  attn (optional) [B,T,Tp]; ctx [B,H,T].
__global__ void __launch_bounds__(256) ps_word_attn_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                            const int64_t* __restrict__ mel2word,
                                                            const int64_t* __restrict__ ph2word, int H, int T, int Tp,
                                                            float* __restrict__ attn, float* __restrict__ ctx) {
  extern __shared__ float sm[];                 // per warp: q [H], p [Tp]
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int t = blockIdx.x * nw + warp;
  if (t >= T) return;
  float* qs = sm + (size_t)warp * (H + Tp);
  float* pr = qs + H;
  const float* kb = kv + (size_t)b * 2 * H * Tp;
  const float* vb = kb + (size_t)H * Tp;
  for (int c = lane; c < H; c += 32) qs[c] = q[((size_t)b * H + c) * T + t];
  __syncwarp();
  const int64_t wt = mel2word[(size_t)b * T + t];
  const int64_t* ps = ph2word + (size_t)b * Tp;
  float mx = -INFINITY;
  for (int s = lane; s < Tp; s += 32) {
    float acc = 0.f;
    for (int c = 0; c < H; ++c) acc = fmaf(qs[c], kb[(size_t)c * Tp + s], acc);
    acc += (ps[s] == wt) ? 0.f : -1e9f;
    pr[s] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int s = lane; s < Tp; s += 32) {
    const float e = expf(pr[s] - mx);
    pr[s] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float rs = 1.f / sum;
  for (int s = lane; s < Tp; s += 32) {
    const float w = pr[s] * rs;
    pr[s] = w;
    if (attn) attn[((size_t)b * T + t) * Tp + s] = w;
  }
  __syncwarp();
  for (int c = 0; c < H; ++c) {
    float acc = 0.f;
    for (int s = lane; s < Tp; s += 32) acc = fmaf(pr[s], vb[(size_t)c * Tp + s], acc);
    acc = warp_sum(acc);
    if (lane == 0) ctx[((size_t)b * H + c) * T + t] = acc;
  }
}
// The same attention with K and V of the utterance and a tile of 64 query frames resident in shared memory (block = one
// utterance x 64 frames, 8 warps x 8 frames): the per-frame kernel above re-reads K | V (2 * H * Tp floats) from L2 for every
// frame -- 2.3 GB for a [60, 192, 400] x [.., 64] call, 795 us.  Scores are summed over c in the same order (attn is
// bit-identical); ctx sums over s sequentially with lanes along the channels instead of a warp reduction per channel.
__global__ void __launch_bounds__(256) ps_word_attn_tile_kernel(const float* __restrict__ q, const float* __restrict__ kv,
                                                                 const int64_t* __restrict__ mel2word,
                                                                 const int64_t* __restrict__ ph2word, int H, int T, int Tp,
                                                                 float* __restrict__ attn, float* __restrict__ ctx) {
  extern __shared__ float sm[];   // ks[H][Tp], vs[H][Tp+1], qt[H][64] (becomes the ctx tile), pr[8][Tp], seg[Tp]
  constexpr int FPB = 64;
  const int b = blockIdx.y, t0 = blockIdx.x * FPB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* ks = sm;
  float* vs = ks + (size_t)H * Tp;
  float* qt = vs + (size_t)H * (Tp + 1);
  float* prs = qt + (size_t)H * FPB;
  int* seg = reinterpret_cast<int*>(prs + (size_t)nw * Tp);
  const float* kb = kv + (size_t)b * 2 * H * Tp;
  const float* vb = kb + (size_t)H * Tp;
#pragma unroll 4
  for (int i = threadIdx.x; i < H * Tp; i += blockDim.x) {
    const float a = __ldg(kb + i), c = __ldg(vb + i);
    ks[i] = a;
    vs[(i / Tp) * (Tp + 1) + (i % Tp)] = c;
  }
  const int nf = min(FPB, T - t0);
#pragma unroll 4
  for (int i = threadIdx.x; i < H * FPB; i += blockDim.x) {
    const int c = i / FPB, f = i - c * FPB;
    qt[i] = f < nf ? __ldg(q + ((size_t)b * H + c) * T + t0 + f) : 0.f;
  }
  for (int i = threadIdx.x; i < Tp; i += blockDim.x) seg[i] = (int)ph2word[(size_t)b * Tp + i];
  __syncthreads();
  float* pr = prs + (size_t)warp * Tp;
  for (int f = warp; f < nf; f += nw) {
    const int t = t0 + f;
    const int wt = (int)mel2word[(size_t)b * T + t];
    float mx = -INFINITY;
    for (int s = lane; s < Tp; s += 32) {
      float acc = 0.f;
      for (int c = 0; c < H; ++c) acc = fmaf(qt[c * FPB + f], ks[c * Tp + s], acc);
      acc += (seg[s] == wt) ? 0.f : -1e9f;
      pr[s] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < Tp; s += 32) {
      const float e = expf(pr[s] - mx);
      pr[s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float rs = 1.f / sum;
    for (int s = lane; s < Tp; s += 32) {
      const float w = pr[s] * rs;
      pr[s] = w;
      if (attn) attn[((size_t)b * T + t) * Tp + s] = w;
    }
    __syncwarp();
    for (int c = lane; c < H; c += 32) {               // lanes along the channels; column f of the q tile is dead by now
      const float* vr = vs + (size_t)c * (Tp + 1);
      float acc = 0.f;
      for (int s = 0; s < Tp; ++s) acc = fmaf(pr[s], vr[s], acc);
      qt[c * FPB + f] = acc;
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * FPB; i += blockDim.x) {
    const int c = i / FPB, f = i - c * FPB;
    if (f < nf) ctx[((size_t)b * H + c) * T + t0 + f] = qt[i];
  }
}

cudaError_t ps_word_attention(const float* q, const float* kv, const int64_t* mel2word, const int64_t* ph2word, int B,
                              int H, int T, int Tp, float* attn, float* ctx, cudaStream_t s) {
  const int nw = 8;
  const size_t tile = ((size_t)H * Tp + (size_t)H * (Tp + 1) + (size_t)H * 64 + (size_t)nw * Tp + Tp) * sizeof(float);
  if (tile <= 200 * 1024) {
    static unsigned long long attr_done = 0;
    {
      const cudaError_t e = ensure_max_dyn_smem(ps_word_attn_tile_kernel, 200 * 1024, &attr_done);
      if (e != cudaSuccess) return e;
    }
    dim3 grid(cdiv(T, 64), B);
    ps_word_attn_tile_kernel<<<grid, nw * 32, tile, s>>>(q, kv, mel2word, ph2word, H, T, Tp, attn, ctx);
    return cudaGetLastError();
  }
  const size_t smem = (size_t)nw * (H + Tp) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ps_word_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid(cdiv(T, nw), B);
  ps_word_attn_kernel<<<grid, nw * 32, smem, s>>>(q, kv, mel2word, ph2word, H, T, Tp, attn, ctx);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// [B,C,T] -> [B,T,C] (tiled through shared memory) and x_mask[b,t] = mel2word > 0.
__global__ void bct_to_btc_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, t = t0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && t < T) ? in[((size_t)b * C + c) * T + t] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, c = c0 + threadIdx.x;
    if (t < T && c < C) out[((size_t)b * T + t) * C + c] = tile[threadIdx.x][r];
  }
}
cudaError_t bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t s) {
  dim3 grid(cdiv(T, 32), cdiv(C, 32), B), block(32, 8);
  bct_to_btc_kernel<<<grid, block, 0, s>>>(in, out, C, T);
  return cudaGetLastError();
}

__global__ void nonpad_mask_kernel(const int64_t* __restrict__ idx, float* __restrict__ mask, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = idx[i] > 0 ? 1.f : 0.f;
}
cudaError_t nonpad_mask(const int64_t* idx, float* mask, size_t n, cudaStream_t s) {
  nonpad_mask_kernel<<<cdiv(n, 256), 256, 0, s>>>(idx, mask, n);
  return cudaGetLastError();
}

}  // namespace dtts
