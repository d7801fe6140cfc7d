// Launchers of the non-convolution kernels of the Dict-TTS path (definitions in text_kernels.cu, lr_kernels.cu,
// repack.cu).  All activations inside the engine are fp32, channels-first [B, C, T] unless stated.
#pragma once
#include "common.cuh"
#include "tc_conv.cuh"

namespace dtts {

// ---- weight repack (create time) ----
// Conv1d weight [C_out][C_in][K] -> [C_in][K][C_out]; reverse_ci / reverse_co implement the folded glow Flip.
cudaError_t repack_conv(const float* w, float* out, int C_out, int C_in, int K, int reverse_ci, int reverse_co,
                        cudaStream_t s);
// ConvTranspose1d weight [C_in][C_out][K], stride S -> [S phases][C_in][K/S][C_out], Wp[ph][ci][m][co] = W[ci][co][m*S+ph]
cudaError_t repack_convT(const float* w, float* out, int C_in, int C_out, int K, int S, cudaStream_t s);
// Linear weight used transposed: in [R][C] -> out [C][R] (i.e. treat W^T as a 1x1 conv weight and pack it)
cudaError_t transpose2d(const float* in, float* out, int R, int C, cudaStream_t s);
cudaError_t reverse_vec(const float* in, float* out, int n, cudaStream_t s);
// stride-4 / pad-2 / k=8 Conv1d weight -> the stride-1 / pad-1 / k=3 weight over the 4x space-to-depth input
cudaError_t repack_s2d4(const float* w, float* out, int C_out, int C_in, cudaStream_t s);
cudaError_t fill_f32(float* p, float v, size_t n, cudaStream_t s);
// WaveNet in_layers / cond_layer rows ([groups][tanh half | sigmoid half][row_len]) re-ordered per N-row block as
// [N/2 tanh | the N/2 sigmoid rows that gate them] (TcConvParams::gate)
// prefetch.global.L2 over [p, p + bytes) (returns at once; the fills proceed in the background)
cudaError_t l2_prefetch(const void* p, size_t bytes, cudaStream_t s);
// C [M][N] = A B with A [M][K] (or [K][M] when transA) and B [K][N], row-major fp32
cudaError_t matmul_f32(const float* A, const float* B, float* C, int M, int N, int K, int transA, cudaStream_t s);
cudaError_t permute_gate_rows(const float* in, float* out, int groups, int hidden, int N, long row_len, cudaStream_t s);

// ---- text encoder ----
// x[b,c,t] = emb[tok[b,t]][c]*scale ; lens[b] = #(tok>0) ; seq_mask[b,t] = t < lens[b] ; tok_mask[b,t] = tok>0
cudaError_t embed_tokens(const int64_t* tok, const float* emb, float scale, int B, int Tw, int H, int vocab,
                         float* x, float* seq_mask, float* tok_mask, int* lens, cudaStream_t s);
// y = LN_c(x * in_mask) * gamma + beta, then * out_mask   (masks may be null) ; layout [B,C,T]
// relu: y = relu(LN(.)) * out_mask (ConvReluNorm, portaspeech/glow_modules.py:65-72)
cudaError_t channel_layernorm(const float* x, float* y, const float* gamma, const float* beta, float eps,
                              const float* in_mask, const float* out_mask, int B, int C, int T, cudaStream_t s,
                              int relu = 0);
// Tiled LayerNorm with optional extra outputs: xw = x * in_mask (may alias x), y [B,C,T] (may be null) and operand planes.
struct PlaneOut;
cudaError_t channel_layernorm_planes(const float* x, float* xw, float* y, const float* gamma, const float* beta, float eps,
                                     const float* in_mask, const float* out_mask, int B, int C, int T, const PlaneOut& po,
                                     cudaStream_t s);
// Shared-memory self attention for short sequences (self_attention_planes_fits: T <= 128 at d_k = 96); q,k,v are the three
// C-channel thirds of one [B,3C,T] tensor.
bool self_attention_planes_fits(int C, int T, int heads);
cudaError_t self_attention_planes(const float* q, const float* k, const float* v, const float* mask, float* out, int B,
                                  int C, int T, int heads, const PlaneOut& po, cudaStream_t s);
cudaError_t wn_gate_planes(const float* a, int B, int H, int T, const PlaneOut& po, cudaStream_t s);
// x *= mask (in place), [B,C,T] with mask [B,T]
cudaError_t apply_mask(float* x, const float* mask, int B, int C, int T, cudaStream_t s);
// multi-head self attention on q,k,v [B,C,T] (C = heads*dk), mask [B,T]; out [B,C,T]
cudaError_t self_attention(const float* q, const float* k, const float* v, const float* mask, float* out, int B,
                           int C, int T, int heads, cudaStream_t s);
// S2PA streaming pass. qk [B,D,Tw] (channels-first, already scaled). Writes weights [B,Tw,Lk], align [B,1,Lk,Tw],
// ctx [B,D,Tw] = sum_l w*values.  key_map float [B,Tw,Lk].
cudaError_t s2pa_stream(const float* keys, const float* values, const float* key_map, const float* qk, int B, int Tw,
                        int Lk, int D, float* weights, float* align, float* ctx, cudaStream_t s,
                        const int64_t* row_off = nullptr, const int32_t* row_len = nullptr);
// S2PA over projected rows (s2pa_route = 1): kv [B][2H][Tw*Lk] (k | v), q [B][H][Tw] scaled; same outputs as
// s2pa_stream with ctx [B,H,Tw] already in the projected space (only W_o remains).
cudaError_t s2pa_attend(const float* kv, const float* q, const float* key_map, int B, int Tw, int Lk, int H,
                        float* weights, float* align, float* ctx, cudaStream_t s);
// GPU-resident dictionary bank: per-batch key_map / pinyin / pinyin_map in the collater's padded
// layout plus the row window of each character in the keys / values bank.  *err != 0: an id or a length was out of range.
cudaError_t dict_bank_gather(const int64_t* ids, const int64_t* tok_off, const int64_t* pin_off,
                             const float* bank_key_map, const int64_t* bank_pinyin, const int64_t* bank_pinyin_map,
                             int n_entries, int B, int Tw, int Lk, int Lp, float* key_map, int64_t* pinyin,
                             int64_t* pinyin_map, int64_t* row_off, int32_t* row_len, int* err, cudaStream_t s);
// global maxima of key_map (float, as int) and pinyin_map (int64) -> maxes[0], maxes[1]
cudaError_t dict_maxes(const float* key_map, size_t n_key, const int64_t* pinyin_map, size_t n_pin, int* maxes,
                       cudaStream_t s);
// pronunciation mixing: pron_attn [B,Tw,Lp]; x2[b,h,t] = context[b,h,t]*seq_mask[b,t] + sum_p pron_w*pinyin_emb
cudaError_t s2pa_pron(const float* weights, const float* key_map, const int64_t* pinyin, const int64_t* pinyin_map,
                      const int64_t* pron_modified, const int* maxes, const float* pinyin_emb, int pinyin_vocab,
                      const float* context, const float* seq_mask, int B, int Tw, int Lk, int Lp, int H,
                      int apply_rule, float* pron_attn, float* x2, cudaStream_t s);
// word_encoder_out[b,t,h] = x[b,h,t]*tok_mask ; dur_in[b,h,t] same (channels-first) ; keep[b,t] = (sum_h |.| != 0)
cudaError_t finish_text(const float* x, const float* tok_mask, int B, int Tw, int H, float* enc_btc, float* dur_in,
                        float* keep, cudaStream_t s);
// ilens[b] = sum_t keep
cudaError_t count_keep(const float* keep, int B, int Tw, int64_t* ilens, cudaStream_t s);
// dur[b,t] = softplus(w . xs[b,:,t] + bias) * keep ; dur_int = clamp(rint(exp(dur)-1), 0)
cudaError_t dur_head(const float* xs, const float* w, const float* bias, const float* keep, int B, int C, int Tw,
                     float* dur, int64_t* dur_int, cudaStream_t s);

// ---- length regulator ----
// cum[b,w] inclusive prefix sum over the first ilens[b] durations (all-zero row -> durations 1); totals[b];
// atomicMax into *t_max (must be zeroed by the caller)
cudaError_t lr_scan(const int64_t* dur, const int64_t* ilens, int B, int Tw, int* cum, int* totals, int* t_max,
                    cudaStream_t s);
// mel2word[b,t] for t < T_raw from cum; columns [T_raw, T) repeat column T_raw-1
cudaError_t lr_fill(const int* cum, const int64_t* ilens, int B, int Tw, int T_raw, int T, int64_t* mel2word,
                    cudaStream_t s);
// gather: out_btc[b,t,:] = enc[b, m-1, :] (0 if m==0) ; out_bct = transpose ; nonpad[b,t] = m>0
cudaError_t lr_gather(const float* enc_btc, const int64_t* mel2word, int B, int Tw, int T, int H, float* out_btc,
                      float* out_bct, float* nonpad, cudaStream_t s);

// ---- PortaSpeech sibling (ps_kernels.cu) ----
// self attention of any length on q,k,v slices of one [B,3C,T] tensor, with the relative-position terms of
// rel_transformer_encoder.py:117-233 when rel_k / rel_v ([2*window+1, C/heads], shared by the heads) are given
cudaError_t rel_self_attention(const float* q, const float* k, const float* v, const float* mask, const float* rel_k,
                               const float* rel_v, int window, float* out, int B, int C, int T, int heads,
                               const PlaneOut& po, cudaStream_t s);
cudaError_t ps_finish_ph(const float* x, const int64_t* tok, int B, int Tp, int H, float* ph_btc, float* ph_bct,
                         float* keep, cudaStream_t s);
cudaError_t ps_group_by_segs(const float* ph, const int64_t* ph2word, int B, int Tp, int Tw, int H, float* out,
                             cudaStream_t s);
cudaError_t ps_fft_prepare(float* x, const float* table, int table_rows, const float* alpha, int B, int C, int T,
                           float* keep, cudaStream_t s);
cudaError_t ps_word_durations(const float* dur_ph, const float* keep_ph, const int64_t* ph2word, int B, int Tp, int Tw,
                              float* dur, int64_t* dur_int, int64_t* ilens, cudaStream_t s);
// out [B,2H,N] = [features ; sinusoidal in-word positions]; features = feat_btc[b,n,:] (gather == 0) or
// feat_btc[b, x2word[b,n]-1, :] (gather != 0, feat_btc is [B,Tw,H], zero row for x2word == 0)
cudaError_t ps_build_cat(const float* feat_btc, int gather, const int64_t* x2word, const float* freqs, int B, int N,
                         int Tw, int H, float* out, cudaStream_t s);
cudaError_t ps_word_attention(const float* q, const float* kv, const int64_t* mel2word, const int64_t* ph2word, int B,
                              int H, int T, int Tp, float* attn, float* ctx, cudaStream_t s);
cudaError_t bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t s);
cudaError_t nonpad_mask(const int64_t* idx, float* mask, size_t n, cudaStream_t s);

// ---- FVAE decoder pre_net ----
// ConvTranspose1d(C_in -> C_out, kernel = stride = 4) (fvae_semantics.py:40-41,53): out[b,co,4q+ph] = bias[co] +
// sum_ci w[ph][ci][co] * z[b,ci,q]  (w in repack_convT layout) -> fp32 [B,C_out,4*Tq] and operand planes (halo zeroed)
cudaError_t fvae_pre_net_planes(const float* z, const float* w, const float* bias, int B, int C_in, int C_out, int Tq,
                                float* out, const PlaneOut& po, cudaStream_t s);

// ---- WaveNet gate ----
// acts[b,c,t] = tanh(a[b,c,t]) * sigmoid(a[b,c+H,t]),  a [B,2H,T]
cudaError_t wn_gate(const float* a, float* acts, int B, int H, int T, cudaStream_t s);
cudaError_t copy_f32(const float* in, float* out, size_t n, cudaStream_t s);
// out[b,co,t] = alpha * (sum_ci w[ci][co] * x[b,ci,t] + bias[co]) + res[b,co,t]; x / out / res [.,C,T] with batch strides;
// w packed [C_in][C_out] (ConvW of a 1x1 convolution), C_out % 8 == 0
cudaError_t pointwise_small(const float* x, long x_bs, const float* w, const float* bias, int C_in, int C_out, int B, int T,
                            float alpha, const float* res, long r_bs, float* out, long o_bs, cudaStream_t s);

}  // namespace dtts
