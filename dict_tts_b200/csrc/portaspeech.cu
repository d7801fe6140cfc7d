// PortaSpeech (non-dict) sibling behind the same handle type: everything PortaSpeech.run_text_encoder
// computes at dur_level = word (modules/portaspeech/model.py:239-262) -- phoneme encoder with relative-position
// attention, segment mean, FFT-block word encoder, phoneme-level durations summed per word, in-word positions and the
// word-to-phoneme attention.  The length regulator, the duration predictor and the FVAE decoder are the dict model's
// (acoustic.cu); dense layers run on tc_conv_kernel (precision 1) or conv1d_f32_kernel (precision 0) like there.
#include <cmath>

#include "acoustic.cuh"

using namespace dtts;
using namespace dtts::ac;

namespace dtts {
namespace ac {

struct DenseW {                     // one dense layer in both packings
  ConvW f;
  TcConvW t;
};
struct FftLayerW {                  // EncSALayer (commons/common_layers.py:624-673), kernel size 1
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  DenseW qkv, o, ffn1, ffn2;
};
struct PsW {
  const float* ph_emb = nullptr;
  static constexpr int kPreLayers = 3, kPreKernel = 5;       // ConvReluNorm(..., kernel_size=5, n_layers=3) (model.py:103-105)
  DenseW pre_conv[kPreLayers], pre_proj;
  const float *pre_g[kPreLayers], *pre_b[kPreLayers];
  EncoderW enc;                                                // post-LN Encoder (pre_ln = False: no last_ln)
  std::vector<const float*> rel_k, rel_v;                      // [2w+1][dk] per layer
  std::vector<FftLayerW> word;
  const float *word_ln_g = nullptr, *word_ln_b = nullptr, *pos_alpha = nullptr;
  int word_filter = 0;
  float* sin_table = nullptr;                                  // [sin_rows][H], SinusoidalPositionalEmbedding
  int sin_rows = 0;
  float* freqs = nullptr;                                      // [H/2]: exp(i * -(ln 10000 / (H/2 - 1)))
  DenseW enc_pos_proj, dec_query_proj, dec_res_proj, attn_in, attn_out;
};

namespace {

int pack_dense(dtts_acoustic* h, const std::string& name, bool has_bias, int C_out, int C_in, int K, DenseW* w,
               cudaStream_t s, int N = 0, const char* wsuffix = ".weight") {
  DTTS_TRY(pack(h, name, C_out, C_in, K, has_bias, &w->f, s, 0, 0, wsuffix));
  if (!h->precision) return DTTS_OK;
  const float* wp = h->tab.get(name + wsuffix, (uint64_t)C_out * C_in * K);
  if (!wp) return DTTS_ERR_MISSING_WEIGHT;
  return tc_pack(h, &wp, 1, w->f.bias, C_out, C_in, K, 0, N, &w->t, s);
}

// One dense layer on either pipe: y[b,co,t] = post * (act(conv(x) + bias) * alpha * mask + res) over output channels
// [co_off, co_off + co_n) of w.  x, y, res: fp32 [B,C,T] channel-first.
struct PsRun {
  dtts_acoustic* h;
  Launcher* L;
  TcRun* tc;
  int B;
  void dense(const float* x, int T, const DenseW& w, float* y, int pad, const TcRun::Epi& e, int co_off = 0, int co_n = 0) {
    if (co_n <= 0) co_n = w.f.C_out - co_off;
    if (tc) {
      Planes& P = tc->P[0];
      tc->stage_nct(P, x, w.f.C_in, T);
      if (co_off % w.t.N || co_n % w.t.N) { (*L)(cudaErrorInvalidValue); return; }
      tc->conv_nct(P, w.t, y, T, 1, pad, e, co_off / w.t.N, co_n / w.t.N);
      return;
    }
    ConvParams p = conv_params(x, T, w.f, co_off, co_n, y, T, 1, 1, pad);
    p.act = e.act; p.alpha = e.alpha; p.post = e.post; p.accumulate = e.accumulate;
    p.mask = e.mask; p.m_bs = e.m_bs;
    p.res = e.res; p.r_bs = e.r_bs; p.r_cs = (int)e.r_cs; p.r_ts = (int)e.r_ts;
    (*L)(launch_conv1d_f32(p, B, L->stream));
  }
};

TcRun::Epi epi_res(const float* res, int C, int T) {
  TcRun::Epi e;
  e.res = res; e.r_bs = (long)C * T; e.r_cs = T; e.r_ts = 1;
  return e;
}

}  // namespace

int destroy_ps(dtts_acoustic* h) {
  delete h->ps;
  h->ps = nullptr;
  return DTTS_OK;
}

int create_ps(dtts_acoustic* h, cudaStream_t s) {
  const dtts_acoustic_desc& d = h->d;
  const int H = d.hidden, F = d.ffn_filter, dk = H / d.n_heads;
  PsW* P = h->ps = new PsW();
  P->ph_emb = h->tab.get("ph_encoder.emb.weight", (uint64_t)d.ph_size * H);
  if (!P->ph_emb) return DTTS_ERR_MISSING_WEIGHT;
  for (int i = 0; i < PsW::kPreLayers; ++i) {
    const std::string q = "ph_encoder.pre";
    DTTS_TRY(pack_dense(h, q + ".conv_layers." + std::to_string(i), true, H, H, PsW::kPreKernel, &P->pre_conv[i], s));
    P->pre_g[i] = h->tab.get(q + ".norm_layers." + std::to_string(i) + ".gamma", H);
    P->pre_b[i] = h->tab.get(q + ".norm_layers." + std::to_string(i) + ".beta", H);
    if (!P->pre_g[i] || !P->pre_b[i]) return DTTS_ERR_MISSING_WEIGHT;
  }
  DTTS_TRY(pack_dense(h, "ph_encoder.pre.proj", true, H, H, 1, &P->pre_proj, s));
  DTTS_TRY(pack_encoder(h, "ph_encoder.encoder", &P->enc, s));
  for (int i = 0; i < d.enc_layers; ++i) {
    const std::string a = "ph_encoder.encoder.attn_layers." + std::to_string(i);
    const float *rk = nullptr, *rv = nullptr;
    if (d.rel_window > 0) {
      rk = h->tab.get(a + ".emb_rel_k", (uint64_t)(2 * d.rel_window + 1) * dk);
      rv = h->tab.get(a + ".emb_rel_v", (uint64_t)(2 * d.rel_window + 1) * dk);
      if (!rk || !rv) return DTTS_ERR_MISSING_WEIGHT;
    }
    P->rel_k.push_back(rk);
    P->rel_v.push_back(rv);
  }
  // FFT-block word encoder: FastspeechDecoder(hidden, word_enc_layers, kernel 1, num_heads) (model.py:151-152)
  for (int i = 0; i < d.word_enc_layers; ++i) {
    const std::string q = "word_encoder.layers." + std::to_string(i) + ".op";
    FftLayerW Lw;
    Lw.ln1_g = h->tab.get(q + ".layer_norm1.weight", H); Lw.ln1_b = h->tab.get(q + ".layer_norm1.bias", H);
    Lw.ln2_g = h->tab.get(q + ".layer_norm2.weight", H); Lw.ln2_b = h->tab.get(q + ".layer_norm2.bias", H);
    if (!Lw.ln1_g || !Lw.ln1_b || !Lw.ln2_g || !Lw.ln2_b) return DTTS_ERR_MISSING_WEIGHT;
    auto it = h->tab.entries.find(q + ".ffn.ffn_1.bias");
    if (it == h->tab.entries.end()) return fail(DTTS_ERR_MISSING_WEIGHT, "missing weight: " + q + ".ffn.ffn_1.bias");
    const int Fw = (int)it->second.second;
    if (P->word_filter && P->word_filter != Fw) return fail(DTTS_ERR_BAD_SHAPE, "word encoder layers differ in FFN width");
    P->word_filter = Fw;
    DTTS_TRY(pack_dense(h, q + ".self_attn.in_proj_weight", false, 3 * H, H, 1, &Lw.qkv, s, H, ""));   // N = H: q | k | v blocks
    DTTS_TRY(pack_dense(h, q + ".self_attn.out_proj", false, H, H, 1, &Lw.o, s));
    DTTS_TRY(pack_dense(h, q + ".ffn.ffn_1", true, Fw, H, 1, &Lw.ffn1, s));     // kernel size 1: x k^-1/2 is x 1
    DTTS_TRY(pack_dense(h, q + ".ffn.ffn_2", true, H, Fw, 1, &Lw.ffn2, s));
    P->word.push_back(Lw);
  }
  P->word_ln_g = h->tab.get("word_encoder.layer_norm.weight", H);
  P->word_ln_b = h->tab.get("word_encoder.layer_norm.bias", H);
  P->pos_alpha = h->tab.get("word_encoder.pos_embed_alpha", 1);
  if (!P->word_ln_g || !P->word_ln_b || !P->pos_alpha) return DTTS_ERR_MISSING_WEIGHT;
  (void)F;
  // sinusoidal tables, computed once on the host in fp32 like the reference (common_layers.py:110-127, model.py:22-33):
  // freq_i = exp(i * -(ln 10000 / (half - 1))); table[p] = [sin(p * freq) | cos(p * freq)], row 0 (padding_idx) zero
  {
    const int half = H / 2, rows = 2048 + 1;
    std::vector<float> fr(half), tab((size_t)rows * H, 0.f);
    const float step = (float)(-(std::log(10000.0) / (double)(half - 1)));
    for (int i = 0; i < half; ++i) fr[i] = std::exp((float)i * step);
    for (int p = 1; p < rows; ++p)
      for (int i = 0; i < half; ++i) {
        const float a = (float)p * fr[i];
        tab[(size_t)p * H + i] = std::sin(a);
        tab[(size_t)p * H + half + i] = std::cos(a);
      }
    P->sin_table = h->pool.take(tab.size());
    P->freqs = h->pool.take(half);
    if (!P->sin_table || !P->freqs) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    DTTS_CUDA(cudaMemcpyAsync(P->sin_table, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    DTTS_CUDA(cudaMemcpyAsync(P->freqs, fr.data(), half * sizeof(float), cudaMemcpyHostToDevice, s));
    DTTS_CUDA(cudaStreamSynchronize(s));               // the host vectors die with this scope
    P->sin_rows = rows;
  }
  DTTS_TRY(pack_dense(h, "enc_pos_proj", true, H, 2 * H, 1, &P->enc_pos_proj, s));
  DTTS_TRY(pack_dense(h, "dec_query_proj", true, H, 2 * H, 1, &P->dec_query_proj, s));
  DTTS_TRY(pack_dense(h, "dec_res_proj", true, H, 2 * H, 1, &P->dec_res_proj, s));
  DTTS_TRY(pack_dense(h, "attn.in_proj_weight", false, 3 * H, H, 1, &P->attn_in, s, H, ""));
  DTTS_TRY(pack_dense(h, "attn.out_proj", false, H, H, 1, &P->attn_out, s));
  return DTTS_OK;
}

}  // namespace ac
}  // namespace dtts

// ---------------------------------------------------------------------------------------------------------------
static size_t ps_plane_cap(const dtts_acoustic* h, int B, int T) {
  const size_t H = h->d.hidden;
  size_t c = std::max<size_t>({(size_t)h->d.ffn_filter, 2 * H, (size_t)h->d.dur_chans, (size_t)(h->ps ? h->ps->word_filter : 0)});
  return (size_t)B * c * tc_rows(T);
}

extern "C" uint64_t dtts_ps_text_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tp, int32_t Tw) {
  if (!h || !h->ps || B <= 0 || Tp <= 0 || Tw <= 0) return 0;
  const size_t H = h->d.hidden, F = std::max<size_t>(h->d.ffn_filter, h->ps->word_filter), C = h->d.dur_chans;
  const size_t bp = (size_t)B * Tp, bw = (size_t)B * Tw, bm = std::max(bp, bw);
  size_t n = 0;
  auto add = [&](size_t floats) { n += ws_round(floats * sizeof(float)); };
  add(bp * H); add(bp * H); add(bp * H);                 // x, x_org / tmp, y
  add(bm * 3 * H); add(bm * H); add(bm * F);             // qkv, att, ffn
  add(bp); add(bp); add(bp); add(B);                     // seq_mask, tok_mask, keep_ph, lens
  add(bp * H);                                           // ph_bct
  add(bw * H); add(bw * H); add(bw * H); add(bw);        // wx, wh, wout, keep_w
  add(bp * C); add(bp * C); add(bp); add(2 * bp);        // d1, d2, dur_ph, dur_int scratch (int64)
  if (h->precision) n += 4 * ws_round(ps_plane_cap(h, B, std::max(Tp, Tw)) * sizeof(tc16));
  return n + 4096;
}

extern "C" int dtts_ps_text_encode(dtts_acoustic* h, const dtts_ps_text_in* in, const dtts_ps_text_out* out, void* ws,
                                   uint64_t ws_bytes, void* stream) {
  if (!h || !in || !out || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_ps_text_encode: null argument");
  if (h->d.model != DTTS_MODEL_PORTASPEECH || !h->ps)
    return fail(DTTS_ERR_BAD_ARG, "dtts_ps_text_encode: the handle does not hold a PortaSpeech model");
  const int B = in->B, Tp = in->Tp, Tw = in->Tw;
  if (B <= 0 || Tp <= 0 || Tw <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_ps_text_encode: empty shape");
  if (!in->txt_tokens_dev || !in->ph2word_dev || !out->ph_encoder_out_dev || !out->word_encoder_out_dev ||
      !out->dur_dev || !out->dur_int_dev || !out->ilens_dev)
    return fail(DTTS_ERR_BAD_ARG, "dtts_ps_text_encode: null tensor");
  if (ws_bytes < dtts_ps_text_workspace_bytes(h, B, Tp, Tw))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_ps_text_encode: workspace too small");
  const dtts_acoustic_desc& d = h->d;
  const PsW& P = *h->ps;
  if (Tw + 1 > P.sin_rows) return fail(DTTS_ERR_BAD_SHAPE, "dtts_ps_text_encode: more than 2048 words per utterance");
  const int H = d.hidden, F = d.ffn_filter, Fw = P.word_filter, C = d.dur_chans, K = d.ffn_kernel;
  const size_t bp = (size_t)B * Tp, bw = (size_t)B * Tw, bm = std::max(bp, bw);
  Bump bump(ws, ws_bytes);
  float* x = bump.take<float>(bp * H);
  float* t1 = bump.take<float>(bp * H);
  float* t2 = bump.take<float>(bp * H);
  float* qkv = bump.take<float>(bm * 3 * H);
  float* att = bump.take<float>(bm * H);
  float* ffn = bump.take<float>(bm * std::max(F, Fw));
  float* seq_mask = bump.take<float>(bp);
  float* tok_mask = bump.take<float>(bp);
  float* keep_ph = bump.take<float>(bp);
  int* lens = bump.take<int>(B);
  float* ph_bct = bump.take<float>(bp * H);
  float* wx = bump.take<float>(bw * H);
  float* wh = bump.take<float>(bw * H);
  float* wout = bump.take<float>(bw * H);
  float* keep_w = bump.take<float>(bw);
  float* d1 = bump.take<float>(bp * C);
  float* d2 = bump.take<float>(bp * C);
  float* dur_ph = bump.take<float>(bp);
  int64_t* dur_int_ph = bump.take<int64_t>(bp);
  TcRun tcr{};
  TcRun* tc = nullptr;
  if (h->precision) {
    tcr.B = B;
    tcr.take(bump, 2, ps_plane_cap(h, B, std::max(Tp, Tw)));
    tc = &tcr;
  }
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_ps_text_encode: workspace too small");
  Launcher L;
  L.stream = (cudaStream_t)stream;
  L.counter = &h->launches;
  cudaStream_t s = L.stream;
  tcr.h = h; tcr.L = &L;
  PsRun R{h, &L, tc, B};
  TcRun::Epi none;

  if (tc && ac_fuse_enabled()) L(l2_prefetch(h->tc_pool, h->tc_text_end * sizeof(tc16), s));   // weights -> L2 (kernels.cuh)
  // ---- TextEncoder.forward (model.py:119-129): embedding * sqrt(H), prefix mask by count, pre-net, post-LN encoder ----
  L(embed_tokens(in->txt_tokens_dev, P.ph_emb, sqrtf((float)H), B, Tp, H, d.ph_size, x, seq_mask, tok_mask, lens, s));
  L(apply_mask(x, seq_mask, B, H, Tp, s));               // x_org * x_mask: every use of it below is masked
  {
    // ConvReluNorm (glow_modules.py:65-72): 3 x {conv k5 on x * mask -> channel LN -> ReLU}, 1x1 proj, + x_org, * mask.
    // Values at padded positions never reach valid ones (every convolution input is masked), so the ReLU output is
    // masked here instead of at the next convolution's input.
    const float* cur = x;
    float* bufs[2] = {t1, t2};
    for (int i = 0; i < PsW::kPreLayers; ++i) {
      R.dense(cur, Tp, P.pre_conv[i], bufs[0], PsW::kPreKernel / 2, none);
      L(channel_layernorm(bufs[0], bufs[1], P.pre_g[i], P.pre_b[i], 1e-4f, nullptr, seq_mask, B, H, Tp, s, 1));
      cur = bufs[1];                                     // the next convolution reads it and overwrites bufs[0]
    }
    TcRun::Epi e = epi_res(x, H, Tp);
    e.mask = seq_mask; e.m_bs = Tp;
    R.dense(cur, Tp, P.pre_proj, x, 0, e);               // x = (x_org + proj(.)) * mask, in place (res aliases out)
  }
  for (size_t i = 0; i < P.enc.layers.size(); ++i) {
    // Encoder.forward with pre_ln = False (rel_transformer_encoder.py:55-79); x is masked on entry
    const EncLayerW& W = P.enc.layers[i];
    DenseW qkv_w{W.qkv, W.t_qkv}, o_w{W.o, W.t_o}, f1_w{W.ffn1, W.t_ffn1}, f2_w{W.ffn2, W.t_ffn2};
    R.dense(x, Tp, qkv_w, qkv, 0, none);
    L(rel_self_attention(qkv, qkv + (size_t)H * Tp, qkv + (size_t)2 * H * Tp, seq_mask, P.rel_k[i], P.rel_v[i],
                         d.rel_window, att, B, H, Tp, d.n_heads, PlaneOut(), s));
    R.dense(att, Tp, o_w, t1, 0, epi_res(x, H, Tp));                                        // x + y
    L(channel_layernorm(t1, x, W.g1, W.b1, 1e-4f, nullptr, seq_mask, B, H, Tp, s));         // norm_layers_1 (* mask: FFN input)
    TcRun::Epi e1;
    e1.act = ACT_RELU; e1.mask = seq_mask; e1.m_bs = Tp;
    R.dense(x, Tp, f1_w, ffn, K / 2, e1);
    TcRun::Epi e2 = epi_res(x, H, Tp);
    e2.mask = seq_mask; e2.m_bs = Tp;
    R.dense(ffn, Tp, f2_w, t1, 0, e2);                                                      // x + ffn(x) * mask
    L(channel_layernorm(t1, x, W.g2, W.b2, 1e-4f, nullptr, seq_mask, B, H, Tp, s));         // norm_layers_2, * mask
  }
  // ret['ph_encoder_out'] = ph_encoder(txt) * src_nonpadding; src_padding of add_dur
  L(ps_finish_ph(x, in->txt_tokens_dev, B, Tp, H, out->ph_encoder_out_dev, ph_bct, keep_ph, s));

  // ---- word level: group_hidden_by_segs + FFTBlocks (tts_modules.py:493-518) ----
  L(ps_group_by_segs(out->ph_encoder_out_dev, in->ph2word_dev, B, Tp, Tw, H, wx, s));
  L(ps_fft_prepare(wx, P.sin_table, P.sin_rows, P.pos_alpha, B, H, Tw, keep_w, s));
  for (size_t i = 0; i < P.word.size(); ++i) {
    const FftLayerW& W = P.word[i];
    L(channel_layernorm(wx, wh, W.ln1_g, W.ln1_b, 1e-5f, nullptr, nullptr, B, H, Tw, s));
    R.dense(wh, Tw, W.qkv, qkv, 0, none);
    L(rel_self_attention(qkv, qkv + (size_t)H * Tw, qkv + (size_t)2 * H * Tw, keep_w, nullptr, nullptr, 0, att, B, H, Tw,
                         d.n_heads, PlaneOut(), s));
    TcRun::Epi eo = epi_res(wx, H, Tw);
    eo.mask = keep_w; eo.m_bs = Tw;
    R.dense(att, Tw, W.o, wx, 0, eo);                                                       // (residual + attn) * keep
    L(channel_layernorm(wx, wh, W.ln2_g, W.ln2_b, 1e-5f, nullptr, nullptr, B, H, Tw, s));
    TcRun::Epi eg;
    eg.act = ACT_GELU;
    R.dense(wh, Tw, W.ffn1, ffn, 0, eg);
    R.dense(ffn, Tw, W.ffn2, wx, 0, eo);                                                    // (residual + ffn) * keep
  }
  L(channel_layernorm(wx, wout, P.word_ln_g, P.word_ln_b, 1e-5f, nullptr, keep_w, B, H, Tw, s));
  L(bct_to_btc(wout, out->word_encoder_out_dev, B, H, Tw, s));

  // ---- add_dur (model.py:317-340): phoneme-level predictor, summed per word ----
  run_dur_predictor(h, ph_bct, keep_ph, d1, d2, B, Tp, dur_ph, dur_int_ph, L, tc);
  L(ps_word_durations(dur_ph, keep_ph, in->ph2word_dev, B, Tp, Tw, out->dur_dev, out->dur_int_dev, out->ilens_dev, s));
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_ps_text_encode: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" uint64_t dtts_ps_attend_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tp, int32_t Tw, int32_t T) {
  if (!h || !h->ps || B <= 0 || Tp <= 0 || Tw <= 0 || T <= 0) return 0;
  const size_t H = h->d.hidden;
  size_t n = 0;
  auto add = [&](size_t floats) { n += ws_round(floats * sizeof(float)); };
  add((size_t)B * 2 * H * Tp); add((size_t)B * H * Tp); add((size_t)B * 2 * H * Tp);      // cat_ph, ph_kv, kv
  add((size_t)B * 2 * H * T); add((size_t)B * H * T); add((size_t)B * H * T); add((size_t)B * H * T);
  add((size_t)B * H * T);                                                                    // cat_q, dec_q, xres, q, ctx
  if (h->precision) n += 4 * ws_round(ps_plane_cap(h, B, std::max(Tp, T)) * sizeof(tc16));
  return n + 4096;
}

extern "C" int dtts_ps_attend(dtts_acoustic* h, const float* ph_enc, const float* word_enc, const int64_t* ph2word,
                              const int64_t* mel2word, int32_t B, int32_t Tp, int32_t Tw, int32_t T, float* attn,
                              float* decoder_inp, float* g_bct, float* x_mask, void* ws, uint64_t ws_bytes, void* stream) {
  if (!h || !ph_enc || !word_enc || !ph2word || !mel2word || !decoder_inp || !g_bct || !x_mask || !ws)
    return fail(DTTS_ERR_BAD_ARG, "dtts_ps_attend: null argument");
  if (h->d.model != DTTS_MODEL_PORTASPEECH || !h->ps)
    return fail(DTTS_ERR_BAD_ARG, "dtts_ps_attend: the handle does not hold a PortaSpeech model");
  if (B <= 0 || Tp <= 0 || Tw <= 0 || T <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_ps_attend: empty shape");
  if (ws_bytes < dtts_ps_attend_workspace_bytes(h, B, Tp, Tw, T))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_ps_attend: workspace too small");
  const PsW& P = *h->ps;
  const int H = h->d.hidden;
  Bump bump(ws, ws_bytes);
  float* cat_ph = bump.take<float>((size_t)B * 2 * H * Tp);
  float* ph_kv = bump.take<float>((size_t)B * H * Tp);
  float* kv = bump.take<float>((size_t)B * 2 * H * Tp);
  float* cat_q = bump.take<float>((size_t)B * 2 * H * T);
  float* dec_q = bump.take<float>((size_t)B * H * T);
  float* xres = bump.take<float>((size_t)B * H * T);
  float* q = bump.take<float>((size_t)B * H * T);
  float* ctx = bump.take<float>((size_t)B * H * T);
  TcRun tcr{};
  TcRun* tc = nullptr;
  if (h->precision) {
    tcr.B = B;
    tcr.take(bump, 2, ps_plane_cap(h, B, std::max(Tp, T)));
    tc = &tcr;
  }
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_ps_attend: workspace too small");
  Launcher L;
  L.stream = (cudaStream_t)stream;
  L.counter = &h->launches;
  cudaStream_t s = L.stream;
  tcr.h = h; tcr.L = &L;
  PsRun R{h, &L, tc, B};
  TcRun::Epi none;

  L(nonpad_mask(mel2word, x_mask, (size_t)B * T, s));                                  // tgt_nonpadding (model.py:254)
  // keys / values: enc_pos_proj([ph_encoder_out ; enc_pos]) then W_k | W_v of the one-head attention (model.py:279,284)
  L(ps_build_cat(ph_enc, 0, ph2word, P.freqs, B, Tp, Tw, H, cat_ph, s));
  R.dense(cat_ph, Tp, P.enc_pos_proj, ph_kv, 0, none);
  R.dense(ph_kv, Tp, P.attn_in, kv, 0, none, H, 2 * H);
  // queries: dec_query_proj([word_encoder_out[mel2word] ; dec_pos]), W_q, * H^-1/2; residual path dec_res_proj
  L(ps_build_cat(word_enc, 1, mel2word, P.freqs, B, T, Tw, H, cat_q, s));
  R.dense(cat_q, T, P.dec_query_proj, dec_q, 0, none);
  TcRun::Epi em;
  em.mask = x_mask; em.m_bs = T;
  R.dense(cat_q, T, P.dec_res_proj, xres, 0, em);                                      // x_res * tgt_nonpadding
  TcRun::Epi eq;
  eq.alpha = 1.f / sqrtf((float)H);
  R.dense(dec_q, T, P.attn_in, q, 0, eq, 0, H);
  L(ps_word_attention(q, kv, mel2word, ph2word, B, H, T, Tp, attn, ctx, s));
  TcRun::Epi eo = epi_res(xres, H, T);
  eo.mask = x_mask; eo.m_bs = T;
  R.dense(ctx, T, P.attn_out, g_bct, 0, eo);                                           // (out_proj(ctx) + x_res) * tgt_nonpadding
  L(bct_to_btc(g_bct, decoder_inp, B, H, T, s));
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_ps_attend: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}
