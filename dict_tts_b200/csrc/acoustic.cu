// Acoustic model behind the C ABI: S2PA text encoder, duration predictor, length regulator, FVAE decoder with its
// residual-coupling prior flow.  Mirrors PortaSpeech_dict.forward(infer=True) (modules/dict_tts/model.py:36-122).
#include "acoustic.cuh"

using namespace dtts;
using namespace dtts::ac;

namespace dtts {
namespace ac {

// conv weight [C_out][C_in][K] (+ optional bias) -> packed
int pack(dtts_acoustic* h, const std::string& name, int C_out, int C_in, int K, bool has_bias, ConvW* cw,
         cudaStream_t s, int rci, int rco, const char* wsuffix) {
  const float* w = h->tab.get(name + wsuffix, (uint64_t)C_out * C_in * K);
  if (!w) return DTTS_ERR_MISSING_WEIGHT;
  const float* b = nullptr;
  if (has_bias) {
    b = h->tab.get(name + ".bias", C_out);
    if (!b) return DTTS_ERR_MISSING_WEIGHT;
    if (rco) {
      float* rb = h->pool.take(C_out);
      if (!rb) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
      DTTS_CUDA(reverse_vec(b, rb, C_out, s));
      b = rb;
    }
  }
  float* dst = h->pool.take((size_t)C_out * C_in * K);
  if (!dst) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
  DTTS_CUDA(repack_conv(w, dst, C_out, C_in, K, rci, rco, s));
  cw->w = dst; cw->bias = b; cw->C_out = C_out; cw->C_in = C_in; cw->ktaps = K; cw->phases = 1;
  return DTTS_OK;
}

int pick_n(int C_out) {
  for (int n = 256; n >= 32; n -= 32)
    if (C_out % n == 0) return n;
  return 0;
}

// Packs `parts` convolutions that share (C_in, K) side by side along C_out into one tensor-core weight set (each part
// becomes whole N-blocks).  w[i] in reference layout [C_out_i][C_in][K] ([C_in][C_out_i][K] when transposed).
int tc_pack(dtts_acoustic* h, const float* const* w, int parts, const float* bias, int C_out_part, int C_in, int K,
            int transposed, int N, TcConvW* cw, cudaStream_t s) {
  if (!h->precision) return DTTS_OK;
  cw->C_in = C_in; cw->C_out = C_out_part * parts;
  cw->N = N ? N : pick_n(C_out_part);
  cw->KC = (C_in % 32 == 0) ? 32 : 16;
  cw->ktaps = K; cw->phases = 1;
  cw->set_mode(h->mode);
  cw->bias = bias;
  if (!cw->N || C_out_part % cw->N || C_in % cw->KC)
    return fail(DTTS_ERR_BAD_SHAPE, "tensor-core acoustic path: unsupported channel count");
  const size_t part_elems = cw->elems() / parts;
  h->tc_used = (h->tc_used + 63) & ~(size_t)63;
  if (h->tc_used + cw->elems() > h->tc_cap) return fail(DTTS_ERR_CUDA, "tensor-core weight pool exhausted");
  tc16* dst = h->tc_pool + h->tc_used;
  cw->w = dst;
  for (int i = 0; i < parts; ++i)
    DTTS_CUDA(tc_pack_weights(w[i], dst + (size_t)i * part_elems, C_out_part, C_in, K, transposed, 1, cw->N, cw->KC,
                              cw->planes, cw->fmt, cw->stack, s, 0, cw->pair));
  h->tc_used += cw->elems();
  return DTTS_OK;
}
int tc_pack1(dtts_acoustic* h, const std::string& name, bool has_bias, int C_out, int C_in, int K, int N, TcConvW* cw,
             cudaStream_t s, int transposed) {
  if (!h->precision) return DTTS_OK;
  const float* w = h->tab.get(name + ".weight", (uint64_t)C_out * C_in * K);
  if (!w) return DTTS_ERR_MISSING_WEIGHT;
  const float* b = nullptr;
  if (has_bias && !(b = h->tab.get(name + ".bias", C_out))) return DTTS_ERR_MISSING_WEIGHT;
  return tc_pack(h, &w, 1, b, C_out, C_in, K, transposed, N, cw, s);
}

int pack_qkv(dtts_acoustic* h, const std::string names[3], const std::string* bias_names, ConvW* qkv, TcConvW* t_qkv,
             cudaStream_t s) {
  const int H = h->d.hidden;
  // q, k, v projections packed side by side: one [H][1][3H] weight, one launch
  float* wqkv = h->pool.take((size_t)3 * H * H);
  float* bqkv = bias_names ? h->pool.take((size_t)3 * H) : nullptr;
  float* tmp = h->pool.take((size_t)H * H);
  if (!wqkv || (bias_names && !bqkv) || !tmp) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
  const float* ws[3];
  for (int j = 0; j < 3; ++j) {
    ws[j] = h->tab.get(names[j], (uint64_t)H * H);
    if (!ws[j]) return DTTS_ERR_MISSING_WEIGHT;
    DTTS_CUDA(repack_conv(ws[j], tmp, H, H, 1, 0, 0, s));                    // [ci][co]
    DTTS_CUDA(cudaMemcpy2DAsync(wqkv + j * H, 3 * H * sizeof(float), tmp, H * sizeof(float), H * sizeof(float), H,
                                cudaMemcpyDeviceToDevice, s));
    if (bias_names) {
      const float* b = h->tab.get(bias_names[j], H);
      if (!b) return DTTS_ERR_MISSING_WEIGHT;
      DTTS_CUDA(cudaMemcpyAsync(bqkv + j * H, b, H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  }
  qkv->w = wqkv; qkv->bias = bqkv; qkv->C_out = 3 * H; qkv->C_in = H; qkv->ktaps = 1; qkv->phases = 1;
  if (h->precision) DTTS_TRY(tc_pack(h, ws, 3, bqkv, H, H, 1, 0, 0, t_qkv, s));
  return DTTS_OK;
}

int pack_encoder(dtts_acoustic* h, const std::string& p, EncoderW* e, cudaStream_t s) {
  const int H = h->d.hidden, F = h->d.ffn_filter, K = h->d.ffn_kernel;
  for (int i = 0; i < h->d.enc_layers; ++i) {
    EncLayerW L;
    const std::string a = p + ".attn_layers." + std::to_string(i);
    const std::string wn[3] = {a + ".conv_q.weight", a + ".conv_k.weight", a + ".conv_v.weight"};
    const std::string bn[3] = {a + ".conv_q.bias", a + ".conv_k.bias", a + ".conv_v.bias"};
    DTTS_TRY(pack_qkv(h, wn, bn, &L.qkv, &L.t_qkv, s));
    DTTS_TRY(pack(h, a + ".conv_o", H, H, 1, true, &L.o, s));
    const std::string f = p + ".ffn_layers." + std::to_string(i);
    DTTS_TRY(pack(h, f + ".conv_1", F, H, K, true, &L.ffn1, s));
    DTTS_TRY(pack(h, f + ".conv_2", H, F, 1, true, &L.ffn2, s));
    if (h->precision) {
      DTTS_TRY(tc_pack1(h, a + ".conv_o", true, H, H, 1, 0, &L.t_o, s));
      DTTS_TRY(tc_pack1(h, f + ".conv_1", true, F, H, K, 0, &L.t_ffn1, s));
      DTTS_TRY(tc_pack1(h, f + ".conv_2", true, H, F, 1, 0, &L.t_ffn2, s));
    }
    L.g1 = h->tab.get(p + ".norm_layers_1." + std::to_string(i) + ".gamma", H);
    L.b1 = h->tab.get(p + ".norm_layers_1." + std::to_string(i) + ".beta", H);
    L.g2 = h->tab.get(p + ".norm_layers_2." + std::to_string(i) + ".gamma", H);
    L.b2 = h->tab.get(p + ".norm_layers_2." + std::to_string(i) + ".beta", H);
    if (!L.g1 || !L.b1 || !L.g2 || !L.b2) return DTTS_ERR_MISSING_WEIGHT;
    e->layers.push_back(L);
  }
  if (h->tab.entries.count(p + ".last_ln.gamma")) {          // pre-LN encoders only (rel_transformer_encoder.py:52-53)
    e->last_g = h->tab.get(p + ".last_ln.gamma", H);
    e->last_b = h->tab.get(p + ".last_ln.beta", H);
    if (!e->last_g || !e->last_b) return DTTS_ERR_MISSING_WEIGHT;
  } else {
    e->last_g = e->last_b = nullptr;
  }
  return DTTS_OK;
}

int pack_wn(dtts_acoustic* h, const std::string& p, int hidden, int n_layers, int K, int gin, WNW* wn,
            cudaStream_t s, bool gated) {
  DTTS_TRY(pack(h, p + ".cond_layer", 2 * hidden * n_layers, gin, 1, true, &wn->cond, s));
  for (int i = 0; i < n_layers; ++i) {
    ConvW a, r;
    DTTS_TRY(pack(h, p + ".in_layers." + std::to_string(i), 2 * hidden, hidden, K, true, &a, s));
    const int rs = (i < n_layers - 1) ? 2 * hidden : hidden;
    DTTS_TRY(pack(h, p + ".res_skip_layers." + std::to_string(i), rs, hidden, 1, true, &r, s));
    wn->in_layers.push_back(a);
    wn->res_skip.push_back(r);
    if (h->precision) {
      TcConvW ta, tr;
      DTTS_TRY(tc_pack1(h, p + ".in_layers." + std::to_string(i), true, 2 * hidden, hidden, K, 0, &ta, s));
      DTTS_TRY(tc_pack1(h, p + ".res_skip_layers." + std::to_string(i), true, rs, hidden, 1, hidden, &tr, s));
      wn->t_in.push_back(ta);
      wn->t_rs.push_back(tr);
    }
  }
  DTTS_TRY(tc_pack1(h, p + ".cond_layer", true, 2 * hidden * n_layers, gin, 1, 0, &wn->t_cond, s));
  const int Ng = pick_n(2 * hidden);
  if (gated && h->precision && Ng % 64 == 0 && hidden % (Ng / 2) == 0) {
    // gate in the epilogue (TcConvParams::gate): in_layers and cond_layer once more with [tanh | sigmoid] halves per N block
    const size_t in_n = (size_t)2 * hidden * hidden * K, cond_n = (size_t)2 * hidden * n_layers * gin;
    for (int i = 0; i < n_layers; ++i) {
      const std::string q = p + ".in_layers." + std::to_string(i);
      const float* w = h->tab.get(q + ".weight", in_n);
      const float* b = h->tab.get(q + ".bias", 2 * hidden);
      float* wp = h->pool.take(in_n);
      float* bp = h->pool.take(2 * hidden);
      if (!w || !b) return DTTS_ERR_MISSING_WEIGHT;
      if (!wp || !bp) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
      DTTS_CUDA(permute_gate_rows(w, wp, 1, hidden, Ng, (long)hidden * K, s));
      DTTS_CUDA(permute_gate_rows(b, bp, 1, hidden, Ng, 1, s));
      TcConvW t;
      const float* wpc = wp;
      DTTS_TRY(tc_pack(h, &wpc, 1, bp, 2 * hidden, hidden, K, 0, Ng, &t, s));
      wn->t_in_g.push_back(t);
    }
    const float* cw = h->tab.get(p + ".cond_layer.weight", cond_n);
    const float* cb = h->tab.get(p + ".cond_layer.bias", (uint64_t)2 * hidden * n_layers);
    float* cwp = h->pool.take(cond_n);
    float* cbp = h->pool.take((size_t)2 * hidden * n_layers);
    if (!cw || !cb) return DTTS_ERR_MISSING_WEIGHT;
    if (!cwp || !cbp) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    DTTS_CUDA(permute_gate_rows(cw, cwp, n_layers, hidden, Ng, gin, s));
    DTTS_CUDA(permute_gate_rows(cb, cbp, n_layers, hidden, Ng, 1, s));
    const float* cwpc = cwp;
    DTTS_TRY(tc_pack(h, &cwpc, 1, cbp, 2 * hidden * n_layers, gin, 1, 0, wn->t_cond.N, &wn->t_cond_g, s));
  }
  return DTTS_OK;
}

// Pre-LN transformer encoder (rel_transformer_encoder.py:55-79).  x is updated in place; the result is written to hbuf.
void run_encoder(dtts_acoustic* h, const EncoderW& E, float* x, float* hbuf, float* qkv, float* att, float* ffn,
                 const float* seq_mask, int B, int Tw, Launcher& L, TcRun* tc) {
  const int H = h->d.hidden, F = h->d.ffn_filter, K = h->d.ffn_kernel;
  cudaStream_t s = L.stream;
  if (tc) {
    Planes &P0 = tc->P[0], &P1 = tc->P[1];
    TcRun::Epi res_x;
    res_x.res = x; res_x.r_bs = (long)H * Tw; res_x.r_cs = Tw; res_x.r_ts = 1;
    // Fused form: the LayerNorm that follows a residual update runs in the epilogue of the convolution that makes it
    // (O -> LN2, FFN2 -> LN1 of the next layer / the last LN): one standalone LayerNorm per encoder instead of nine
    const size_t nl = E.layers.size();
    const bool fuse_ln = ac_fuse_enabled() && nl > 0 && TcRun::ln_fusable(E.layers[0].t_o, Tw) &&
                         TcRun::ln_fusable(E.layers[0].t_ffn2, Tw);
    for (size_t i = 0; i < nl; ++i) {
      const EncLayerW& W = E.layers[i];
      // x = x * x_mask ; LN1 -> operand planes of the fused q|k|v projection
      if (!fuse_ln || i == 0)
        L(channel_layernorm_planes(x, x, nullptr, W.g1, W.b1, 1e-4f, seq_mask, nullptr, B, H, Tw,
                                   tc->out_of(P0, H, Tw, fuse_ln), s));       // (fused: halo zeroed once, for every FFN)
      tc->conv_nct(P0, W.t_qkv, qkv, Tw, 1, 0, TcRun::Epi());
      if (self_attention_planes_fits(H, Tw, h->d.n_heads)) {
        L(self_attention_planes(qkv, qkv + (size_t)H * Tw, qkv + (size_t)2 * H * Tw, seq_mask, nullptr, B, H, Tw,
                                h->d.n_heads, tc->out_of(P0, H, Tw, false), s));
      } else {
        L(self_attention(qkv, qkv + (size_t)H * Tw, qkv + (size_t)2 * H * Tw, seq_mask, att, B, H, Tw, h->d.n_heads, s));
        tc->stage_nct(P0, att, H, Tw);
      }
      if (fuse_ln) {
        TcRun::Epi eo = res_x;                                                // x += O(att) ; LN2(x) * mask -> P0 (rows [0,T))
        eo.ln_gamma = W.g2; eo.ln_beta = W.b2; eo.ln_eps = 1e-4f; eo.ln_out_mask = seq_mask;
        tc->conv_nct(P0, W.t_o, x, Tw, 1, 0, eo, 0, 0, &P0);
      } else {
        tc->conv_nct(P0, W.t_o, x, Tw, 1, 0, res_x);
        // LN2 (* x_mask) -> planes with zeroed halo (the FFN's first convolution has k = 5)
        L(channel_layernorm_planes(x, nullptr, nullptr, W.g2, W.b2, 1e-4f, nullptr, seq_mask, B, H, Tw,
                                   tc->out_of(P0, H, Tw, true), s));
      }
      TcRun::Epi e1;
      e1.act = 1; e1.mask = seq_mask; e1.m_bs = Tw;
      tc->conv_nct(P0, W.t_ffn1, nullptr, Tw, 1, K / 2, e1, 0, 0, &P1);      // relu(.) * mask straight into planes
      TcRun::Epi e2 = res_x;
      e2.mask = seq_mask; e2.m_bs = Tw;
      if (fuse_ln) {
        if (i + 1 < nl) {                    // next layer: x = x * x_mask ; LN1(x) -> P0
          e2.ln_gamma = E.layers[i + 1].g1; e2.ln_beta = E.layers[i + 1].b1; e2.ln_in_mask = seq_mask;
        } else {                             // last LN: fp32 [B,H,T] for the consumers and planes for the S2PA query projection
          e2.ln_gamma = E.last_g; e2.ln_beta = E.last_b; e2.ln_out_mask = seq_mask; e2.ln_y = hbuf;
        }
        e2.ln_eps = 1e-4f;
        tc->conv_nct(P1, W.t_ffn2, x, Tw, 1, 0, e2, 0, 0, &P0);
      } else {
        tc->conv_nct(P1, W.t_ffn2, x, Tw, 1, 0, e2);
      }
    }
    // last LN: fp32 [B,H,T] for the consumers and planes (P0) for the S2PA query projection
    if (!fuse_ln)
      L(channel_layernorm_planes(x, nullptr, hbuf, E.last_g, E.last_b, 1e-4f, nullptr, seq_mask, B, H, Tw,
                                 tc->out_of(P0, H, Tw, false), s));
    return;
  }
  for (size_t i = 0; i < E.layers.size(); ++i) {
    const EncLayerW& W = E.layers[i];
    L(apply_mask(x, seq_mask, B, H, Tw, s));
    L(channel_layernorm(x, hbuf, W.g1, W.b1, 1e-4f, nullptr, nullptr, B, H, Tw, s));
    L(launch_conv1d_f32(conv_params(hbuf, Tw, W.qkv, 0, 3 * H, qkv, Tw, 1, 1, 0), B, s));
    L(self_attention(qkv, qkv + (size_t)H * Tw, qkv + (size_t)2 * H * Tw, seq_mask, att, B, H, Tw, h->d.n_heads, s));
    {
      ConvParams p = conv_params(att, Tw, W.o, 0, H, x, Tw, 1, 1, 0);
      p.res = x; p.r_bs = (long)H * Tw; p.r_cs = Tw; p.r_ts = 1;
      L(launch_conv1d_f32(p, B, s));
    }
    L(channel_layernorm(x, hbuf, W.g2, W.b2, 1e-4f, nullptr, seq_mask, B, H, Tw, s));   // FFN input is x * x_mask
    {
      ConvParams p = conv_params(hbuf, Tw, W.ffn1, 0, F, ffn, Tw, 1, 1, K / 2);
      p.act = ACT_RELU; p.mask = seq_mask; p.m_bs = Tw;
      L(launch_conv1d_f32(p, B, s));
    }
    {
      ConvParams p = conv_params(ffn, Tw, W.ffn2, 0, H, x, Tw, 1, 1, 0);
      p.mask = seq_mask; p.m_bs = Tw;
      p.res = x; p.r_bs = (long)H * Tw; p.r_cs = Tw; p.r_ts = 1;
      L(launch_conv1d_f32(p, B, s));
    }
  }
  L(channel_layernorm(x, hbuf, E.last_g, E.last_b, 1e-4f, nullptr, seq_mask, B, H, Tw, s));
}

// WN.forward with x_mask = 1 (modules/commons/wavenet.py:54-78).  hx [B,hidden,T] is updated in place,
// skip [B,hidden,T] receives the output.
// x_planes_ready: tc->P[1] already holds hx as operand planes with a zeroed halo (its producer wrote them);
// skip_po (optional): receives the final skip sum as operand planes (the consumer's staging launch is then not needed).
void run_wn(const WNW& W, int hidden, int K, float* hx, const float* g, int gin, float* cond, float* a, float* acts,
            float* skip, int B, int T, Launcher& L, TcRun* tc, bool x_planes_ready = false, Planes* skip_po = nullptr) {
  cudaStream_t s = L.stream;
  const int n = (int)W.in_layers.size();
  if (tc) {
    Planes &Pg = tc->P[0], &Px = tc->P[1], &Pa = tc->P[2];
    const bool gate_epi = ac_fuse_enabled() && (int)W.t_in_g.size() == n && W.t_cond_g.w;
    tc->stage_nct(Pg, g, gin, T);
    tc->conv_nct(Pg, gate_epi ? W.t_cond_g : W.t_cond, cond, T, 1, 0, TcRun::Epi());
    if (!x_planes_ready) tc->stage_nct(Px, hx, hidden, T);              // halo zeroed once; epilogues refresh rows [0,T)
    for (int i = 0; i < n; ++i) {
      TcRun::Epi ea;
      ea.res = cond + (size_t)2 * hidden * i * T; ea.r_bs = (long)W.cond.C_out * T; ea.r_cs = T; ea.r_ts = 1;
      if (gate_epi) {                                                     // tanh * sigmoid in the epilogue -> planes Pa
        ea.gate = 1;
        tc->conv_nct(Px, W.t_in_g[i], nullptr, T, 1, K / 2, ea, 0, 0, &Pa);
      } else {
        tc->conv_nct(Px, W.t_in[i], a, T, 1, K / 2, ea);
        L(wn_gate_planes(a, B, hidden, T, tc->out_of(Pa, hidden, T, false), s));
      }
      TcRun::Epi ex, es;
      ex.res = hx; ex.r_bs = (long)hidden * T; ex.r_cs = T; ex.r_ts = 1;
      es.accumulate = (i > 0);
      if (i < n - 1) {
        tc->conv_nct(Pa, W.t_rs[i], hx, T, 1, 0, ex, 0, 1, &Px);        // x = x + rs[:hidden]  (fp32 + planes)
        tc->conv_nct(Pa, W.t_rs[i], skip, T, 1, 0, es, 1, 1);           // out += rs[hidden:]
      } else {
        tc->conv_nct(Pa, W.t_rs[i], skip, T, 1, 0, es, 0, 1, skip_po);
      }
    }
    return;
  }
  (void)gin;
  L(launch_conv1d_f32(conv_params(g, T, W.cond, 0, W.cond.C_out, cond, T, 1, 1, 0), B, s));
  for (int i = 0; i < n; ++i) {
    {
      ConvParams p = conv_params(hx, T, W.in_layers[i], 0, 2 * hidden, a, T, 1, 1, K / 2);
      p.res = cond + (size_t)2 * hidden * i * T; p.r_bs = (long)W.cond.C_out * T; p.r_cs = T; p.r_ts = 1;
      L(launch_conv1d_f32(p, B, s));
    }
    L(wn_gate(a, acts, B, hidden, T, s));
    if (i < n - 1) {
      ConvParams p = conv_params(acts, T, W.res_skip[i], 0, hidden, hx, T, 1, 1, 0);          // x = x + rs[:hidden]
      p.res = hx; p.r_bs = (long)hidden * T; p.r_cs = T; p.r_ts = 1;
      L(launch_conv1d_f32(p, B, s));
      ConvParams q = conv_params(acts, T, W.res_skip[i], hidden, hidden, skip, T, 1, 1, 0);    // out += rs[hidden:]
      q.accumulate = (i > 0);
      L(launch_conv1d_f32(q, B, s));
    } else {
      ConvParams q = conv_params(acts, T, W.res_skip[i], 0, hidden, skip, T, 1, 1, 0);
      q.accumulate = (i > 0);
      L(launch_conv1d_f32(q, B, s));
    }
  }
}

void run_dur_predictor(dtts_acoustic* h, const float* dur_in, const float* keep, float* d1, float* d2, int B, int Tw,
                       float* dur, int64_t* dur_int, Launcher& L, TcRun* tc) {
  const dtts_acoustic_desc& d = h->d;
  const int C = d.dur_chans;
  cudaStream_t s = L.stream;
  const int C_in = h->dur_conv.empty() ? 0 : h->dur_conv[0].C_in;
  const float* cur = dur_in;
  float* bufs[2] = {d1, d2};
  if (tc) tc->stage_nct(tc->P[0], dur_in, C_in, Tw);
  for (int i = 0; i < d.dur_layers; ++i) {
    float* c = bufs[0];
    float* y = bufs[1];
    if (tc) {
      TcRun::Epi er;
      er.act = 1;
      tc->conv_nct(tc->P[0], h->t_dur[i], c, Tw, 1, (d.dur_kernel - 1) / 2, er);
      const bool last = i == d.dur_layers - 1;              // LayerNorm -> next convolution's planes (k = 5: zero halo)
      L(channel_layernorm_planes(c, nullptr, y, h->dur_ln_g[i], h->dur_ln_b[i], 1e-5f, nullptr, keep, B, C, Tw,
                                 last ? PlaneOut() : tc->out_of(tc->P[0], C, Tw, true), s));
    } else {
      ConvParams p = conv_params(cur, Tw, h->dur_conv[i], 0, C, c, Tw, 1, 1, (d.dur_kernel - 1) / 2);
      p.act = ACT_RELU;
      L(launch_conv1d_f32(p, B, s));
      L(channel_layernorm(c, y, h->dur_ln_g[i], h->dur_ln_b[i], 1e-5f, nullptr, keep, B, C, Tw, s));
    }
    cur = y;
    bufs[0] = c;      // conv output buffer can be reused: next conv reads y, writes c
  }
  L(dur_head(cur, h->dur_w, h->dur_b, keep, B, C, Tw, dur, dur_int, s));
}

int pack_dur_predictor(dtts_acoustic* h, int c_in, cudaStream_t s) {
  const dtts_acoustic_desc* d = &h->d;
  const int H = c_in;
  for (int i = 0; i < d->dur_layers; ++i) {
    ConvW c;
    const std::string q = "dur_predictor.conv." + std::to_string(i);
    DTTS_TRY(pack(h, q + ".1", d->dur_chans, i == 0 ? H : d->dur_chans, d->dur_kernel, true, &c, s));
    h->dur_conv.push_back(c);
    if (h->precision) {
      TcConvW tcw;
      DTTS_TRY(tc_pack1(h, q + ".1", true, d->dur_chans, i == 0 ? H : d->dur_chans, d->dur_kernel, 0, &tcw, s));
      h->t_dur.push_back(tcw);
    }
    const float* g = h->tab.get(q + ".3.weight", d->dur_chans);
    const float* b = h->tab.get(q + ".3.bias", d->dur_chans);
    if (!g || !b) return DTTS_ERR_MISSING_WEIGHT;
    h->dur_ln_g.push_back(g);
    h->dur_ln_b.push_back(b);
  }
  h->dur_w = h->tab.get("dur_predictor.linear.0.weight", d->dur_chans);
  h->dur_b = h->tab.get("dur_predictor.linear.0.bias", 1);
  if (!h->dur_w || !h->dur_b) return DTTS_ERR_MISSING_WEIGHT;
  return DTTS_OK;
}

int pack_decoder(dtts_acoustic* h, cudaStream_t s) {
  const dtts_acoustic_desc* d = &h->d;
  const int H = d->hidden;
  DTTS_TRY(pack(h, "fvae.g_pre_net.0", H, H, 8, true, &h->g_pre, s));
  if (h->precision) {
    // Conv1d(H,H,k=8,s=4,p=2) == Conv1d(4H,H,k=3,p=1) over x'[(s,ci), q] = g[ci, 4q+s] with
    // w'[co,(s,ci),a] = w[co,ci,4(a-1)+s+2] (zero where that tap does not exist)
    const float* w8 = h->tab.get("fvae.g_pre_net.0.weight", (uint64_t)H * H * 8);
    const float* b8 = h->tab.get("fvae.g_pre_net.0.bias", H);
    if (!w8 || !b8) return DTTS_ERR_MISSING_WEIGHT;
    float* w3 = h->pool.take((size_t)H * 4 * H * 3);
    if (!w3) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    DTTS_CUDA(repack_s2d4(w8, w3, H, H, s));
    const float* w3c = w3;
    DTTS_TRY(tc_pack(h, &w3c, 1, b8, H, 4 * H, 3, 0, 0, &h->t_gpre, s));
  }
  const int half = d->latent / 2;
  for (int f = 0; f < d->flow_blocks; ++f) {
    FlowW F;
    // reversed(flows) = Flip, RCL_{n-1}, Flip, RCL_{n-2}, ...: RCL_f runs after (n - f) flips.
    F.odd = ((d->flow_blocks - f) & 1);
    const std::string q = "fvae.prior_flow.flows." + std::to_string(2 * f);
    DTTS_TRY(pack(h, q + ".pre", d->flow_hidden, half, 1, true, &F.pre, s, F.odd, 0));
    DTTS_TRY(pack(h, q + ".post", half, d->flow_hidden, 1, true, &F.post, s, 0, F.odd));
    DTTS_TRY(pack_wn(h, q + ".enc", d->flow_hidden, d->flow_layers, d->flow_kernel, H, &F.wn, s));
    h->flows.push_back(F);
  }
  if (h->precision && d->flow_blocks > 0 &&
      flow_fused_supported(H, d->flow_hidden, d->latent, d->flow_kernel, d->flow_layers, d->flow_blocks)) {
    // the same weights once more as the stream of flow_fused_kernel, in execution order (reverse pass: last layer first)
    FlowFusedW& W = h->flow_fused;
    W.n_flows = d->flow_blocks; W.n_layers = d->flow_layers; W.n_chunks = H / 64; W.H = H;
    W.par_stride = flow_fused_par_floats(d->flow_layers);
    W.par = h->pool.take((size_t)W.par_stride * W.n_flows);
    if (!W.par) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    if (cudaMalloc((void**)&W.stream, W.stream_bytes()) != cudaSuccess) {
      W.stream = nullptr;
      return fail(DTTS_ERR_CUDA, "fused prior flow: weight stream allocation failed");
    }
    const int FH = d->flow_hidden, Ln = d->flow_layers;
    for (int e = 0; e < W.n_flows; ++e) {
      const int f = d->flow_blocks - 1 - e;
      const FlowW& F = h->flows[f];
      if (F.odd) W.odd_mask |= 1u << e;
      const std::string q = "fvae.prior_flow.flows." + std::to_string(2 * f) + ".enc";
      const float* cw = h->tab.get(q + ".cond_layer.weight", (uint64_t)2 * FH * Ln * H);
      FlowParSrc src{};
      src.n_layers = Ln;
      src.cond_b = h->tab.get(q + ".cond_layer.bias", (uint64_t)2 * FH * Ln);
      if (!cw || !src.cond_b) return DTTS_ERR_MISSING_WEIGHT;
      for (int i = 0; i < Ln; ++i) {
        const int rs = (i < Ln - 1) ? 2 * FH : FH;
        const std::string li = std::to_string(i);
        const float* iw = h->tab.get(q + ".in_layers." + li + ".weight", (uint64_t)2 * FH * FH * d->flow_kernel);
        const float* rw = h->tab.get(q + ".res_skip_layers." + li + ".weight", (uint64_t)rs * FH);
        src.in_b[i] = h->tab.get(q + ".in_layers." + li + ".bias", 2 * FH);
        src.rs_b[i] = h->tab.get(q + ".res_skip_layers." + li + ".bias", rs);
        if (!iw || !rw || !src.in_b[i] || !src.rs_b[i]) return DTTS_ERR_MISSING_WEIGHT;
        DTTS_CUDA(flow_fused_pack_layer(iw, cw + (size_t)i * 2 * FH * H, rw, rs, H,
                                        W.stream + ((size_t)e * Ln + i) * (W.n_chunks + 4) * 32768, s));
      }
      src.pre_w = F.pre.w; src.pre_b = F.pre.bias; src.post_w = F.post.w; src.post_b = F.post.bias;
      DTTS_CUDA(flow_fused_pack_par(src, W.par + (size_t)e * W.par_stride, s));
    }
  }
  {
    // ConvTranspose1d(latent -> H, k=4, s=4): weight [latent][H][4]
    const float* w = h->tab.get("fvae.decoder.pre_net.0.weight", (uint64_t)d->latent * H * 4);
    const float* b = h->tab.get("fvae.decoder.pre_net.0.bias", H);
    if (!w || !b) return DTTS_ERR_MISSING_WEIGHT;
    float* dst = h->pool.take((size_t)d->latent * H * 4);
    if (!dst) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    DTTS_CUDA(repack_convT(w, dst, d->latent, H, 4, 4, s));
    h->dec_pre.w = dst; h->dec_pre.bias = b; h->dec_pre.C_out = H; h->dec_pre.C_in = d->latent;
    h->dec_pre.ktaps = 1; h->dec_pre.phases = 4;
  }
  DTTS_TRY(pack_wn(h, "fvae.decoder.wn", H, d->dec_layers, d->dec_kernel, H, &h->dec_wn, s, true));
  DTTS_TRY(pack(h, "fvae.decoder.out_proj", d->n_mel, H, 1, true, &h->dec_out, s));
  if (h->precision) {
    // out_proj: C_out = n_mel (80) padded with zero rows to a multiple of 32 for the MMA's N
    const int np = (d->n_mel + 31) / 32 * 32;
    const float* wo = h->tab.get("fvae.decoder.out_proj.weight", (uint64_t)d->n_mel * H);
    const float* bo = h->tab.get("fvae.decoder.out_proj.bias", d->n_mel);
    if (!wo || !bo) return DTTS_ERR_MISSING_WEIGHT;
    float* wp = h->pool.take((size_t)np * H);
    float* bp = h->pool.take(np);
    if (!wp || !bp) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
    DTTS_CUDA(cudaMemsetAsync(wp, 0, (size_t)np * H * sizeof(float), s));
    DTTS_CUDA(cudaMemsetAsync(bp, 0, np * sizeof(float), s));
    DTTS_CUDA(cudaMemcpyAsync(wp, wo, (size_t)d->n_mel * H * sizeof(float), cudaMemcpyDeviceToDevice, s));
    DTTS_CUDA(cudaMemcpyAsync(bp, bo, d->n_mel * sizeof(float), cudaMemcpyDeviceToDevice, s));
    const float* wpc = wp;
    DTTS_TRY(tc_pack(h, &wpc, 1, bp, np, H, 1, 0, 0, &h->t_out, s));
    h->out_pad = np;
  }
  return DTTS_OK;
}

}  // namespace ac
}  // namespace dtts

extern "C" int dtts_acoustic_create(const dtts_acoustic_desc* d, const float* arena_dev, uint64_t arena_floats,
                                    const dtts_weight_entry* table, int32_t n_entries, void* stream,
                                    dtts_acoustic** out) {
  if (!d || !out) return fail(DTTS_ERR_BAD_ARG, "null descriptor/out");
  *out = nullptr;
  if (d->hidden <= 0 || d->hidden > 256 || d->n_heads <= 0 || d->hidden % d->n_heads || d->dict_dim <= 0 ||
      d->dict_dim % 4 || d->latent <= 0 || d->latent % 2 || d->frames_multiple != 4 || d->enc_layers < 0 ||
      d->ffn_kernel <= 0 || d->ffn_filter <= 0 || d->dur_layers < 0 || d->dur_chans <= 0 || d->flow_blocks < 0 ||
      d->flow_hidden <= 0 || d->n_mel <= 0 || d->word_size <= 0 || d->pinyin_size <= 0)
    return fail(DTTS_ERR_BAD_SHAPE, "unsupported acoustic configuration");
  if (d->model != DTTS_MODEL_DICT && d->model != DTTS_MODEL_PORTASPEECH)
    return fail(DTTS_ERR_BAD_ARG, "dtts_acoustic_desc.model must be DTTS_MODEL_DICT or DTTS_MODEL_PORTASPEECH");
  if (d->model == DTTS_MODEL_PORTASPEECH && (d->ph_size <= 0 || d->word_enc_layers < 0 || d->rel_window < 0 ||
                                             d->hidden % 2 || d->hidden < 4))
    return fail(DTTS_ERR_BAD_SHAPE, "unsupported PortaSpeech configuration");
  if (d->precision != 0 && d->precision != 1)
    return fail(DTTS_ERR_BAD_ARG, "acoustic precision must be 0 (fp32 FMA) or 1 (tcgen05, bf16 hi/lo split)");
  if (d->s2pa_route != 0 && d->s2pa_route != 1)
    return fail(DTTS_ERR_BAD_ARG, "s2pa_route must be 0 (folded streaming pass) or 1 (K/V projection GEMM)");
  if (d->s2pa_route == 1 && (d->precision != 1 || d->hidden % 32 || d->dict_dim % 32))
    return fail(DTTS_ERR_BAD_ARG, "s2pa_route = 1 runs on tcgen05: it needs precision = 1, hidden and dict_dim % 32 == 0");
  DTTS_TRY(arch_check());
  dtts_acoustic* h = new dtts_acoustic();
  h->d = *d;
  h->precision = d->precision;
  h->mode = tc_mode(1);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = h->tab.init(arena_dev, arena_floats, table, n_entries);
  if (rc != DTTS_OK) { delete h; return rc; }
  size_t total = 0;
  for (auto& e : h->tab.entries) total += e.second.second + 64;
  // slack: q|k|v side-by-side copies, the space-to-depth g_pre_net weight, the gate-permuted decoder WaveNet, alignment
  const size_t gate_extra = (size_t)d->dec_layers * 2 * d->hidden * d->hidden * (d->dec_kernel + 1) + 4096 +
                            (size_t)2 * d->hidden * d->dict_dim;     // + the pre-multiplied S2PA projections
  rc = h->pool.reserve(total + 2 * 1024 * 1024 + gate_extra);
  if (rc != DTTS_OK) { delete h; return rc; }
  if (h->precision) {
    h->tc_cap = 2 * total + (1 << 20) + 2 * gate_extra;   // two 16-bit planes per weight (+ the gate-permuted copies)
    if (d->s2pa_route == 1) h->tc_cap += (size_t)4 * d->hidden * d->dict_dim + 256;   // second copy of W_k, W_v
    cudaError_t e = cudaMalloc((void**)&h->tc_pool, h->tc_cap * sizeof(tc16));
    if (e == cudaSuccess) e = tc_conv_init();
    if (e != cudaSuccess) {
      h->pool.release();
      delete h;
      return fail(DTTS_ERR_CUDA, std::string("acoustic tensor-core pool: ") + cudaGetErrorString(e));
    }
  }
  auto build = [&]() -> int {
    const int H = d->hidden, D = d->dict_dim;
    if (d->model == DTTS_MODEL_PORTASPEECH) {          // own text side, shared predictor / decoder
      DTTS_TRY(create_ps(h, s));
      DTTS_TRY(pack_dur_predictor(h, H, s));
      h->tc_text_end = h->tc_used;
      return pack_decoder(h, s);
    }
    const std::string p = "dict_encoder.S2PA_module";
    h->word_emb = h->tab.get(p + ".word_emb.weight", (uint64_t)d->word_size * H);
    h->pinyin_emb = h->tab.get(p + ".s2pa_attention.pinyin_embedding.weight", (uint64_t)d->pinyin_size * H);
    if (!h->word_emb || !h->pinyin_emb) return DTTS_ERR_MISSING_WEIGHT;
    DTTS_TRY(pack_encoder(h, p + ".semantic_encoder", &h->sem, s));
    DTTS_TRY(pack_encoder(h, p + ".linguistic_encoder", &h->lin, s));
    if (!h->sem.last_g || !h->lin.last_g) return fail(DTTS_ERR_MISSING_WEIGHT, "missing weight: <encoder>.last_ln.gamma");
    const std::string a = p + ".s2pa_attention";
    DTTS_TRY(pack(h, a + ".q_transform", H, H, 1, false, &h->s2pa_q, s));
    DTTS_TRY(pack(h, a + ".v_transform", H, D, 1, false, &h->s2pa_v, s));
    DTTS_TRY(pack(h, a + ".output_transform", H, H, 1, false, &h->s2pa_o, s));
    {
      // k_transform.weight is [H][D]; the folded form needs W_k^T as a 1x1 conv H -> D, packed [ci=H][D] = as stored.
      const float* wk = h->tab.get(a + ".k_transform.weight", (uint64_t)H * D);
      if (!wk) return DTTS_ERR_MISSING_WEIGHT;
      h->s2pa_kT.w = wk; h->s2pa_kT.bias = nullptr; h->s2pa_kT.C_out = D; h->s2pa_kT.C_in = H; h->s2pa_kT.ktaps = 1;
      // tensor-core copy: [H][D] is the ConvTranspose layout [C_in][C_out][1]
      DTTS_TRY(tc_pack(h, &wk, 1, nullptr, D, H, 1, 1, 0, &h->t_s2pa_kT, s));
    }
    DTTS_TRY(tc_pack1(h, a + ".q_transform", false, H, H, 1, 0, &h->t_s2pa_q, s));
    DTTS_TRY(tc_pack1(h, a + ".v_transform", false, H, D, 1, 0, &h->t_s2pa_v, s));
    DTTS_TRY(tc_pack1(h, a + ".output_transform", false, H, H, 1, 0, &h->t_s2pa_o, s));
    if (h->precision) {
      const float* wq = h->tab.get(a + ".q_transform.weight", (uint64_t)H * H);
      const float* wk = h->tab.get(a + ".k_transform.weight", (uint64_t)H * D);
      const float* wv = h->tab.get(a + ".v_transform.weight", (uint64_t)H * D);
      const float* wo = h->tab.get(a + ".output_transform.weight", (uint64_t)H * H);
      float* wqk = h->pool.take((size_t)D * H);
      float* wvo = h->pool.take((size_t)H * D);
      if (!wq || !wk || !wv || !wo) return DTTS_ERR_MISSING_WEIGHT;
      if (!wqk || !wvo) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
      DTTS_CUDA(matmul_f32(wk, wq, wqk, D, H, H, 1, s));        // [D][H] = W_k^T [D][H'] W_q [H'][H]
      DTTS_CUDA(matmul_f32(wo, wv, wvo, H, D, H, 0, s));        // [H][D] = W_o [H][H'] W_v [H'][D]
      const float* c1 = wqk;
      const float* c2 = wvo;
      DTTS_TRY(tc_pack(h, &c1, 1, nullptr, D, H, 1, 0, 0, &h->t_s2pa_qk, s));
      DTTS_TRY(tc_pack(h, &c2, 1, nullptr, H, D, 1, 0, 0, &h->t_s2pa_vo, s));
    }
    if (d->s2pa_route == 1) {
      const float* wkv[2] = {h->tab.get(a + ".k_transform.weight", (uint64_t)H * D),
                             h->tab.get(a + ".v_transform.weight", (uint64_t)H * D)};
      if (!wkv[0] || !wkv[1]) return DTTS_ERR_MISSING_WEIGHT;
      DTTS_TRY(tc_pack(h, wkv, 2, nullptr, H, D, 1, 0, H, &h->t_s2pa_kv, s));      // N = H: one block per projection
    }
    DTTS_TRY(pack_dur_predictor(h, H, s));
    h->tc_text_end = h->tc_used;
    DTTS_TRY(pack_decoder(h, s));
    return DTTS_OK;
  };
  rc = build();
  if (rc == DTTS_OK) {
    cudaError_t e = cudaMalloc((void**)&h->bank_err_dev, sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(h->bank_err_dev, 0, sizeof(int), s);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->bank_err_host, sizeof(int));
    if (e == cudaSuccess) { *h->bank_err_host = 0; e = cudaEventCreateWithFlags(&h->bank_err_evt, cudaEventDisableTiming); }
    if (e != cudaSuccess) rc = fail(DTTS_ERR_CUDA, std::string("acoustic create (status words): ") + cudaGetErrorString(e));
  }
  if (rc == DTTS_OK) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = fail(DTTS_ERR_CUDA, std::string("acoustic create: ") + cudaGetErrorString(e));
  }
  if (rc != DTTS_OK) {
    dtts_acoustic_destroy(h);
    return rc;
  }
  *out = h;
  return DTTS_OK;
}

extern "C" int dtts_acoustic_destroy(dtts_acoustic* h) {
  if (!h) return DTTS_OK;
  h->pool.release();
  destroy_ps(h);
  if (h->tc_pool) cudaFree(h->tc_pool);
  if (h->flow_fused.stream) cudaFree(h->flow_fused.stream);
  if (h->bank_err_dev) cudaFree(h->bank_err_dev);
  if (h->bank_err_host) cudaFreeHost(h->bank_err_host);
  if (h->bank_err_evt) cudaEventDestroy(h->bank_err_evt);
  delete h;
  return DTTS_OK;
}

// Folds a landed gather status into the sticky word and reports it (0 = clean).
static int bank_status(dtts_acoustic* h) {
  if (h->bank_err_host && *h->bank_err_host) {
    h->bank_err_sticky |= *h->bank_err_host;
    *h->bank_err_host = 0;
  }
  const int st = h->bank_err_sticky;
  if (!st) return DTTS_OK;
  h->bank_err_sticky = 0;
  if (st & 1) return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode_bank: a dict_id was >= bank.n_entries (that character was encoded as an all-zero row)");
  return fail(DTTS_ERR_BAD_SHAPE, "dtts_text_encode_bank: a bank entry is longer than the Lk / Lp of the call (it was truncated)");
}

extern "C" int dtts_acoustic_status(dtts_acoustic* h, void* stream, int32_t sync) {
  if (!h) return fail(DTTS_ERR_BAD_ARG, "dtts_acoustic_status: null handle");
  if (sync) DTTS_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  else if (h->bank_err_evt && cudaEventQuery(h->bank_err_evt) == cudaErrorNotReady) return DTTS_OK;   // nothing landed yet
  return bank_status(h);
}

extern "C" uint64_t dtts_acoustic_launch_count(const dtts_acoustic* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------------------------
extern "C" uint64_t dtts_text_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tw, int32_t Lk, int32_t Lp) {
  if (!h || B <= 0 || Tw <= 0 || Lk <= 0 || Lp <= 0) return 0;
  const size_t H = h->d.hidden, F = h->d.ffn_filter, D = h->d.dict_dim, C = h->d.dur_chans;
  const size_t bt = (size_t)B * Tw;
  size_t n = 0;
  auto add = [&](size_t floats) { n += ws_round(floats * sizeof(float)); };
  add(bt * H); add(bt * H); add(bt * 3 * H); add(bt * H); add(bt * F);        // x, h, qkv, att, ffn
  add(bt); add(bt); add(bt); add(B); add(64);                                  // masks, lens, maxes
  add(bt * H); add(bt * D); add(bt * Lk); add(bt * D); add(bt * H); add(bt * H);  // q, qk, weights, ctx, ctxv, context
  add(bt * H); add(bt * C); add(bt * C);                                        // dur_in, d1, d2
  if (h->precision) n += 4 * ws_round((size_t)B * (F > D ? F : D) * tc_rows(Tw) * sizeof(tc16));   // 2 operand-plane sets
  if (h->d.s2pa_route == 1) {                                                  // gloss-token planes (hi, lo) + k|v
    n += 2 * ws_round((size_t)B * D * tc_rows(Tw * Lk) * sizeof(tc16));
    add(bt * Lk * 2 * H);
  }
  return n + 4096;
}

// Shared body of dtts_text_encode / dtts_text_encode_bank.  row_off / row_len != null: in->keys_dev / values_dev are the
// dictionary bank [rows][dict_dim] and character (b,t) owns rows [row_off, row_off + row_len) of it.
static int text_encode_impl(dtts_acoustic* h, const dtts_text_in* in, const dtts_text_out* out, void* ws,
                            uint64_t ws_bytes, void* stream, const int64_t* row_off, const int32_t* row_len) {
  if (!h || !in || !out || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode: null argument");
  if (h->d.model != DTTS_MODEL_DICT)
    return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode: this handle holds a PortaSpeech model (use dtts_ps_text_encode)");
  const int B = in->B, Tw = in->Tw, Lk = in->Lk, Lp = in->Lp;
  if (B <= 0 || Tw <= 0 || Lk <= 0 || Lp <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_text_encode: empty shape");
  if (!in->word_tokens_dev || !in->keys_dev || !in->values_dev || !in->key_map_dev || !in->pinyin_dev ||
      !in->pinyin_map_dev || !out->word_encoder_out_dev || !out->dict_attn_dev || !out->pron_attn_dev ||
      !out->dur_dev || !out->dur_int_dev || !out->ilens_dev)
    return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode: null tensor");
  if (((uintptr_t)in->keys_dev & 15) || ((uintptr_t)in->values_dev & 15))
    return fail(DTTS_ERR_ALIGNMENT, "dtts_text_encode: keys/values must be 16-byte aligned");
  if (ws_bytes < dtts_text_workspace_bytes(h, B, Tw, Lk, Lp))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_text_encode: workspace too small");
  const dtts_acoustic_desc& d = h->d;
  const int H = d.hidden, F = d.ffn_filter, D = d.dict_dim, C = d.dur_chans;
  const size_t bt = (size_t)B * Tw;
  Bump bump(ws, ws_bytes);
  float* x = bump.take<float>(bt * H);
  float* hb = bump.take<float>(bt * H);
  float* qkv = bump.take<float>(bt * 3 * H);
  float* att = bump.take<float>(bt * H);
  float* ffn = bump.take<float>(bt * F);
  float* seq_mask = bump.take<float>(bt);
  float* tok_mask = bump.take<float>(bt);
  float* keep = bump.take<float>(bt);
  int* lens = bump.take<int>(B);
  int* maxes = bump.take<int>(64);
  float* q = bump.take<float>(bt * H);
  float* qk = bump.take<float>(bt * D);
  float* weights = bump.take<float>(bt * Lk);
  float* ctx = bump.take<float>(bt * D);
  float* ctxv = bump.take<float>(bt * H);
  float* context = bump.take<float>(bt * H);
  float* dur_in = bump.take<float>(bt * H);
  float* d1 = bump.take<float>(bt * C);
  float* d2 = bump.take<float>(bt * C);
  TcRun tcr{};
  TcRun* tc = nullptr;
  if (h->precision) {
    tcr.B = B;
    tcr.take(bump, 2, (size_t)B * (F > D ? F : D) * tc_rows(Tw));
    tc = &tcr;
  }
  const bool gemm_route = d.s2pa_route == 1;
  if (gemm_route && row_off)
    return fail(DTTS_ERR_BAD_ARG, "s2pa_route = 1 (K/V projection GEMM) is not available with the dictionary bank");
  Planes KP;                                                   // gloss tokens as operand planes: C = D, T = Tw * Lk
  float* kv = nullptr;
  if (gemm_route) {
    KP.cap = (size_t)B * D * tc_rows(Tw * Lk);
    KP.hi = bump.take<tc16>(KP.cap);
    KP.lo = bump.take<tc16>(KP.cap);
    kv = bump.take<float>(bt * Lk * 2 * H);
  }
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_text_encode: workspace too small");
  Launcher L;
  L.stream = (cudaStream_t)stream;
  L.counter = &h->launches;
  cudaStream_t s = L.stream;
  tcr.h = h; tcr.L = &L; tcr.B = B;

  if (tc && ac_fuse_enabled()) L(l2_prefetch(h->tc_pool, h->tc_text_end * sizeof(tc16), s));   // weights -> L2 (kernels.cuh)
  // word embedding * sqrt(H), masks (dict_encoder.py:131-136)
  L(embed_tokens(in->word_tokens_dev, h->word_emb, sqrtf((float)H), B, Tw, H, d.word_size, x, seq_mask, tok_mask, lens,
                 s));
  run_encoder(h, h->sem, x, hb, qkv, att, ffn, seq_mask, B, Tw, L, tc);       // semantic encoder -> hb
  if (gemm_route) {
    // S2PA as written (dict_encoder.py:40-58): q = W_q x * D^-1/2; k = W_k keys and v = W_v values for every gloss
    // token -- the [B*Tw*Lk, D] x [D, 2H] "dict-attention GEMM" on tcgen05 -- then scores / softmax / weighted sum.
    const int TL = Tw * Lk;
    TcRun::Epi eq;
    eq.alpha = 1.f / sqrtf((float)D);
    tc->conv_nct(tc->P[0], h->t_s2pa_q, q, Tw, 1, 0, eq);                         // P[0] = last LN of the encoder
    const long gbs = (long)TL * D;                                                // gloss row (b, t*Lk + l) = D floats
    tc->stage(KP, in->keys_dev, gbs, 1, D, D, TL);
    if (in->values_dev == in->keys_dev) {
      tc->conv(KP, h->t_s2pa_kv, 0, 2, kv, (long)2 * H * TL, TL, 1, TL, 1, 0, TcRun::Epi());
    } else {
      tc->conv(KP, h->t_s2pa_kv, 0, 1, kv, (long)2 * H * TL, TL, 1, TL, 1, 0, TcRun::Epi());
      tc->stage(KP, in->values_dev, gbs, 1, D, D, TL);
      tc->conv(KP, h->t_s2pa_kv, 1, 1, kv + (size_t)H * TL, (long)2 * H * TL, TL, 1, TL, 1, 0, TcRun::Epi());
    }
    L(s2pa_attend(kv, q, in->key_map_dev, B, Tw, Lk, H, weights, out->dict_attn_dev, ctxv, s));
    tc->stage_nct(tc->P[0], ctxv, H, Tw);
    tc->conv_nct(tc->P[0], h->t_s2pa_o, context, Tw, 1, 0, TcRun::Epi());
  } else {
  // S2PA (dict_encoder.py:32-66), folded: logits = keys . (W_k^T (W_q x) * D^-1/2)
  if (tc) {
    TcRun::Epi ek;
    ek.alpha = 1.f / sqrtf((float)D);
    if (h->t_s2pa_qk.w && ac_fuse_enabled()) {
      tc->conv_nct(tc->P[0], h->t_s2pa_qk, qk, Tw, 1, 0, ek);                                 // P[0] = last LN of the encoder
    } else {
      tc->conv_nct(tc->P[0], h->t_s2pa_q, nullptr, Tw, 1, 0, TcRun::Epi(), 0, 0, &tc->P[1]);
      tc->conv_nct(tc->P[1], h->t_s2pa_kT, qk, Tw, 1, 0, ek);
    }
  } else {
    L(launch_conv1d_f32(conv_params(hb, Tw, h->s2pa_q, 0, H, q, Tw, 1, 1, 0), B, s));
    ConvParams p = conv_params(q, Tw, h->s2pa_kT, 0, D, qk, Tw, 1, 1, 0);
    p.alpha = 1.f / sqrtf((float)D);
    L(launch_conv1d_f32(p, B, s));
  }
  L(s2pa_stream(in->keys_dev, in->values_dev, in->key_map_dev, qk, B, Tw, Lk, D, weights, out->dict_attn_dev, ctx, s,
                row_off, row_len));
  if (tc) {
    tc->stage_nct(tc->P[0], ctx, D, Tw);
    if (h->t_s2pa_vo.w && ac_fuse_enabled()) {
      tc->conv_nct(tc->P[0], h->t_s2pa_vo, context, Tw, 1, 0, TcRun::Epi());
    } else {
      tc->conv_nct(tc->P[0], h->t_s2pa_v, nullptr, Tw, 1, 0, TcRun::Epi(), 0, 0, &tc->P[1]);
      tc->conv_nct(tc->P[1], h->t_s2pa_o, context, Tw, 1, 0, TcRun::Epi());
    }
  } else {
    L(launch_conv1d_f32(conv_params(ctx, Tw, h->s2pa_v, 0, H, ctxv, Tw, 1, 1, 0), B, s));
    L(launch_conv1d_f32(conv_params(ctxv, Tw, h->s2pa_o, 0, H, context, Tw, 1, 1, 0), B, s));
  }
  }   // folded route
  L(dict_maxes(in->key_map_dev, bt * Lk, in->pinyin_map_dev, bt * Lp, maxes, s));
  // x2 = context * x_mask + pron   (written into x, the input of the linguistic encoder)
  L(s2pa_pron(weights, in->key_map_dev, in->pinyin_dev, in->pinyin_map_dev, in->pron_modified_dev, maxes,
              h->pinyin_emb, d.pinyin_size, context, seq_mask, B, Tw, Lk, Lp, H, d.language_zh, out->pron_attn_dev, x,
              s));
  run_encoder(h, h->lin, x, hb, qkv, att, ffn, seq_mask, B, Tw, L, tc);       // linguistic encoder -> hb
  // word_encoder_out = x^T * (tokens > 0); dur_input; src_padding  (dict_encoder.py:168-170, model.py:94-96,73)
  L(finish_text(hb, tok_mask, B, Tw, H, out->word_encoder_out_dev, dur_in, keep, s));
  L(count_keep(keep, B, Tw, out->ilens_dev, s));
  // duration predictor (portaspeech/model.py:58-66)
  run_dur_predictor(h, dur_in, keep, d1, d2, B, Tw, out->dur_dev, out->dur_int_dev, L, tc);
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_text_encode: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}

extern "C" int dtts_text_encode(dtts_acoustic* h, const dtts_text_in* in, const dtts_text_out* out, void* ws,
                                uint64_t ws_bytes, void* stream) {
  return text_encode_impl(h, in, out, ws, ws_bytes, stream, nullptr, nullptr);
}

// ---- GPU-resident dictionary bank: the per-batch dict_msg is a function of the character ids ----
static size_t bank_gather_bytes(int B, int Tw, int Lk, int Lp) {
  const size_t bt = (size_t)B * Tw;
  return ws_round(bt * Lk * sizeof(float)) + 2 * ws_round(bt * Lp * sizeof(int64_t)) + ws_round(bt * sizeof(int64_t)) +
         ws_round(bt * sizeof(int32_t)) + 2048;
}

extern "C" uint64_t dtts_text_bank_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tw, int32_t Lk,
                                                   int32_t Lp) {
  const uint64_t base = dtts_text_workspace_bytes(h, B, Tw, Lk, Lp);
  return base ? base + bank_gather_bytes(B, Tw, Lk, Lp) : 0;
}

extern "C" int dtts_text_encode_bank(dtts_acoustic* h, const dtts_dict_bank* bank, const dtts_text_in_bank* in,
                                     const dtts_text_out* out, void* ws, uint64_t ws_bytes, void* stream) {
  if (!h || !bank || !in || !out || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode_bank: null argument");
  if (!bank->keys_dev || !bank->values_dev || !bank->key_map_dev || !bank->tok_offsets_dev || !bank->pinyin_dev ||
      !bank->pinyin_map_dev || !bank->pin_offsets_dev || bank->n_entries <= 0 || !in->dict_ids_dev ||
      !in->word_tokens_dev)
    return fail(DTTS_ERR_BAD_ARG, "dtts_text_encode_bank: null bank / id tensor");
  const int B = in->B, Tw = in->Tw, Lk = in->Lk, Lp = in->Lp;
  if (B <= 0 || Tw <= 0 || Lk <= 0 || Lp <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_text_encode_bank: empty shape");
  if (ws_bytes < dtts_text_bank_workspace_bytes(h, B, Tw, Lk, Lp))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_text_encode_bank: workspace too small");
  const size_t gbytes = bank_gather_bytes(B, Tw, Lk, Lp);
  const size_t bt = (size_t)B * Tw;
  Bump bump(ws, gbytes);
  float* key_map = bump.take<float>(bt * Lk);
  int64_t* pinyin = bump.take<int64_t>(bt * Lp);
  int64_t* pinyin_map = bump.take<int64_t>(bt * Lp);
  int64_t* row_off = bump.take<int64_t>(bt);
  int32_t* row_len = bump.take<int32_t>(bt);
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_text_encode_bank: workspace too small");
  // Status word: cleared, OR-ed by the kernel (1: id >= n_entries, 2: entry wider than Lk / Lp) and copied to the
  // handle's pinned word on the same stream.  An earlier call's status that has landed but was never collected stays
  // sticky; it is reported by dtts_acoustic_status and by dtts_length_regulate_scan (the path's one synchronising call).
  cudaStream_t gs = (cudaStream_t)stream;
  if (h->bank_err_host && *h->bank_err_host && cudaEventQuery(h->bank_err_evt) == cudaSuccess) {
    h->bank_err_sticky |= *h->bank_err_host;
    *h->bank_err_host = 0;
  }
  DTTS_CUDA(cudaMemsetAsync(h->bank_err_dev, 0, sizeof(int), gs));
  DTTS_CUDA(dict_bank_gather(in->dict_ids_dev, bank->tok_offsets_dev, bank->pin_offsets_dev, bank->key_map_dev,
                             bank->pinyin_dev, bank->pinyin_map_dev, bank->n_entries, B, Tw, Lk, Lp, key_map, pinyin,
                             pinyin_map, row_off, row_len, h->bank_err_dev, gs));
  DTTS_CUDA(cudaMemcpyAsync(h->bank_err_host, h->bank_err_dev, sizeof(int), cudaMemcpyDeviceToHost, gs));
  DTTS_CUDA(cudaEventRecord(h->bank_err_evt, gs));
  h->launches++;
  dtts_text_in tin{};
  tin.word_tokens_dev = in->word_tokens_dev;
  tin.pron_modified_dev = in->pron_modified_dev;
  tin.keys_dev = bank->keys_dev;
  tin.values_dev = bank->values_dev;
  tin.key_map_dev = key_map;
  tin.pinyin_dev = pinyin;
  tin.pinyin_map_dev = pinyin_map;
  tin.B = B; tin.Tw = Tw; tin.Lk = Lk; tin.Lp = Lp;
  return text_encode_impl(h, &tin, out, (char*)ws + gbytes, ws_bytes - gbytes, stream, row_off, row_len);
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" int dtts_length_regulate_scan(dtts_acoustic* h, const int64_t* dur_int, const int64_t* ilens, int32_t B,
                                         int32_t Tw, int32_t* cum, int32_t* totals, int32_t* t_raw_host,
                                         void* stream) {
  if (!h || !dur_int || !ilens || !cum || !totals || !t_raw_host)
    return fail(DTTS_ERR_BAD_ARG, "dtts_length_regulate_scan: null argument");
  if (B <= 0 || Tw <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_length_regulate_scan: empty shape");
  cudaStream_t s = (cudaStream_t)stream;
  // totals has B+1 entries: the last one receives the batch maximum
  DTTS_CUDA(cudaMemsetAsync(totals + B, 0, sizeof(int32_t), s));
  DTTS_CUDA(lr_scan(dur_int, ilens, B, Tw, cum, totals, totals + B, s));
  h->launches++;
  DTTS_CUDA(cudaMemcpyAsync(t_raw_host, totals + B, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DTTS_CUDA(cudaStreamSynchronize(s));
  return bank_status(h);                     // a bad dictionary id / width of the text stage surfaces at this sync
}

extern "C" int dtts_length_regulate_fill(dtts_acoustic* h, const int32_t* cum, const int64_t* ilens, int32_t B,
                                         int32_t Tw, int32_t t_raw, int32_t T, int64_t* mel2word, void* stream) {
  if (!h || !cum || !ilens || !mel2word) return fail(DTTS_ERR_BAD_ARG, "dtts_length_regulate_fill: null argument");
  if (B <= 0 || Tw <= 0 || T <= 0 || t_raw <= 0 || t_raw > T)
    return fail(DTTS_ERR_BAD_SHAPE, "dtts_length_regulate_fill: need 0 < t_raw <= T");
  DTTS_CUDA(lr_fill(cum, ilens, B, Tw, t_raw, T, mel2word, (cudaStream_t)stream));
  h->launches++;
  return DTTS_OK;
}

extern "C" int dtts_expand(dtts_acoustic* h, const float* enc, const int64_t* mel2word, int32_t B, int32_t Tw,
                           int32_t T, float* decoder_inp, float* g_bct, float* x_mask, void* stream) {
  if (!h || !enc || !mel2word || !decoder_inp || !g_bct || !x_mask)
    return fail(DTTS_ERR_BAD_ARG, "dtts_expand: null argument");
  if (B <= 0 || Tw <= 0 || T <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_expand: empty shape");
  DTTS_CUDA(lr_gather(enc, mel2word, B, Tw, T, h->d.hidden, decoder_inp, g_bct, x_mask, (cudaStream_t)stream));
  h->launches++;
  return DTTS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" uint64_t dtts_decode_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  const size_t H = h->d.hidden, FH = h->d.flow_hidden, T4 = T / 4;
  size_t n = 0;
  auto add = [&](size_t floats) { n += ws_round(floats * sizeof(float)); };
  add(B * H * T4);                                     // g_sqz
  add(B * 2 * FH * h->d.flow_layers * T4);             // flow cond
  add(B * FH * T4); add(B * 2 * FH * T4); add(B * FH * T4); add(B * FH * T4);   // h, a, acts, skip
  add(B * H * T);                                      // x
  add(B * 2 * H * h->d.dec_layers * T);                // cond
  add(B * 2 * H * T); add(B * H * T); add(B * H * T);  // a, acts, skip
  if (h->precision)                                                                    // 3 operand-plane sets
    n += 6 * ws_round((size_t)B * H * (size_t)std::max(tc_rows(T), 4 * tc_rows(T / 4)) * sizeof(tc16));
  return n + 4096;
}

extern "C" int dtts_decode_mel(dtts_acoustic* h, const float* g, const float* z_in, int32_t B, int32_t T, float* mel,
                               float* z_p, void* ws, uint64_t ws_bytes, void* stream) {
  if (!h || !g || !z_in || !mel || !z_p || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_decode_mel: null argument");
  if (B <= 0 || T <= 0 || T % h->d.frames_multiple)
    return fail(DTTS_ERR_BAD_SHAPE, "dtts_decode_mel: T must be a positive multiple of frames_multiple");
  if (ws_bytes < dtts_decode_workspace_bytes(h, B, T))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_decode_mel: workspace too small");
  const dtts_acoustic_desc& d = h->d;
  const int H = d.hidden, FH = d.flow_hidden, T4 = T / 4, half = d.latent / 2;
  Bump bump(ws, ws_bytes);
  float* g_sqz = bump.take<float>((size_t)B * H * T4);
  float* fcond = bump.take<float>((size_t)B * 2 * FH * d.flow_layers * T4);
  float* fh = bump.take<float>((size_t)B * FH * T4);
  float* fa = bump.take<float>((size_t)B * 2 * FH * T4);
  float* facts = bump.take<float>((size_t)B * FH * T4);
  float* fskip = bump.take<float>((size_t)B * FH * T4);
  float* x = bump.take<float>((size_t)B * H * T);
  float* cond = bump.take<float>((size_t)B * 2 * H * d.dec_layers * T);
  float* a = bump.take<float>((size_t)B * 2 * H * T);
  float* acts = bump.take<float>((size_t)B * H * T);
  float* skip = bump.take<float>((size_t)B * H * T);
  TcRun tcr{};
  TcRun* tc = nullptr;
  if (h->precision) {
    tcr.B = B;
    tcr.take(bump, 3, (size_t)B * H * (size_t)std::max(tc_rows(T), 4 * tc_rows(T4)));
    tc = &tcr;
  }
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_decode_mel: workspace too small");
  Launcher L;
  L.stream = (cudaStream_t)stream;
  L.counter = &h->launches;
  cudaStream_t s = L.stream;
  tcr.h = h; tcr.L = &L; tcr.B = B;

  const bool fuse_flow = tc && h->flow_fused.ready() && ac_fuse_enabled() && z_in != z_p;   // (its tiles read z_in as halo)
  if (tc && ac_fuse_enabled()) {               // decoder weights -> L2 while the first launches run (kernels.cuh)
    L(l2_prefetch(h->tc_pool + h->tc_text_end, (h->tc_used - h->tc_text_end) * sizeof(tc16), s));
    if (fuse_flow) L(l2_prefetch(h->flow_fused.stream, h->flow_fused.stream_bytes(), s));
  }
  // g_sqz = Conv1d(H,H,k=8,s=4,p=2)(g)  (fvae_semantics.py:93-94; semantics == 0)
  if (tc) {
    Planes& P0 = tc->P[0];
    if (tc->shape(P0, 4 * H, T4)) {
      // x'[(sp, ci), q] = g[ci, 4q + sp]: one staging launch for the four phases
      L(tc_to_planes_full(g, (long)H * T, T, 4, B, 4 * H, T4, 1.f, P0.hi, h->mode.a_planes == 2 ? P0.lo : nullptr, P0.rows,
                          TC_PADF, h->mode.fmt, s, 0, 0, H));
      // fused prior flow: g_sqz is only needed as the operand planes of the conditioning GEMMs
      if (fuse_flow) tc->conv_nct(P0, h->t_gpre, nullptr, T4, 1, 1, TcRun::Epi(), 0, 0, &tc->P[1]);
      else tc->conv_nct(P0, h->t_gpre, g_sqz, T4, 1, 1, TcRun::Epi());
    }
  } else {
    L(launch_conv1d_f32(conv_params(g, T, h->g_pre, 0, H, g_sqz, T4, 1, 4, 2), B, s));
  }
  // prior flow, reverse (glow_modules.py:108-128,157-163).  The channel Flip is folded into the pre/post weights:
  // on "odd" layers the conditioning half is physical channels [half, 2*half) and the updated half is [0, half).
  if (fuse_flow) {
    // every coupling layer in ONE launch (flow_fused.cu): activations stay in shared memory / TMEM
    const Planes& Pg = tc->P[1];
    FlowFusedParams fp{};
    fp.g_hi = Pg.hi; fp.g_lo = Pg.lo; fp.g_bs = (long)Pg.C * Pg.rows; fp.g_rows = Pg.rows; fp.g_pad = TC_PADF;
    fp.z_in = z_in; fp.z_out = z_p; fp.T = T4; fp.B = B;
    L(launch_flow_fused(h->flow_fused, fp, s));
  } else {
    L(copy_f32(z_in, z_p, (size_t)B * d.latent * T4, s));
  }
  for (int f = d.flow_blocks - 1; f >= 0 && !fuse_flow; --f) {
    const FlowW& F = h->flows[f];
    const int c_x0 = F.odd ? half : 0, c_x1 = F.odd ? 0 : half;
    if (FH % 8 == 0 && half % 8 == 0) {      // 8 <-> 64 channels: one thread per position (pointwise_small_kernel)
      L(pointwise_small(z_p + (size_t)c_x0 * T4, (long)d.latent * T4, F.pre.w, F.pre.bias, half, FH, B, T4, 1.f, nullptr,
                        0, fh, (long)FH * T4, s));
    } else {
      ConvParams p = conv_params(z_p + (size_t)c_x0 * T4, T4, F.pre, 0, FH, fh, T4, 1, 1, 0);
      p.x_bs = (long)d.latent * T4;
      L(launch_conv1d_f32(p, B, s));
    }
    run_wn(F.wn, FH, d.flow_kernel, fh, g_sqz, H, fcond, fa, facts, fskip, B, T4, L, tc);
    {
      // x1 = x1 - m   (mean_only, logs = 0)
      float* x1 = z_p + (size_t)c_x1 * T4;
      if (FH % 8 == 0 && half % 8 == 0) {
        L(pointwise_small(fskip, (long)FH * T4, F.post.w, F.post.bias, FH, half, B, T4, -1.f, x1, (long)d.latent * T4, x1,
                          (long)d.latent * T4, s));
      } else {
        ConvParams p = conv_params(fskip, T4, F.post, 0, half, x1, T4, 1, 1, 0);
        p.o_bs = (long)d.latent * T4;
        p.alpha = -1.f;
        p.res = x1; p.r_bs = (long)d.latent * T4; p.r_cs = T4; p.r_ts = 1;
        L(launch_conv1d_f32(p, B, s));
      }
    }
  }
  // decoder (fvae_semantics.py:53-58)
  const bool fuse_dec = tc && ac_fuse_enabled() && H % 8 == 0 && (d.latent == 16 || d.latent == 8) &&
                        (size_t)4 * d.latent * H * sizeof(float) <= 48 * 1024;     // pre_net weights in shared memory
  if (fuse_dec) {
    // ConvTranspose1d(latent -> H, k = 4, s = 4) straight into the fp32 stream AND the operand planes of the first
    // WaveNet convolution (one launch instead of the generic fp32 kernel + a staging pass)
    L(fvae_pre_net_planes(z_p, h->dec_pre.w, h->dec_pre.bias, B, d.latent, H, T4, x, tc->out_of(tc->P[1], H, T, true), s));
    run_wn(h->dec_wn, H, d.dec_kernel, x, g, H, cond, a, acts, skip, B, T, L, tc, true, &tc->P[0]);
  } else {
    L(launch_conv1d_f32(convT_params(z_p, T4, h->dec_pre, x, T, 4, 0), B, s));
    run_wn(h->dec_wn, H, d.dec_kernel, x, g, H, cond, a, acts, skip, B, T, L, tc);
  }
  if (tc) {
    TcRun::Epi eo;
    eo.c_valid = d.n_mel;
    if (!fuse_dec) tc->stage_nct(tc->P[0], skip, H, T);     // (fused: the last res_skip epilogue wrote the planes)
    tc->conv(tc->P[0], h->t_out, 0, 0, mel, (long)T * d.n_mel, 1, d.n_mel, T, 1, 0, eo);    // mel_out is [B,T,80]
  } else {
    ConvParams p = conv_params(skip, T, h->dec_out, 0, d.n_mel, mel, T, 1, 1, 0);
    p.o_bs = (long)T * d.n_mel; p.o_cs = 1; p.o_ts = d.n_mel;                 // mel_out is [B,T,80]
    L(launch_conv1d_f32(p, B, s));
  }
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_decode_mel: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}

extern "C" int dtts_debug_set_acoustic_fuse(int32_t mode) {
  if (mode < -1 || mode > 1) return fail(DTTS_ERR_BAD_ARG, "dtts_debug_set_acoustic_fuse: mode must be -1, 0 or 1");
  ac_fuse_override(mode);
  return DTTS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
extern "C" int dtts_debug_conv1d(const float* x, const float* w, const float* bias, float* out, int32_t B, int32_t C_in,
                                 int32_t T_in, int32_t C_out, int32_t K, int32_t stride, int32_t padding,
                                 int32_t dilation, int32_t transposed, float pre_slope, float* scratch_w,
                                 void* stream) {
  if (!x || !w || !out || !scratch_w) return fail(DTTS_ERR_BAD_ARG, "dtts_debug_conv1d: null argument");
  DTTS_TRY(arch_check());
  cudaStream_t s = (cudaStream_t)stream;
  ConvW cw;
  cw.bias = bias; cw.C_out = C_out; cw.C_in = C_in;
  ConvParams p;
  if (transposed) {
    if (K % stride) return fail(DTTS_ERR_BAD_SHAPE, "transposed conv needs K % stride == 0");
    DTTS_CUDA(repack_convT(w, scratch_w, C_in, C_out, K, stride, s));
    cw.w = scratch_w; cw.ktaps = K / stride; cw.phases = stride;
    const int T_out = (T_in - 1) * stride - 2 * padding + K;
    p = convT_params(x, T_in, cw, out, T_out, stride, padding);
  } else {
    DTTS_CUDA(repack_conv(w, scratch_w, C_out, C_in, K, 0, 0, s));
    cw.w = scratch_w; cw.ktaps = K; cw.phases = 1;
    const int T_out = (T_in + 2 * padding - dilation * (K - 1) - 1) / stride + 1;
    p = conv_params(x, T_in, cw, 0, C_out, out, T_out, dilation, stride, padding);
  }
  p.pre_slope = pre_slope;
  DTTS_CUDA(launch_conv1d_f32(p, B, s));
  return DTTS_OK;
}
