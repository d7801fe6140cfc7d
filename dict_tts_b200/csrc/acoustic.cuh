// Shared declarations of the acoustic-model handle (acoustic.cu: Dict-TTS text side + FVAE decoder; portaspeech.cu: the
// PortaSpeech sibling's text side, SURVEY.md §8f-3, which shares the decoder, the length regulator and the duration
// predictor).
#pragma once
#include <algorithm>

#include "engine.cuh"
#include "flow_fused.cuh"
#include "tc_conv.cuh"
#include "tc16.cuh"

namespace dtts {
namespace ac {

// Every dense convolution exists twice: packed for the fp32 FMA kernel (precision 0, the exact path) and as tcgen05
// blobs (precision 1: bf16 hi/lo on both operands, 3 MMAs per product, fp32-class accuracy -- durations must round the
// same way as the reference's fp32 forward).
struct EncLayerW {
  ConvW qkv, o, ffn1, ffn2;
  TcConvW t_qkv, t_o, t_ffn1, t_ffn2;
  const float *g1, *b1, *g2, *b2;
};
struct EncoderW {
  std::vector<EncLayerW> layers;
  const float *last_g, *last_b;
};
struct WNW {
  ConvW cond;
  std::vector<ConvW> in_layers, res_skip;
  TcConvW t_cond;
  std::vector<TcConvW> t_in, t_rs;      // t_rs blocks of `hidden` channels: [0] -> x update, [1] -> skip
  // the same in_layers / cond_layer with their output channels permuted so that every N block holds [N/2 tanh channels |
  // the N/2 sigmoid channels that gate them]: the gate then runs in the convolution epilogue (empty: not built)
  std::vector<TcConvW> t_in_g;
  TcConvW t_cond_g;
};
struct FlowW {
  ConvW pre, post;
  WNW wn;
  int odd;     // 1: the latent is logically channel-flipped while this coupling layer runs
};

struct PsW;     // PortaSpeech text-side weights (portaspeech.cu)

}  // namespace ac
}  // namespace dtts

struct dtts_acoustic {
  dtts_acoustic_desc d;
  dtts::WeightTable tab;
  dtts::Pool pool;
  const float* word_emb;
  const float* pinyin_emb;
  dtts::ac::EncoderW sem, lin;
  dtts::ConvW s2pa_q, s2pa_kT, s2pa_v, s2pa_o;
  dtts::TcConvW t_s2pa_q, t_s2pa_kT, t_s2pa_v, t_s2pa_o;
  // folded route with the projections that are applied back to back multiplied once at create time:
  // qk = (W_k^T W_q) x and context = (W_o W_v) ctx -- one tcgen05 launch each instead of two (empty: not built)
  dtts::TcConvW t_s2pa_qk, t_s2pa_vo;
  dtts::TcConvW t_s2pa_kv;              // s2pa_route = 1: [W_k ; W_v] side by side, dict_dim -> 2 * hidden (block 0 = k, block 1 = v)
  dtts::TcConvW t_gpre;                 // g_pre_net as a stride-1 k=3 convolution over the 4x space-to-depth input (C' = 4H)
  dtts::TcConvW t_out;                  // out_proj with C_out zero-padded to a multiple of 32
  int out_pad = 0;
  std::vector<dtts::ConvW> dur_conv;
  std::vector<dtts::TcConvW> t_dur;
  int precision = 0;              // 0: fp32 FMA pipe; 1: tcgen05 (bf16 hi/lo x hi/lo)
  dtts::TcMode mode;
  dtts::tc16* tc_pool = nullptr;
  size_t tc_cap = 0, tc_used = 0;
  size_t tc_text_end = 0;         // tc_pool[0, tc_text_end): text side + duration predictor; the rest: decoder (L2 prefetch ranges)
  std::vector<const float*> dur_ln_g, dur_ln_b;
  const float *dur_w, *dur_b;
  dtts::ConvW g_pre, dec_pre, dec_out;
  std::vector<dtts::ac::FlowW> flows;   // in reference order (flows.0, .2, .4, .6)
  dtts::ac::WNW dec_wn;
  dtts::FlowFusedW flow_fused;          // the whole prior flow as one launch (flow_fused.cu); empty: per-layer launches
  dtts::ac::PsW* ps = nullptr;     // model = 1 (PortaSpeech sibling): its text-side weights; the dict-encoder members stay empty
  uint64_t launches = 0;
  // dictionary-bank gather status (dtts_text_encode_bank): device word written by dict_bank_gather_kernel, copied to the
  // pinned host word after every gather; sticky until dtts_acoustic_status reports it
  int* bank_err_dev = nullptr;
  int* bank_err_host = nullptr;
  cudaEvent_t bank_err_evt = nullptr;
  int bank_err_sticky = 0;
};


namespace dtts {
namespace ac {

// conv weight [C_out][C_in][K] (+ optional bias) -> packed for the fp32 kernel
int pack(dtts_acoustic* h, const std::string& name, int C_out, int C_in, int K, bool has_bias, ConvW* cw,
         cudaStream_t s, int rci = 0, int rco = 0, const char* wsuffix = ".weight");
int pick_n(int C_out);
// Packs `parts` convolutions that share (C_in, K) side by side along C_out into one tensor-core weight set
int tc_pack(dtts_acoustic* h, const float* const* w, int parts, const float* bias, int C_out_part, int C_in, int K,
            int transposed, int N, TcConvW* cw, cudaStream_t s);
int tc_pack1(dtts_acoustic* h, const std::string& name, bool has_bias, int C_out, int C_in, int K, int N, TcConvW* cw,
             cudaStream_t s, int transposed = 0);

// One set of operand planes in the caller's workspace; (C, T, rows) describe what it currently holds.
struct Planes {
  tc16 *hi = nullptr, *lo = nullptr;
  size_t cap = 0;                 // elements per plane
  int C = 0, T = 0, rows = 0;
};

// Per-call context of the tensor-core convolutions: operand-plane sets (inputs are written there by the producing
// kernel -- LayerNorm, attention, gate, a convolution epilogue -- or converted from fp32 by stage()) and launch glue.
struct TcRun {
  dtts_acoustic* h;
  Launcher* L;
  int B;
  Planes P[3];
  struct Epi {
    const float* res = nullptr; long r_bs = 0, r_cs = 0, r_ts = 1;
    const float* mask = nullptr; int m_bs = 0;
    int act = 0; float alpha = 1.f, post = 1.f; int accumulate = 0;
    int c_valid = 0;              // > 0: number of real output channels (the rest is zero padding)
    int gate = 0;                 // 1: WaveNet gate in the epilogue (TcConvParams::gate): `po` receives C_out / 2 channels
    // channel LayerNorm of the output in the epilogue (TcConvParams::ln_*): `po` receives LN(out) instead of out
    const float* ln_gamma = nullptr; const float* ln_beta = nullptr; float ln_eps = 0.f;
    const float* ln_in_mask = nullptr; const float* ln_out_mask = nullptr; float* ln_y = nullptr;
  };
  // can conv(w) carry a fused LayerNorm over its C_out channels for sequences of T steps?
  static bool ln_fusable(const TcConvW& w, int T) { return w.C_out == w.N && w.N <= 256 && !w.pair && !w.stack && T <= 128; }
  bool shape(Planes& p, int C, int T) {
    p.C = C; p.T = T; p.rows = tc_rows(T);
    if ((size_t)B * C * p.rows > p.cap) { (*L)(cudaErrorInvalidValue); return false; }
    return true;
  }
  // destination descriptor for a producer kernel
  PlaneOut out_of(Planes& p, int C, int T, bool zero_halo) {
    PlaneOut o;
    if (!shape(p, C, T)) return o;
    o.hi = p.hi; o.lo = h->mode.a_planes == 2 ? p.lo : nullptr;
    o.rows = p.rows; o.pad = TC_PADF; o.fmt = h->mode.fmt; o.zero_halo = zero_halo ? 1 : 0;
    return o;
  }
  // fp32 x (element (c,t) of batch b at x[b*bs + c*cs + t*ts]) -> planes, halo rows zeroed
  void stage(Planes& p, const float* x, long bs, long cs, long ts, int C, int T) {
    if (!shape(p, C, T)) return;
    (*L)(tc_to_planes_full(x, bs, cs, ts, B, C, T, 1.f, p.hi, h->mode.a_planes == 2 ? p.lo : nullptr, p.rows, TC_PADF,
                           h->mode.fmt, L->stream));
  }
  void stage_nct(Planes& p, const float* x, int C, int T) { stage(p, x, (long)C * T, T, 1, C, T); }
  // channels [c_off, c_off + C) of planes already shaped to (C_total, T)
  void stage_sub(Planes& p, const float* x, long bs, long cs, long ts, int C, int c_off) {
    (*L)(tc_to_planes_full(x, bs, cs, ts, B, C, p.T, 1.f, p.hi, h->mode.a_planes == 2 ? p.lo : nullptr, p.rows, TC_PADF,
                           h->mode.fmt, L->stream, p.C, c_off));
  }
  // blocks [blk0, blk0+nblk) of w applied to `in`; fp32 result (optional) element (c,t) at out[b*o_bs + c*o_cs + t*o_ts]
  // with c counted from the first block; `po` (optional) receives the result as operand planes of the next convolution.
  void conv(const Planes& in, const TcConvW& w, int blk0, int nblk, float* out, long o_bs, long o_cs, long o_ts,
            int T_out, int dil, int pad, const Epi& e, Planes* po = nullptr) {
    if (w.C_in != in.C) { (*L)(cudaErrorInvalidValue); return; }
    TcConvW sub = w;
    const int nblocks = w.C_out / w.N;
    if (nblk <= 0) nblk = nblocks - blk0;
    sub.C_out = nblk * w.N;
    sub.w = w.w + (size_t)blk0 * (w.elems() / nblocks);
    sub.bias = w.bias ? w.bias + (size_t)blk0 * w.N : nullptr;
    TcConvParams p{};
    p.a_hi = in.hi; p.a_lo = h->mode.a_planes == 2 ? in.lo : nullptr;
    p.a_bs = (long)in.C * in.rows; p.a_rows = in.rows; p.a_pad = TC_PADF;
    p.tap_off0 = -pad; p.tap_step = dil;
    tc_conv_plan(&p, sub, T_out, h->mode.a_planes);
    p.ot_mul = 1; p.ot_add = 0; p.T_out = T_out;
    p.o32 = out; p.o32_bs = o_bs; p.o_nct = 1; p.o_cs = o_cs; p.o_ts = o_ts;
    p.res = e.res; p.r_bs = e.r_bs; p.r_cs = e.r_cs; p.r_ts = e.r_ts;
    p.mask = e.mask; p.m_bs = e.m_bs; p.act = e.act; p.alpha = e.alpha; p.post = e.post; p.accumulate = e.accumulate;
    p.slope = 1.f;
    if (e.c_valid > 0) p.c_valid = e.c_valid;
    p.gate = e.gate;
    p.ln_gamma = e.ln_gamma; p.ln_beta = e.ln_beta; p.ln_eps = e.ln_eps;
    p.ln_in_mask = e.ln_in_mask; p.ln_out_mask = e.ln_out_mask; p.ln_y = e.ln_y;
    if (e.ln_gamma && !e.mask) p.m_bs = T_out;       // the LN masks are [B][T] like the convolution's
    if (po) {
      if (!shape(*po, e.gate ? sub.C_out / 2 : sub.C_out, T_out)) return;
      p.o_hi = po->hi; p.o_lo = h->mode.a_planes == 2 ? po->lo : nullptr;
      p.op_bs = (long)po->C * po->rows; p.op_rows = po->rows; p.op_pad = TC_PADF;
    }
    (*L)(launch_tc_conv(p, B, L->stream));
  }
  void conv_nct(const Planes& in, const TcConvW& w, float* out, int T_out, int dil, int pad, const Epi& e, int blk0 = 0,
                int nblk = 0, Planes* po = nullptr) {
    const int nblocks = w.C_out / w.N;
    const int co = (nblk > 0 ? nblk : nblocks - blk0) * w.N;
    conv(in, w, blk0, nblk, out, (long)co * T_out, T_out, 1, T_out, dil, pad, e, po);
  }
  // carve `n` plane sets of `cap` elements per plane out of the workspace
  void take(Bump& bump, int n, size_t cap) {
    for (int i = 0; i < n; ++i) {
      P[i].cap = cap;
      P[i].hi = bump.take<tc16>(cap);
      P[i].lo = bump.take<tc16>(cap);
    }
  }
};


// q | k | v of one attention layer side by side ([H][1][3H] for the fp32 kernel, three N-blocks for tcgen05); names[j] are
// the three weight keys (bias keys optional: bias_names == nullptr -> no bias)
int pack_qkv(dtts_acoustic* h, const std::string names[3], const std::string* bias_names, ConvW* qkv, TcConvW* t_qkv,
             cudaStream_t s);
int pack_encoder(dtts_acoustic* h, const std::string& p, EncoderW* e, cudaStream_t s);
int pack_wn(dtts_acoustic* h, const std::string& p, int hidden, int n_layers, int K, int gin, WNW* wn, cudaStream_t s,
            bool gated = false);
int pack_dur_predictor(dtts_acoustic* h, int c_in, cudaStream_t s);
int pack_decoder(dtts_acoustic* h, cudaStream_t s);
// DurationPredictor.forward (portaspeech/model.py:58-66) + softplus head: dur_in [B,H,Tw] channel-first, keep [B,Tw];
// d1 / d2: [B,dur_chans,Tw] scratch.  Writes dur [B,Tw] and (optional) dur_int.
void run_dur_predictor(dtts_acoustic* h, const float* dur_in, const float* keep, float* d1, float* d2, int B, int Tw,
                       float* dur, int64_t* dur_int, Launcher& L, TcRun* tc);
int destroy_ps(dtts_acoustic* h);
int create_ps(dtts_acoustic* h, cudaStream_t s);

}  // namespace ac
}  // namespace dtts
