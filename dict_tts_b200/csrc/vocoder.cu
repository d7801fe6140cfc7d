// HiFi-GAN V1 generator (modules/hifigan/hifigan.py:27-58,101-142) behind dtts_vocode.
// fp32 path: every convolution is one launch of conv1d_f32 with the leaky-ReLU fused into the operand load and the
// residual add / 1/3 resblock mean / tanh fused into the epilogue (the reference launches 78 cuDNN convolutions
// plus 127 element-wise kernels per call).
#include "engine.cuh"

using namespace dtts;

struct dtts_vocoder {
  dtts_vocoder_desc desc;
  WeightTable tab;
  Pool pool;
  ConvW conv_pre, conv_post;
  std::vector<ConvW> ups;
  std::vector<ConvW> rb1, rb2;   // [stage][rb][m] flattened
  uint64_t launches = 0;
  int hop = 1;
  size_t unit = 0;               // max over stages of C*T_len per input frame
};

namespace {

int pack_conv(dtts_vocoder* h, const std::string& name, int C_out, int C_in, int K, ConvW* cw, cudaStream_t s) {
  const float* w = h->tab.get(name + ".weight", (uint64_t)C_out * C_in * K);
  if (!w) return DTTS_ERR_MISSING_WEIGHT;
  const float* b = h->tab.get(name + ".bias", C_out);
  if (!b) return DTTS_ERR_MISSING_WEIGHT;
  float* dst = h->pool.take((size_t)C_out * C_in * K);
  if (!dst) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
  DTTS_CUDA(repack_conv(w, dst, C_out, C_in, K, 0, 0, s));
  cw->w = dst; cw->bias = b; cw->C_out = C_out; cw->C_in = C_in; cw->ktaps = K; cw->phases = 1;
  return DTTS_OK;
}

int pack_convT(dtts_vocoder* h, const std::string& name, int C_in, int C_out, int K, int S, ConvW* cw,
               cudaStream_t s) {
  const float* w = h->tab.get(name + ".weight", (uint64_t)C_out * C_in * K);
  if (!w) return DTTS_ERR_MISSING_WEIGHT;
  const float* b = h->tab.get(name + ".bias", C_out);
  if (!b) return DTTS_ERR_MISSING_WEIGHT;
  float* dst = h->pool.take((size_t)C_out * C_in * K);
  if (!dst) return fail(DTTS_ERR_CUDA, "weight pool exhausted");
  DTTS_CUDA(repack_convT(w, dst, C_in, C_out, K, S, s));
  cw->w = dst; cw->bias = b; cw->C_out = C_out; cw->C_in = C_in; cw->ktaps = K / S; cw->phases = S;
  return DTTS_OK;
}

}  // namespace

extern "C" int dtts_vocoder_create(const dtts_vocoder_desc* d, const float* arena_dev, uint64_t arena_floats,
                                   const dtts_weight_entry* table, int32_t n_entries, void* stream,
                                   dtts_vocoder** out) {
  if (!d || !out) return fail(DTTS_ERR_BAD_ARG, "null descriptor/out");
  *out = nullptr;
  if (d->n_ups < 1 || d->n_ups > DTTS_MAX_UPS || d->n_rb < 1 || d->n_rb > DTTS_MAX_RB)
    return fail(DTTS_ERR_BAD_SHAPE, "unsupported number of upsample stages / resblocks");
  for (int i = 0; i < d->n_ups; ++i) {
    const int u = d->up_rates[i], k = d->up_kernels[i];
    if (u < 1 || k % u != 0 || (k - u) % 2 != 0)
      return fail(DTTS_ERR_BAD_SHAPE, "upsample kernel must be a multiple of its rate with even (k-u)");
  }
  if (d->precision != 0) return fail(DTTS_ERR_BAD_ARG, "vocoder precision mode not available in this build");
  DTTS_TRY(arch_check());
  dtts_vocoder* h = new dtts_vocoder();
  h->desc = *d;
  cudaStream_t s = (cudaStream_t)stream;
  int rc = h->tab.init(arena_dev, arena_floats, table, n_entries);
  if (rc != DTTS_OK) { delete h; return rc; }
  size_t total = 0;
  for (auto& e : h->tab.entries) total += e.second.second + 64;
  rc = h->pool.reserve(total);
  if (rc != DTTS_OK) { delete h; return rc; }
  auto bail = [&](int code) { h->pool.release(); delete h; return code; };
  rc = pack_conv(h, "conv_pre", d->init_ch, d->n_mel, 7, &h->conv_pre, s);
  if (rc != DTTS_OK) return bail(rc);
  int ch = d->init_ch;
  size_t len = 1;
  h->unit = (size_t)d->init_ch;
  for (int i = 0; i < d->n_ups; ++i) {
    ConvW u;
    rc = pack_convT(h, "ups." + std::to_string(i), ch, ch / 2, d->up_kernels[i], d->up_rates[i], &u, s);
    if (rc != DTTS_OK) return bail(rc);
    h->ups.push_back(u);
    ch /= 2;
    len *= d->up_rates[i];
    if ((size_t)ch * len > h->unit) h->unit = (size_t)ch * len;
    for (int j = 0; j < d->n_rb; ++j) {
      const std::string r = "resblocks." + std::to_string(i * d->n_rb + j);
      for (int m = 0; m < 3; ++m) {
        ConvW c1, c2;
        rc = pack_conv(h, r + ".convs1." + std::to_string(m), ch, ch, d->rb_kernels[j], &c1, s);
        if (rc != DTTS_OK) return bail(rc);
        rc = pack_conv(h, r + ".convs2." + std::to_string(m), ch, ch, d->rb_kernels[j], &c2, s);
        if (rc != DTTS_OK) return bail(rc);
        h->rb1.push_back(c1);
        h->rb2.push_back(c2);
      }
    }
  }
  h->hop = (int)len;
  rc = pack_conv(h, "conv_post", 1, ch, 7, &h->conv_post, s);
  if (rc != DTTS_OK) return bail(rc);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return bail(fail(DTTS_ERR_CUDA, std::string("vocoder create: ") + cudaGetErrorString(e)));
  *out = h;
  return DTTS_OK;
}

extern "C" int dtts_vocoder_destroy(dtts_vocoder* h) {
  if (!h) return DTTS_OK;
  h->pool.release();
  delete h;
  return DTTS_OK;
}

extern "C" uint64_t dtts_vocoder_launch_count(const dtts_vocoder* h) { return h ? h->launches : 0; }

extern "C" uint64_t dtts_vocode_workspace_bytes(const dtts_vocoder* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return 5 * ws_round((size_t)B * T * h->unit * sizeof(float)) + 1024;
}

extern "C" int dtts_vocode(dtts_vocoder* h, const float* mel, int32_t B, int32_t T, float* wav, void* ws,
                           uint64_t ws_bytes, void* stream) {
  if (!h || !mel || !wav || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_vocode: null argument");
  if (B <= 0 || T <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_vocode: B and T must be positive");
  if (ws_bytes < dtts_vocode_workspace_bytes(h, B, T))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_vocode: workspace too small");
  const dtts_vocoder_desc& d = h->desc;
  Bump bump(ws, ws_bytes);
  float* buf[5];
  for (int i = 0; i < 5; ++i) buf[i] = bump.take<float>((size_t)B * T * h->unit);
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_vocode: workspace too small");
  Launcher L;
  L.stream = (cudaStream_t)stream;
  L.counter = &h->launches;
  cudaStream_t s = L.stream;

  float* x = buf[0];      // stage input
  float* xu = buf[1];     // upsampled
  float* t1 = buf[2];     // resblock temporary
  float* y = buf[3];      // resblock running value
  float* acc = buf[4];    // mean over resblocks

  // conv_pre reads the mel in its native [B,T,n_mel] layout
  {
    ConvParams p = conv_params(mel, T, h->conv_pre, 0, d.init_ch, x, T, 1, 1, 3);
    p.x_bs = (long)T * d.n_mel; p.x_cs = 1; p.x_ts = d.n_mel;
    L(launch_conv1d_f32(p, B, s));
  }
  int ch = d.init_ch, len = T;
  for (int i = 0; i < d.n_ups; ++i) {
    const int u = d.up_rates[i], k = d.up_kernels[i];
    const int len_o = len * u;
    {
      ConvParams p = convT_params(x, len, h->ups[i], xu, len_o, u, (k - u) / 2);
      p.pre_slope = 0.1f;
      L(launch_conv1d_f32(p, B, s));
    }
    ch /= 2;
    len = len_o;
    for (int j = 0; j < d.n_rb; ++j) {
      const int kr = d.rb_kernels[j];
      for (int m = 0; m < 3; ++m) {
        const int dil = d.rb_dilations[j][m];
        const ConvW& c1 = h->rb1[(i * d.n_rb + j) * 3 + m];
        const ConvW& c2 = h->rb2[(i * d.n_rb + j) * 3 + m];
        const float* yin = (m == 0) ? xu : y;
        ConvParams p1 = conv_params(yin, len, c1, 0, ch, t1, len, dil, 1, (kr * dil - dil) / 2);
        p1.pre_slope = 0.1f;
        L(launch_conv1d_f32(p1, B, s));
        float* dst = (m == 2) ? acc : y;
        ConvParams p2 = conv_params(t1, len, c2, 0, ch, dst, len, 1, 1, (kr - 1) / 2);
        p2.pre_slope = 0.1f;
        p2.res = yin; p2.r_bs = (long)ch * len; p2.r_cs = len; p2.r_ts = 1;
        if (m == 2) {                       // xs (+)= resblock_j(x);  x = xs / num_kernels
          p2.post = 1.f / (float)d.n_rb;
          p2.accumulate = (j > 0);
        }
        L(launch_conv1d_f32(p2, B, s));
      }
    }
    float* tmp = x; x = acc; acc = tmp;
  }
  {
    ConvParams p = conv_params(x, len, h->conv_post, 0, 1, wav, len, 1, 1, 3);
    p.pre_slope = 0.01f;                    // F.leaky_relu default slope (hifigan.py:138)
    p.act = ACT_TANH;
    L(launch_conv1d_f32(p, B, s));
  }
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_vocode: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}
