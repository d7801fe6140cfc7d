// HiFi-GAN V1 generator (modules/hifigan/hifigan.py:27-58,101-142) behind dtts_vocode.
// fp32 path: every convolution is one launch of conv1d_f32 with the leaky-ReLU fused into the operand load and the
// residual add / 1/3 resblock mean / tanh fused into the epilogue (the reference launches 78 cuDNN convolutions
// plus 127 element-wise kernels per call).
#include "engine.cuh"
#include "tc_conv.cuh"
#include "tc16.cuh"

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
  // tensor-core path (precision >= 1, see tc_mode() in tc_conv.cuh)
  TcMode mode;
  tc16* tc_pool = nullptr;
  uint8_t* tc_pool8 = nullptr;   // e5m2 lo planes (precision 6)
  uint8_t* tc_pool_s = nullptr;  // per-CTA weight streams of the fused C = 128 ResBlock pairs (rb_pair128.cu)
  TcConvW tc_pre;
  std::vector<TcConvW> tc_ups, tc_rb1, tc_rb2;
  const float *post_w = nullptr, *post_b = nullptr;
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

int tc_pack(dtts_vocoder* h, tc16** cursor, const std::string& name, int C_out, int C_in, int K,
            int transposed, int stride, TcMode mode, TcConvW* cw, cudaStream_t s, uint8_t** cursor8 = nullptr) {
  const float* w = h->tab.get(name + ".weight", (uint64_t)C_out * C_in * K);
  if (!w) return DTTS_ERR_MISSING_WEIGHT;
  const float* b = h->tab.get(name + ".bias", C_out);
  if (!b) return DTTS_ERR_MISSING_WEIGHT;
  cw->C_in = C_in; cw->C_out = C_out;
  cw->KC = (C_in % 32 == 0) ? 32 : 16;
  cw->ktaps = transposed ? K / stride : K;
  cw->phases = 1;
  if (transposed) {                                   // all polyphase components stacked along N (tc_conv.cuh)
    cw->il_u = stride;
    cw->il_cb = tc_il_block(C_out, stride);
    cw->N = cw->il_cb * stride;
    if (!cw->il_cb) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core vocoder: unsupported transposed convolution " + name);
  } else {
    cw->N = C_out > tc_nmax() ? tc_nmax() : C_out;
    if (C_out % cw->N) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core vocoder: unsupported channel count in " + name);
  }
  cw->set_mode(mode);
  if (cw->lo8) {                                      // 2^10-scaled hi plane must stay inside fp16: else two fp16 planes
    int fits = 0;
    DTTS_CUDA(tc_lo8_weights_fit(w, (size_t)C_out * C_in * K, s, &fits));
    if (!fits) cw->set_mode(mode, 0);
  }
  cw->bias = b;
  if (cw->N % 32 || cw->N > 256 || C_in % cw->KC)
    return fail(DTTS_ERR_BAD_SHAPE, "tensor-core vocoder: unsupported channel count in " + name);
  cw->w = *cursor;
  if (cw->lo8) {
    if (!cursor8 || !*cursor8) return fail(DTTS_ERR_CUDA, "tensor-core vocoder: no pool for the e5m2 weight planes");
    cw->w8 = *cursor8;
    DTTS_CUDA(tc_pack_weights_lo8(w, *cursor8, C_out, C_in, K, cw->N, cw->KC, cw->fmt, cw->pair, s));
    *cursor8 += (cw->elems8() + 63) / 64 * 64;
  }
  DTTS_CUDA(tc_pack_weights(w, *cursor, C_out, C_in, K, transposed, stride, cw->N, cw->KC, cw->planes, cw->fmt, cw->stack,
                            s, cw->il_cb, cw->pair, cw->lo8 ? kLo8WScale : 1.f));
  *cursor += (cw->elems() + 63) / 64 * 64;
  return DTTS_OK;
}

int tc_create(dtts_vocoder* h, cudaStream_t s) {
  const dtts_vocoder_desc& d = h->desc;
  h->mode = tc_mode(d.precision);
  const TcMode mode = h->mode;
  const int wp = mode.w_planes;
  // 16-bit elements needed: weight planes * (all conv weights) + alignment slack
  size_t total = (size_t)d.init_ch * d.n_mel * 7 * wp + 64;
  int ch = d.init_ch;
  for (int i = 0; i < d.n_ups; ++i) {
    total += (size_t)ch * (ch / 2) * d.up_kernels[i] * wp + 64;
    ch /= 2;
    for (int j = 0; j < d.n_rb; ++j) total += 6 * ((size_t)ch * ch * d.rb_kernels[j] * wp + 64);
  }
  cudaError_t e = cudaMalloc((void**)&h->tc_pool, total * sizeof(tc16));
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaMalloc(tc weight pool): ") + cudaGetErrorString(e));
  e = tc_conv_init();
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("tc_conv_init: ") + cudaGetErrorString(e));
  tc16* cur = h->tc_pool;
  uint8_t* cur8 = nullptr;
  if (mode.lo8) {                                    // one byte per ResBlock weight, 64-byte aligned entries
    size_t total8 = 0;
    int c8 = d.init_ch;
    for (int i = 0; i < d.n_ups; ++i) {
      c8 /= 2;
      for (int j = 0; j < d.n_rb; ++j) total8 += 6 * ((size_t)c8 * c8 * d.rb_kernels[j] + 64);
    }
    e = cudaMalloc((void**)&h->tc_pool8, total8);
    if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaMalloc(e5m2 weight pool): ") + cudaGetErrorString(e));
    cur8 = h->tc_pool8;
  }
  DTTS_TRY(tc_pack(h, &cur, "conv_pre", d.init_ch, d.n_mel, 7, 0, 1, mode, &h->tc_pre, s));
  ch = d.init_ch;
  for (int i = 0; i < d.n_ups; ++i) {
    TcConvW u;
    DTTS_TRY(tc_pack(h, &cur, "ups." + std::to_string(i), ch / 2, ch, d.up_kernels[i], 1, d.up_rates[i], mode, &u, s));
    h->tc_ups.push_back(u);
    ch /= 2;
    for (int j = 0; j < d.n_rb; ++j) {
      const std::string r = "resblocks." + std::to_string(i * d.n_rb + j);
      for (int m = 0; m < 3; ++m) {
        TcConvW c1, c2;
        DTTS_TRY(tc_pack(h, &cur, r + ".convs1." + std::to_string(m), ch, ch, d.rb_kernels[j], 0, 1, mode, &c1, s, &cur8));
        DTTS_TRY(tc_pack(h, &cur, r + ".convs2." + std::to_string(m), ch, ch, d.rb_kernels[j], 0, 1, mode, &c2, s, &cur8));
        h->tc_rb1.push_back(c1);
        h->tc_rb2.push_back(c2);
      }
    }
  }
  if ((size_t)(cur - h->tc_pool) > total) return fail(DTTS_ERR_CUDA, "tc weight pool overrun");
  {
    // the C = 128 CTA-pair convolutions once more as contiguous per-CTA streams (a weight stage of rb_pair128_kernel is then
    // ONE bulk copy); ~2 MB for HiFi-GAN V1
    auto wants = [](const TcConvW& c) { return c.pair && c.C_in == 128 && c.C_out == 128 && c.N == 128 && c.KC == 32 && !c.il_u; };
    size_t total_s = 0;
    for (auto* v : {&h->tc_rb1, &h->tc_rb2})
      for (const TcConvW& c : *v)
        if (wants(c)) total_s += (c.stream_bytes() + 127) / 128 * 128;
    if (total_s) {
      e = cudaMalloc((void**)&h->tc_pool_s, total_s);
      if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("cudaMalloc(weight streams): ") + cudaGetErrorString(e));
      uint8_t* cs = h->tc_pool_s;
      for (auto* v : {&h->tc_rb1, &h->tc_rb2})
        for (TcConvW& c : *v)
          if (wants(c)) {
            DTTS_CUDA(rb_pair128_pack_stream(c, cs, s));
            c.wstream = cs;
            cs += (c.stream_bytes() + 127) / 128 * 128;
          }
    }
  }
  if (ch % 4 || ch > 64) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core vocoder: conv_post needs <= 64 input channels");
  h->post_w = h->tab.get("conv_post.weight", (uint64_t)ch * 7);
  h->post_b = h->tab.get("conv_post.bias", 1);
  if (!h->post_w || !h->post_b) return DTTS_ERR_MISSING_WEIGHT;
  return DTTS_OK;
}

// per-call buffer geometry of the tensor-core path
// rows per slab of the stage buffers: tc_rows + what the fused ResBlock kernel may stage past its last tile (zero rows)
static inline int voc_rows(int T) { return tc_rows(T) + TC_FUSE_EXTRA_ROWS; }

struct TcGeom {
  size_t plane_elems = 0;   // elements of ONE bf16 plane buffer (max over stages)
  size_t stream_elems = 0;  // floats of one fp32 stream buffer (max over stages)
  size_t mel_plane_elems = 0;
};
TcGeom tc_geom(const dtts_vocoder* h, int B, int T) {
  const dtts_vocoder_desc& d = h->desc;
  TcGeom g;
  g.mel_plane_elems = (size_t)B * d.n_mel * tc_rows(T);
  int ch = d.init_ch, len = T;
  g.plane_elems = (size_t)B * ch * voc_rows(len);
  for (int i = 0; i < d.n_ups; ++i) {
    ch /= 2;
    len *= d.up_rates[i];
    const size_t pe = (size_t)B * ch * voc_rows(len) + 64, se = (size_t)B * ch * len + 64;   // + slack for the odd-pad offset
    if (pe > g.plane_elems) g.plane_elems = pe;
    if (se > g.stream_elems) g.stream_elems = se;
  }
  return g;
}

struct PlaneBuf {
  tc16 *hi = nullptr, *lo = nullptr;
  int C = 0, T = 0, rows = 0;
  long bs() const { return (long)C * rows; }
};

// Valid-length schedule of a vocode call (lens = valid mel frames per item, optional).  An item's wav is only wanted up
// to lens[b] * hop samples; working backwards through the generator, every layer only has to be right on the rows those
// samples can see.  Each launch computes rows [0, lens[b] * mul + add) of item b (clamped to the full length):
//   conv_post (k7)      needs the last stage up to  len * hop + 3
//   stage i convs       compute up to  need_i + rb_halo, rb_halo = max over ResBlocks of sum_m (k-1)/2 * (d_m + 1) (+4):
//                       all six convolutions of a ResBlock use the same limit; what they compute from rows the previous
//                       layer did not produce is garbage that stays within rb_halo rows of the limit
//   ups[i] (stride u)   output rows < lim need q < (lim + pad + u - 1) / u + 1 of its input, which is need_{i-1}
// Valid samples are bit-identical to the full-length call (tests/test_gpu_tensorcore.py); at cfg 2 (300-400 of 400 frames
// valid) this skips ~13 % of stages 2-4.
struct LenSched {
  int stage_add[DTTS_MAX_UPS], ups_add[DTTS_MAX_UPS], rpf[DTTS_MAX_UPS + 1];   // rpf[i]: rows per mel frame before ups[i]
  int pre_add;
};
LenSched len_sched(const dtts_vocoder_desc& d) {
  LenSched ls{};
  ls.rpf[0] = 1;
  for (int i = 0; i < d.n_ups; ++i) ls.rpf[i + 1] = ls.rpf[i] * d.up_rates[i];
  int rb_halo = 0;
  for (int j = 0; j < d.n_rb; ++j) {
    int hsum = 0;
    for (int m = 0; m < 3; ++m) hsum += (d.rb_kernels[j] - 1) / 2 * (d.rb_dilations[j][m] + 1);
    rb_halo = hsum > rb_halo ? hsum : rb_halo;
  }
  rb_halo += 4;
  int need = 3 + 1;                                     // conv_post half width (+1 slack)
  for (int i = d.n_ups - 1; i >= 0; --i) {
    ls.stage_add[i] = need + rb_halo;
    const int u = d.up_rates[i], pad = (d.up_kernels[i] - u) / 2;
    ls.ups_add[i] = (ls.stage_add[i] + pad + u - 1) / u + 1;
    need = ls.ups_add[i];
  }
  ls.pre_add = need;
  return ls;
}

int tc_vocode(dtts_vocoder* h, const float* mel, int B, int T, float* wav, void* ws, uint64_t ws_bytes, cudaStream_t s,
              const int* lens = nullptr) {
  const dtts_vocoder_desc& d = h->desc;
  const LenSched ls = len_sched(d);
  const bool split = h->mode.a_planes == 2;
  const int fmt = h->mode.fmt;
  const TcGeom g = tc_geom(h, B, T);
  Bump bump(ws, ws_bytes);
  PlaneBuf PM, PX, PXU, PT, PY;
  PM.hi = bump.take<tc16>(g.mel_plane_elems);
  PM.lo = split ? bump.take<tc16>(g.mel_plane_elems) : nullptr;
  PlaneBuf* pbs[4] = {&PX, &PXU, &PT, &PY};
  for (PlaneBuf* pb : pbs) {
    pb->hi = bump.take<tc16>(g.plane_elems);
    pb->lo = split ? bump.take<tc16>(g.plane_elems) : nullptr;
  }
  float* XU32 = bump.take<float>(g.stream_elems);
  float* Y32 = bump.take<float>(g.stream_elems);
  float* ACC32 = bump.take<float>(g.stream_elems);
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_vocode: workspace too small");
  Launcher L;
  L.stream = s;
  L.counter = &h->launches;

  auto shape = [&](PlaneBuf& pb, int C, int Tn) {      // re-purpose a plane buffer: new geometry + zero halos
    pb.C = C; pb.T = Tn; pb.rows = (&pb == &PM) ? tc_rows(Tn) : voc_rows(Tn);
    L(tc_zero_halo(pb.hi, pb.lo, B * (C / 8), pb.rows, TC_PADF, Tn, s));
  };
  auto base = [&](const TcConvW& w, const PlaneBuf& in, int nq, int off0, int step) {
    TcConvParams p{};
    p.a_hi = in.hi; p.a_lo = in.lo; p.a_bs = in.bs(); p.a_rows = in.rows; p.a_pad = TC_PADF;
    p.tap_off0 = off0; p.tap_step = step;
    tc_conv_plan(&p, w, nq, h->mode.a_planes);
    p.ot_mul = 1; p.ot_add = 0;
    p.post = 1.f; p.slope = 0.1f; p.accumulate = 0;
    return p;
  };
  auto out_planes = [&](TcConvParams& p, const PlaneBuf& o) {
    p.o_hi = o.hi; p.o_lo = o.lo; p.op_bs = o.bs(); p.op_rows = o.rows; p.op_pad = TC_PADF;
  };

  // mel [B,T,n_mel] -> operand planes (no activation in front of conv_pre)
  shape(PM, d.n_mel, T);
  L(tc_to_planes(mel, (long)T * d.n_mel, 1, d.n_mel, B, d.n_mel, T, 1.f, PM.hi, PM.lo, PM.rows, TC_PADF, fmt, s));
  shape(PX, d.init_ch, T);
  {
    TcConvParams p = base(h->tc_pre, PM, T, -3, 1);
    p.T_out = T;
    p.lens = lens; p.len_mul = 1; p.len_add = ls.pre_add;
    out_planes(p, PX);                                  // leaky(., 0.1) of conv_pre feeds ups.0
    L(launch_tc_conv(p, B, s));
  }
  int ch = d.init_ch, len = T;
  bool post_folded = false;                             // conv_post folded into the last fused pair (rb_pair.cu)
  float* post_part = XU32;                              // [B][7][len]: the ups output is dead by then
  for (int i = 0; i < d.n_ups; ++i) {
    const int u = d.up_rates[i], k = d.up_kernels[i];
    const int len_o = len * u, co = ch / 2;
    const bool last_stage = i == d.n_ups - 1;
    // A transposed convolution writes the samples (t, t+1) of one row with a single 32-byte store; with an odd padding
    // t is odd, so its outputs start 16 bytes into a sector to keep those stores sector-aligned.
    const int odd = ((k - u) / 2) & 1;
    PlaneBuf PXUo = PXU;
    PXUo.hi = PXU.hi + 8 * odd;
    PXUo.lo = PXU.lo ? PXU.lo + 8 * odd : nullptr;
    float* XU = XU32 + 4 * odd;
    shape(PXUo, co, len_o);
    shape(PT, co, len_o);
    shape(PY, co, len_o);
    {
      const TcConvW& w = h->tc_ups[i];
      TcConvParams p = base(w, PX, len + w.ktaps - 1, 0, -1);
      p.ot_mul = u; p.ot_add = -(k - u) / 2; p.T_out = len_o;
      p.o32 = XU; p.o32_bs = (long)co * len_o;
      p.lens = lens; p.len_mul = ls.rpf[i]; p.len_add = ls.ups_add[i];
      out_planes(p, PXUo);
      L(launch_tc_conv(p, B, s));
    }
    ch = co; len = len_o;
    if (!last_stage) shape(PX, ch, len);                // PX is free again: it becomes the next stage's input
    for (int j = 0; j < d.n_rb; ++j) {
      const int kr = d.rb_kernels[j];
      {
        // the whole ResBlock as ONE launch where it is HBM bound as three (k = 3, C = 32 / 64; rb_block.cu): the fp32 stream of
        // a row stays in registers, the planes between the pairs in shared memory
        const TcConvW* b1[3]; const TcConvW* b2[3];
        int dils[3];
        for (int m = 0; m < 3; ++m) {
          b1[m] = &h->tc_rb1[(i * d.n_rb + j) * 3 + m];
          b2[m] = &h->tc_rb2[(i * d.n_rb + j) * 3 + m];
          dils[m] = d.rb_dilations[j][m];
        }
        const bool folds_post = j == d.n_rb - 1 && last_stage;       // (the last pair of the last ResBlock carries conv_post)
        if (tc_fuse_block_enabled() && kr == 3 && !folds_post && rb_block_supported(b1, b2, dils, h->mode.a_planes)) {
          RbBlockParams bp{};
          bp.C = ch;
          bp.a_hi = PXUo.hi; bp.a_bs = PXUo.bs(); bp.a_rows = PXUo.rows; bp.a_pad = TC_PADF;
          for (int m = 0; m < 3; ++m) {
            bp.w[2 * m] = b1[m]->w; bp.w[2 * m + 1] = b2[m]->w;
            bp.bias[2 * m] = b1[m]->bias; bp.bias[2 * m + 1] = b2[m]->bias;
            bp.dil[m] = dils[m];
          }
          bp.T = len; bp.fmt = b1[0]->fmt; bp.slope = 0.1f;
          bp.res = XU; bp.o32 = ACC32; bp.o32_bs = (long)ch * len;
          bp.post = 1.f / (float)d.n_rb; bp.accumulate = j > 0;
          if (j == d.n_rb - 1 && !last_stage) { bp.o_hi = PX.hi; bp.op_bs = PX.bs(); bp.op_rows = PX.rows; bp.op_pad = TC_PADF; }
          bp.lens = lens; bp.len_mul = ls.rpf[i + 1]; bp.len_add = ls.stage_add[i]; bp.B = B;
          L(launch_rb_block(bp, s));
          continue;
        }
      }
      for (int m = 0; m < 3; ++m) {
        const int dil = d.rb_dilations[j][m];
        const TcConvW& c1 = h->tc_rb1[(i * d.n_rb + j) * 3 + m];
        const TcConvW& c2 = h->tc_rb2[(i * d.n_rb + j) * 3 + m];
        if (tc_fuse_enabled() && rb_pair_supported(c1, c2, dil, h->mode.a_planes)) {
          // one launch for the pair: the intermediate activation stays in shared memory (rb_pair.cu).  The tiles of a
          // launch read their neighbours' rows as halo, so the output planes must not be the input planes: the three
          // pairs of a ResBlock go PXU -> PY -> PT -> (PX)
          const PlaneBuf& in = m == 0 ? PXUo : (m == 1 ? PY : PT);
          const PlaneBuf& outp = m == 0 ? PY : PT;
          RbPairParams fp{};
          fp.a_hi = in.hi; fp.a_bs = in.bs(); fp.a_rows = in.rows; fp.a_pad = TC_PADF;
          fp.w1 = c1.w; fp.w2 = c2.w; fp.b1 = c1.bias; fp.b2 = c2.bias;
          fp.w1s = c1.wstream; fp.w2s = c2.wstream; fp.w_planes = c1.planes; fp.lo8 = c1.lo8;
          fp.acc_scale = c1.lo8 ? 1.f / kLo8WScale : 1.f;
          fp.k = kr; fp.dil = dil; fp.T = len; fp.fmt = c1.fmt; fp.slope = 0.1f; fp.C = ch;
          fp.res = m == 0 ? XU : Y32;
          fp.o32_bs = (long)ch * len;
          fp.post = 1.f; fp.accumulate = 0;
          if (m < 2) {
            fp.o32 = Y32;
            fp.o_hi = outp.hi; fp.op_bs = outp.bs(); fp.op_rows = outp.rows; fp.op_pad = TC_PADF;
          } else {
            fp.o32 = ACC32;
            fp.post = 1.f / (float)d.n_rb;
            fp.accumulate = j > 0;
            if (j == d.n_rb - 1 && !last_stage) { fp.o_hi = PX.hi; fp.op_bs = PX.bs(); fp.op_rows = PX.rows; fp.op_pad = TC_PADF; }
            if (j == d.n_rb - 1 && last_stage && ch == 32 && tc_fold_post_enabled() &&
                (size_t)7 * B * len <= g.stream_elems) {
              fp.post_w = h->post_w; fp.post_part = post_part; fp.post_slope = 0.01f;     // F.leaky_relu default (hifigan.py:138)
              post_folded = true;
            }
          }
          fp.lens = lens; fp.len_mul = ls.rpf[i + 1]; fp.len_add = ls.stage_add[i]; fp.B = B;
          L(launch_rb_pair(fp, s));
          continue;
        }
        TcConvParams p1 = base(c1, m == 0 ? PXUo : PY, len, -(kr * dil - dil) / 2, dil);
        p1.T_out = len;
        p1.lens = lens; p1.len_mul = ls.rpf[i + 1]; p1.len_add = ls.stage_add[i];
        out_planes(p1, PT);
        L(launch_tc_conv(p1, B, s));
        TcConvParams p2 = base(c2, PT, len, -(kr - 1) / 2, 1);
        p2.T_out = len;
        p2.lens = lens; p2.len_mul = ls.rpf[i + 1]; p2.len_add = ls.stage_add[i];
        p2.res = m == 0 ? XU : Y32;
        p2.o32_bs = (long)ch * len;
        if (m < 2) {
          p2.o32 = Y32;
          out_planes(p2, PY);
        } else {                                         // xs (+)= resblock_j(x); x = xs / num_kernels
          p2.o32 = ACC32;
          p2.post = 1.f / (float)d.n_rb;
          p2.accumulate = j > 0;
          if (j == d.n_rb - 1 && !last_stage) out_planes(p2, PX);
        }
        L(launch_tc_conv(p2, B, s));
      }
    }
  }
  // F.leaky_relu default slope 0.01 (hifigan.py:138), conv_post, tanh
  if (post_folded) L(tc_conv_post_finish(post_part, h->post_b, wav, B, len, s, lens, ls.rpf[d.n_ups]));
  else L(tc_conv_post(ACC32, h->post_w, h->post_b, wav, B, ch, len, 7, 0.01f, s, lens, ls.rpf[d.n_ups]));
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_vocode(tc): ") + cudaGetErrorString(L.err));
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
  if (d->precision < 0 || d->precision > 6) return fail(DTTS_ERR_BAD_ARG, "vocoder precision must be 0..6");
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
  if (d->precision != 0) {
    rc = tc_create(h, s);
    if (rc != DTTS_OK) {
      if (h->tc_pool) cudaFree(h->tc_pool);
      if (h->tc_pool8) cudaFree(h->tc_pool8);
      if (h->tc_pool_s) cudaFree(h->tc_pool_s);
      return bail(rc);
    }
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return bail(fail(DTTS_ERR_CUDA, std::string("vocoder create: ") + cudaGetErrorString(e)));
  *out = h;
  return DTTS_OK;
}

extern "C" int dtts_vocoder_destroy(dtts_vocoder* h) {
  if (!h) return DTTS_OK;
  h->pool.release();
  if (h->tc_pool) cudaFree(h->tc_pool);
  if (h->tc_pool8) cudaFree(h->tc_pool8);
  if (h->tc_pool_s) cudaFree(h->tc_pool_s);
  delete h;
  return DTTS_OK;
}

extern "C" uint64_t dtts_vocoder_launch_count(const dtts_vocoder* h) { return h ? h->launches : 0; }

extern "C" uint64_t dtts_vocode_workspace_bytes(const dtts_vocoder* h, int32_t B, int32_t T) {
  if (!h || B <= 0 || T <= 0) return 0;
  if (h->desc.precision != 0) {
    const TcGeom g = tc_geom(h, B, T);
    const int planes = h->mode.a_planes;
    return planes * (ws_round(g.mel_plane_elems * 2) + 4 * ws_round(g.plane_elems * 2)) +
           3 * ws_round(g.stream_elems * 4) + 4096;
  }
  return 5 * ws_round((size_t)B * T * h->unit * sizeof(float)) + 1024;
}

__global__ void zero_tail_kernel(float* __restrict__ wav, const int* __restrict__ lens, int hop, long n) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t < n && t >= (long)lens[b] * hop) wav[(size_t)b * n + t] = 0.f;
}

static int vocode_impl(dtts_vocoder* h, const float* mel, const int32_t* lens, int32_t B, int32_t T, float* wav,
                       void* ws, uint64_t ws_bytes, void* stream);

extern "C" int dtts_vocode(dtts_vocoder* h, const float* mel, int32_t B, int32_t T, float* wav, void* ws,
                           uint64_t ws_bytes, void* stream) {
  return vocode_impl(h, mel, nullptr, B, T, wav, ws, ws_bytes, stream);
}

extern "C" int dtts_vocode_lens(dtts_vocoder* h, const float* mel, const int32_t* lens_dev, int32_t B, int32_t T,
                                float* wav, void* ws, uint64_t ws_bytes, void* stream) {
  if (!lens_dev) return fail(DTTS_ERR_BAD_ARG, "dtts_vocode_lens: null lengths");
  return vocode_impl(h, mel, lens_dev, B, T, wav, ws, ws_bytes, stream);
}

static int vocode_impl(dtts_vocoder* h, const float* mel, const int32_t* lens, int32_t B, int32_t T, float* wav,
                       void* ws, uint64_t ws_bytes, void* stream) {
  if (!h || !mel || !wav || !ws) return fail(DTTS_ERR_BAD_ARG, "dtts_vocode: null argument");
  if (B <= 0 || T <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_vocode: B and T must be positive");
  if (ws_bytes < dtts_vocode_workspace_bytes(h, B, T))
    return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_vocode: workspace too small");
  const dtts_vocoder_desc& d = h->desc;
  if (d.precision != 0) {
    if (!lens || B <= TC_MAX_RAGGED_ITEMS) return tc_vocode(h, mel, B, T, wav, ws, ws_bytes, (cudaStream_t)stream, lens);
    // the ragged tile schedule keeps its per-item tables in shared memory (TC_MAX_RAGGED_ITEMS): larger batches run as
    // consecutive sub-batches on the same workspace (same stream, so they serialise on it)
    for (int b0 = 0; b0 < B; b0 += TC_MAX_RAGGED_ITEMS) {
      const int nb = B - b0 < TC_MAX_RAGGED_ITEMS ? B - b0 : TC_MAX_RAGGED_ITEMS;
      DTTS_TRY(tc_vocode(h, mel + (size_t)b0 * T * d.n_mel, nb, T, wav + (size_t)b0 * T * h->hop, ws, ws_bytes,
                         (cudaStream_t)stream, lens + b0));
    }
    return DTTS_OK;
  }
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
  if (lens) {                                // the exact fp32 path computes every frame; same contract: zero past the end
    const long n = (long)len;
    zero_tail_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)B), 256, 0, s>>>(wav, lens, h->hop, n);
    L(cudaGetLastError());
  }
  if (L.err != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("dtts_vocode: ") + cudaGetErrorString(L.err));
  return DTTS_OK;
}

extern "C" int dtts_debug_set_tc_fuse(int32_t mode) {
  if (mode < -1 || mode > 4) return fail(DTTS_ERR_BAD_ARG, "dtts_debug_set_tc_fuse: mode must be -1 .. 4");
  tc_fuse_override(mode);
  return DTTS_OK;
}

// Unit-test hook for the tensor-core convolution: fp32 [B,C_in,T_in] in, fp32 [B,C_out,T_out] out (and optionally the
// leaky-ReLU'd operand planes it would hand to the next layer, reconstructed as hi+lo).
extern "C" int dtts_debug_tc_conv1d(const float* x, const float* w, const float* bias, const float* res, float* out,
                                    float* out_act, int32_t B, int32_t C_in, int32_t T_in, int32_t C_out, int32_t K,
                                    int32_t stride, int32_t padding, int32_t dilation, int32_t transposed,
                                    float pre_slope, float post, float act_slope, int32_t precision, void* scratch,
                                    uint64_t scratch_bytes, void* stream) {
  if (precision < 1 || precision > 6) return fail(DTTS_ERR_BAD_ARG, "dtts_debug_tc_conv1d: precision must be 1..6");
  const TcMode mode = tc_mode(precision);
  const bool split = mode.a_planes == 2;
  if (!x || !w || !out || !scratch) return fail(DTTS_ERR_BAD_ARG, "dtts_debug_tc_conv1d: null argument");
  if (!transposed && stride != 1) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core conv: stride must be 1");
  if (transposed && (K % stride || (K - stride) % 2 || padding != (K - stride) / 2))
    return fail(DTTS_ERR_BAD_SHAPE, "tensor-core transposed conv: K % stride == 0 and padding == (K-stride)/2");
  DTTS_TRY(arch_check());
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = tc_conv_init();
  if (e != cudaSuccess) return fail(DTTS_ERR_CUDA, std::string("tc_conv_init: ") + cudaGetErrorString(e));
  TcConvW cw;
  cw.C_in = C_in; cw.C_out = C_out; cw.KC = (C_in % 32 == 0) ? 32 : 16;
  cw.ktaps = transposed ? K / stride : K; cw.phases = 1;
  if (transposed) {
    cw.il_u = stride; cw.il_cb = tc_il_block(C_out, stride); cw.N = cw.il_cb * stride;
    if (!cw.il_cb) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core transposed conv: unsupported stride / channels");
  } else {
    cw.N = C_out > 256 ? 256 : C_out;
    if (C_out % cw.N) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core conv: unsupported channels");
  }
  cw.set_mode(mode); cw.bias = bias;
  if (cw.lo8) {
    int fits = 0;
    DTTS_CUDA(tc_lo8_weights_fit(w, (size_t)C_out * C_in * K, s, &fits));
    if (!fits) cw.set_mode(mode, 0);
  }
  if (cw.N % 32 || cw.N > 256 || C_in % cw.KC) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core conv: unsupported channels");
  const int T_out = transposed ? (T_in - 1) * stride - 2 * padding + K : T_in + 2 * padding - dilation * (K - 1);
  if (T_out <= 0) return fail(DTTS_ERR_BAD_SHAPE, "tensor-core conv: empty output");
  Bump bump(scratch, scratch_bytes);
  const int rows_in = tc_rows(T_in), rows_out = tc_rows(T_out);
  tc16* wp = bump.take<tc16>(cw.elems());
  tc16* a_hi = bump.take<tc16>((size_t)B * C_in * rows_in);
  tc16* a_lo = split ? bump.take<tc16>((size_t)B * C_in * rows_in) : nullptr;
  tc16* o_hi = bump.take<tc16>((size_t)B * C_out * rows_out);
  tc16* o_lo = split ? bump.take<tc16>((size_t)B * C_out * rows_out) : nullptr;
  float* o32 = bump.take<float>((size_t)B * C_out * T_out);
  float* r32 = res ? bump.take<float>((size_t)B * C_out * T_out) : nullptr;
  uint8_t* w8 = cw.lo8 ? bump.take<uint8_t>(cw.elems8() + 64) : nullptr;     // precision 6: e5m2 lo plane
  if (!bump.ok) return fail(DTTS_ERR_WORKSPACE_TOO_SMALL, "dtts_debug_tc_conv1d: scratch too small");
  cw.w = wp;
  if (cw.lo8) {
    cw.w8 = w8;
    DTTS_CUDA(tc_pack_weights_lo8(w, w8, C_out, C_in, K, cw.N, cw.KC, cw.fmt, cw.pair, s));
  }
  DTTS_CUDA(tc_pack_weights(w, wp, C_out, C_in, K, transposed, stride, cw.N, cw.KC, cw.planes, cw.fmt, cw.stack, s,
                            cw.il_cb, cw.pair, cw.lo8 ? kLo8WScale : 1.f));
  DTTS_CUDA(tc_zero_halo(a_hi, a_lo, B * (C_in / 8), rows_in, TC_PADF, T_in, s));
  DTTS_CUDA(tc_to_planes(x, (long)C_in * T_in, T_in, 1, B, C_in, T_in, pre_slope, a_hi, a_lo, rows_in, TC_PADF, mode.fmt, s));
  if (res) DTTS_CUDA(tc_nct_to_stream(res, r32, B, C_out, T_out, s));
  TcConvParams p{};
  p.a_hi = a_hi; p.a_lo = a_lo; p.a_bs = (long)C_in * rows_in; p.a_rows = rows_in; p.a_pad = TC_PADF;
  int nq;
  if (transposed) { p.tap_off0 = 0; p.tap_step = -1; nq = T_in + cw.ktaps - 1; }
  else { p.tap_off0 = -padding; p.tap_step = dilation; nq = T_out; }
  tc_conv_plan(&p, cw, nq, mode.a_planes);
  p.ot_mul = transposed ? stride : 1; p.ot_add = transposed ? -padding : 0; p.T_out = T_out;
  p.o32 = o32; p.res = r32; p.o32_bs = (long)C_out * T_out;
  p.o_hi = o_hi; p.o_lo = o_lo; p.op_bs = (long)C_out * rows_out; p.op_rows = rows_out; p.op_pad = TC_PADF;
  p.post = post; p.slope = act_slope; p.accumulate = 0;
  DTTS_CUDA(launch_tc_conv(p, B, s));
  DTTS_CUDA(tc_stream_to_nct(o32, out, B, C_out, T_out, s));
  if (out_act) DTTS_CUDA(tc_planes_to_nct(o_hi, o_lo, out_act, B, C_out, T_out, rows_out, TC_PADF, mode.fmt, s));
  return DTTS_OK;
}
