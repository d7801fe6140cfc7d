// Tensor-core 1-D convolution for the HiFi-GAN stack: an implicit GEMM on tcgen05 (sm_100a) with TMEM accumulators.
//
//   D[q, co] = sum_{tap j} sum_{ci} A[q + off_j, ci] * W_j[co, ci]        M = time (128 rows / MMA), N = C_out, K = C_in
//
// Operands are 16-bit (bf16 or fp16, TcMode::fmt) and either operand may be fed as hi + lo planes: the kernel issues
// a_hi*w_hi (+ a_hi*w_lo when the weights are split) (+ a_lo*w_hi when the activations are split), fp32 accumulate
// in TMEM.  bf16 hi/lo on both sides (3 MMAs) is fp32-class (~2^-16 per product); fp16 activations against hi/lo fp16
// weights (2 MMAs) leaves only the 2^-12 activation rounding; see DESIGN.md "vocoder precision" for the measured
// waveform error of every mode against the 1e-4 RMS tolerance.
//
// Memory layout ("slab planes", chosen so that a time shift is a pure address offset for the UMMA descriptor):
//   operand plane   bf16  [B][C/8][rows][8]   row = pad + t ; rows outside [pad, pad+T) are zero (conv zero padding)
//   fp32 stream     f32   [B][C/4][T][4]      residual path (resblock x + conv(x), mean over resblocks)
// One slab (8 channels x rows) is contiguous, so a CTA stages its (tile + halo) rows of each slab with ONE
// cp.async.bulk, lands them in shared memory as the canonical no-swizzle K-major UMMA layout (core matrix = 8 rows x 16 B,
// SBO = 128 B, LBO = rows*16 B) and reads every tap of the convolution from the same staged tile by moving the
// descriptor start address -- activations cross L2->SMEM once per K-chunk instead of once per tap.
//
// Reference ops replaced: nn.Conv1d / nn.ConvTranspose1d + F.leaky_relu + residual adds in
// modules/hifigan/hifigan.py:27-58 (ResBlock1), :101-142 (HifiGanGenerator.forward).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dtts {

typedef uint16_t tc16;   // storage of one operand element (bf16 or fp16 bit pattern)

// Operand format of a tensor-core convolution.
struct TcMode {
  int fmt = 1;        // 0: fp16, 1: bf16
  int a_planes = 2;   // activations: 1 = single rounding, 2 = hi + lo
  int w_planes = 2;   // weights:     1 = single rounding, 2 = hi + lo
  int hybrid = 0;     // 1: layers with C_out >= 128 use ONE weight plane (there the second MMA costs time); narrower layers
                      //    keep hi | lo stacked along N, where the second plane is free
  int lo8 = 0;        // 1: in layers with C_out >= 128 the lo-plane correction a * w_lo runs as an FP8 MMA
                      //    (kind::f8f6f4, e5m2(a) x e5m2(w_lo * 2^10), K = 32 per instruction = twice the fp16 rate):
                      //    it corrects a 2^-11 term, so 2 mantissa bits are plenty (measured: same waveform error as two
                      //    fp16 planes) and the layer issues 1.5 instead of 2 units of tensor work
  int mmas() const { return a_planes + w_planes - 1; }
};
// dtts_vocoder_desc.precision -> mode (1: bf16 3-MMA split, 2: bf16, 3: fp16 x fp16 hi/lo weights, 4: fp16,
// 5: fp16 x fp16, hi/lo weights only where C_out < 128, 6: as 3 with the lo plane of the C_out >= 128 convolutions in FP8)
static inline TcMode tc_mode(int precision) {
  TcMode m;
  switch (precision) {
    case 2: m.fmt = 1; m.a_planes = 1; m.w_planes = 1; break;
    case 3: m.fmt = 0; m.a_planes = 1; m.w_planes = 2; break;
    case 4: m.fmt = 0; m.a_planes = 1; m.w_planes = 1; break;
    case 5: m.fmt = 0; m.a_planes = 1; m.w_planes = 2; m.hybrid = 1; break;
    case 6: m.fmt = 0; m.a_planes = 1; m.w_planes = 2; m.lo8 = 1; break;
    default: m.fmt = 1; m.a_planes = 2; m.w_planes = 2; break;
  }
  return m;
}

constexpr int TC_PADF = 32;        // zero rows in front of t = 0 (>= largest left halo: k=11, d=5 -> 25)
constexpr int TC_PADB = 32;        // zero rows kept after the last tile
constexpr int TC_ROW_ALIGN = 512;  // tiles are at most 512 rows
constexpr int TC_MAX_RAGGED_ITEMS = 512;   // batch items of a launch with per-item row limits (TcConvParams::lens)
// rows allocated per slab for a tensor of T time steps
static inline int tc_rows(int T) {
  const int align = T + 8 <= 128 ? 128 : (T + 8 <= 256 ? 256 : TC_ROW_ALIGN);   // short sequences use 128/256-row tiles
  return TC_PADF + (T + 8 + align - 1) / align * align + TC_PADB;
}

// DTTS_TC_PAIR=0 turns the CTA-pair path off (default on)
int tc_pair_enabled();
// widest MMA N of a vocoder convolution (DTTS_TC_NMAX: 64, 128 or 256): C_out = 256 as ONE N = 256 block (a 128-row tile per
// CTA: TMEM holds two 256-column accumulator sets) or as two N = 128 blocks (256-row tiles: every weight byte that reaches
// shared memory is used for twice as many rows, the activation tile is staged once per block)
int tc_nmax();
// fewest taps of a convolution that uses the FP8 lo plane under TcMode::lo8 (DTTS_TC_LO8_MINTAPS, default 7: the k = 3
// layers are latency / HBM bound, the extra conversion step in front of their MMAs costs more than the MMAs it saves)
int tc_lo8_min_taps();

// Packed weights of one convolution: blobs [group][chunk][tap] of {hi plane, lo plane}, each plane [KC/8][N][8] bf16
// (the shared-memory image of the B operand).  group = nblock * phases + phase.
struct TcConvW {
  const tc16* w = nullptr;
  const float* bias = nullptr;   // [C_out] fp32
  int C_in = 0, C_out = 0;       // full channel counts
  int N = 0;                     // output channels per MMA (<= 256); C_out = N * nblocks
  int KC = 0;                    // channels per K chunk (16 or 32)
  int ktaps = 0, phases = 1;
  int planes = 2;                // separately addressed weight planes (hi, lo): one MMA each
  int stack = 0;                 // 1: hi | lo stacked along N inside ONE plane (N <= 64): one MMA of N' = 2N, planes = 1
  int fmt = 1;                   // TcMode::fmt
  // transposed convolution in "interleaved" form: all il_u (= stride) polyphase components of a block of il_cb output
  // channels stacked along N (N = il_u * il_cb, phases = 1); the epilogue de-interleaves (t = q*stride - pad + phase)
  int il_u = 0, il_cb = 0;
  // CTA-pair execution (tcgen05 cta_group::2): the two CTAs of a cluster compute adjacent 128-row tiles with ONE M = 256
  // MMA; each stages its own A tile and only half of every weight blob, which halves the shared-memory traffic of the
  // B operand and of the weight stream per SM (the limiter of the C >= 128 layers).  Weights are packed [tap][half].
  int pair = 0;
  // lo plane as FP8 (TcMode::lo8): `w` then holds the hi plane only (planes = 1), packed x 2^10, and w8 the e5m2 lo
  // plane e5m2((w - fp16(w)) * 2^10), [group][K-chunk][tap]([half])[KC/16][N][16] bytes
  const uint8_t* w8 = nullptr;
  int lo8 = 0;
  // rb_pair128.cu: the same blobs as ONE contiguous stream per CTA of the pair, [half][chunk][tap]{fp16 plane(s), e5m2
  // plane}, so that a weight stage of the fused kernel is a single bulk copy (null: not built for this layer)
  const uint8_t* wstream = nullptr;
  size_t stream_bytes() const { return (size_t)C_out * C_in * ktaps * (2 * planes + (lo8 ? 1 : 0)); }
  size_t elems8() const { return (size_t)C_out * C_in * ktaps; }      // bytes of w8
  size_t elems() const {
    return (size_t)(il_u ? il_u : 1) * C_out * C_in * ktaps * phases * planes * (stack ? 2 : 1);
  }
  // operand mode -> plane arrangement of this layer
  void set_mode(const TcMode& m, int lo8_ok = 1) {
    fmt = m.fmt;
    lo8 = (m.lo8 && lo8_ok && !il_u && C_out >= 128 && N >= 128 && KC == 32 && m.a_planes == 1 && m.w_planes == 2 &&
           ktaps >= tc_lo8_min_taps()) ? 1 : 0;
    const int wp = ((m.hybrid && C_out >= 128) || lo8) ? 1 : m.w_planes;
    stack = (wp == 2 && N <= 64) ? 1 : 0;
    planes = stack ? 1 : wp;
    pair = (!stack && m.a_planes == 1 && N >= 128 && tc_pair_enabled()) ? 1 : 0;
  }
};

struct TcConvParams {
  const tc16* a_hi;     // input operand planes
  const tc16* a_lo;
  long a_bs;                     // batch stride in elements
  int a_rows, a_pad;
  const tc16* w;
  const float* bias;
  int C_in, N, KC, nchunks, ktaps, nblocks, phases;
  int a_planes, w_planes, fmt;   // operand planes of A (activations) and B (weights); 16-bit format
  int stack, NM;                 // weights hi | lo stacked along N; N of the main MMAs (2N when stacked, else N)
  int il_u, il_cb;               // interleaved transposed convolution (see TcConvW); 0 = off
  int pair;                      // CTA-pair (cta_group::2) execution
  int tap_off0, tap_step;        // tap j reads input row q + tap_off0 + j*tap_step
  int min_off, RA;               // staged rows per slab: [q0 + min_off, q0 + min_off + RA)
  int nq, MT, NACC;
  int ot_mul, ot_add, T_out;     // output time t = q*ot_mul + ot_add + phase
  float* o32;                    // fp32 stream out (may be null)
  const float* res;              // fp32 stream residual (may be null), same geometry as o32
  long o32_bs;                   // = C_out * T_out
  tc16* o_hi;           // operand planes out (may be null): leaky(v, slope) split in hi / lo
  tc16* o_lo;
  long op_bs;
  int op_rows, op_pad;
  float post, slope;
  int accumulate;
  // generic-stride fp32 output / residual (acoustic model, [B,C,T] or [B,T,C] tensors) instead of the fp32 stream:
  int o_nct;                     // 1: out element (c,t) at o32[b*o32_bs + c*o_cs + t*o_ts], res at res[b*r_bs + c*r_cs + t*r_ts]
  long o_cs, o_ts, r_bs, r_cs, r_ts;
  const float* mask;             // optional [B][m_bs] multiplier per (b, t)
  int m_bs;
  int act;                       // 0: none, 1: ReLU, 3: exact GELU (applied to conv + bias; codes of conv1d_f32.cuh)
  float alpha;                   // out = post * (act(conv + bias) * alpha * mask + res)  (+ out if accumulate)
  int c_valid;                   // o_nct: only output channels < c_valid are stored (weights zero-padded to N % 32 == 0)
  // WaveNet gate fused into the epilogue (o_nct residual = the conditioning slice, o_hi / o_lo required, o32 unused): the
  // weights are packed so that column j < N/2 of a block is a tanh channel and column j + N/2 its sigmoid channel;
  // planes channel (block * N/2 + j) = tanh(conv_j + bias + res) * sigmoid(conv_{j+N/2} + bias + res).  N % 64 == 0.
  int gate;
  // Channel LayerNorm of the OUTPUT fused into the epilogue (acoustic encoders: x = x + conv(..) followed by LN(x)):
  //   x_new = (conv + bias) * mask + res          -> o32 (generic strides), multiplied by ln_in_mask when given
  //   y     = LN_c(x_new) * gamma + beta, * ln_out_mask -> operand planes o_hi / o_lo and (optional) ln_y [same strides as o32]
  // One N block must hold all channels (nblocks == 1, N == C_out), one 128-row sub-tile per tile; not in the CTA-pair
  // build.  Statistics are single-pass (sum, sum of squares) over the values as stored.
  const float* ln_gamma;         // null: off
  const float* ln_beta;
  float ln_eps;
  const float* ln_in_mask;       // [B][m_bs] or null
  const float* ln_out_mask;      // [B][m_bs] or null
  float* ln_y;
  int a_stages, w_stages;
  int TG;                        // taps per weight stage
  int ntiles, B;                 // time tiles per (group, batch item); batch size (set by the launcher)
  int csize, nu;                 // cluster size; units (= csize consecutive row tiles) per weight group (launcher)
  // Ragged launch (optional): item b only needs rows [0, lim_b), lim_b = clamp(lens[b] * len_mul + len_add, 0, nq).
  // Tiles beyond lim_b are neither computed nor stored (what is there is stale); the caller owns the argument why nobody
  // needs them (receptive field of the layers that follow, see tc_vocode).  lens: device int32 [B], B <= 512.
  const int* lens;
  int len_mul, len_add;
  // FP8 lo-plane correction (TcConvW::lo8): e5m2 lo weights; the e5m2 copy of the activations is cut out of the staged
  // fp16 tile in shared memory (high byte of every element).  acc_scale: the accumulator holds conv / acc_scale
  // (lo8 weights are packed x 2^10 so that w_lo lands in the e5m2 range); the epilogue applies it before the bias.
  const uint8_t* w8;
  int lo8;
  float acc_scale;
};

// Fused ResBlock pair (rb_pair.cu, rb_pair128.cu): conv1 (kernel k, dilation dil) -> leaky -> conv2 (kernel k, dilation 1)
// -> + residual for C = 32 / 64 (hi | lo stacked weights) and C = 128 (CTA pairs, two fp16 weight planes or fp16 + FP8),
// single-plane activations; bit-identical to the two tc_conv launches.
struct RbPairParams {
  const tc16* a_hi;              // input operand planes [B][4][a_rows][8] (already leaky-ReLU'd by their producer)
  long a_bs;
  int a_rows, a_pad;
  const tc16* w1;                // stacked weight blobs [tap][4 slabs][64][8] of conv1 / conv2 (TcConvW::w)
  const tc16* w2;
  const float* b1;               // [32] fp32
  const float* b2;
  int k, dil;
  int T;                         // sequence length (rows)
  int fmt;
  float slope;                   // leaky slope of the intermediate AND of the output planes
  const float* res;              // fp32 stream residual (may be null) / output [B][8][T][4]
  float* o32;
  long o32_bs;
  tc16* o_hi;                    // output operand planes (may be null)
  long op_bs;
  int op_rows, op_pad;
  float post;
  int accumulate;
  const int* lens;               // ragged launch (optional), as TcConvParams
  int len_mul, len_add;
  int B;
  int C;                         // channels: 32 (weights resident in shared memory) or 64 (weights streamed per tile)
  int S, ntiles;                 // set by the launcher: tile stride 256 - (k - 1), tiles per item
  int TG, a_stages, csize;       // C = 64: taps per weight stage, input stages, CTAs sharing the weight stream (launcher)
  int w_stages, pf;              // C = 128: depth of the weight ring, L2 prefetch of the next tile's input (launcher)
  // conv_post folded into the epilogue of the LAST pair (C = 32 only): instead of storing the final fp32 stream (o32 is
  // then only read, for `accumulate`) every row writes the seven per-tap partial dot products
  //   part[b][j][t] = sum_c post_w[c][j] * leaky(out[b, c, t], post_slope)
  // and tc_conv_post_finish adds the shifted partials: wav[t] = tanh(bias + sum_j part[j][t + j - 3]).
  const float* post_w;           // [32][7] (conv_post.weight), null: off
  float* post_part;              // [B][7][T]
  float post_slope;
  // C = 128 (rb_pair128.cu, CTA pairs): the weights as per-CTA streams (TcConvW::wstream); w_planes fp16 planes (2: hi +
  // lo, one MMA each) or, with lo8, one fp16 plane x 2^10 plus the e5m2 lo plane
  const uint8_t* w1s;
  const uint8_t* w2s;
  int w_planes, lo8;
  float acc_scale;               // the accumulators hold conv / acc_scale (TcConvParams::acc_scale)
};
// Whole ResBlock (three pairs, kernel 3, C = 32 or 64) in one launch (rb_block.cu): the fp32 residual stream of a row stays in
// registers, the operand planes between the pairs in shared memory; bit-identical to the three pair launches.
struct RbBlockParams {
  int C;                         // channels: 32 (weights resident in shared memory) or 64 (one convolution per phase streamed)
  const tc16* a_hi;              // input operand planes [B][C/8][a_rows][8] (leaky-ReLU'd by their producer)
  long a_bs;
  int a_rows, a_pad;
  const tc16* w[6];              // stacked weight blobs of conv1(0), conv2(0), conv1(1), conv2(1), conv1(2), conv2(2)
  const float* bias[6];
  int dil[3];                    // dilation of conv1 of each pair (conv2: 1)
  int T, fmt;
  float slope;
  const float* res;              // fp32 stream: x of the ResBlock (may be null)
  float* o32;                    // fp32 stream out: post * y (+ out if accumulate)
  long o32_bs;
  tc16* o_hi;                    // output operand planes of leaky(out) (may be null)
  long op_bs;
  int op_rows, op_pad;
  float post;
  int accumulate;
  const int* lens;               // ragged launch (optional), as TcConvParams
  int len_mul, len_add;
  int B;
  int halo, S, ntiles;           // set by the launcher: sum(dil + 1), tile stride 256 - 2 halo, tiles per item
};
int rb_block_supported(const TcConvW* const c1[3], const TcConvW* const c2[3], const int dil[3], int a_planes);
cudaError_t launch_rb_block(RbBlockParams p, cudaStream_t stream);
// the k = 3 ResBlocks of the C = 32 and C = 64 stages as one launch each (default on with the fused pairs; DTTS_TC_FUSE_BLOCK=0 or a
// tc_fuse_override other than 4: off)
int tc_fuse_block_enabled();
// rows the fused kernel may stage past tc_rows(T): its last tile reads up to 256 + halo rows beyond the tile start
constexpr int TC_FUSE_EXTRA_ROWS = 320;
int rb_pair_supported(const TcConvW& c1, const TcConvW& c2, int dil, int a_planes);
cudaError_t launch_rb_pair(RbPairParams p, cudaStream_t stream);
// DTTS_TC_PDL=0 launches the tcgen05 kernels without programmatic dependent launch (default on)
int tc_pdl_enabled();
// DTTS_TC_PAIR64_CLUSTER=1: the fused C = 64 kernel with the 2-CTA weight multicast.  Built, bit-identical, and measured
// on a B200 without any gain (per-launch times equal to the microsecond: the kernel is bound by shared-memory bandwidth --
// MMA operand fetch at N' = 128 needs the full 128 B/clk next to the bulk copies -- not by the L2 -> SM weight stream),
// so it is off by default.
int tc_pair64_cluster_enabled();
// DTTS_TC_FUSE64=0 keeps the C = 64 stage on the two-launch form (default: fused)
int tc_fuse64_enabled();
// largest kernel size whose C = 128 ResBlock pairs run fused (rb_pair128.cu).  DTTS_TC_FUSE128_MAXK, default 3: measured on
// a B200 the fused k = 3 pairs take 535 us against 594 for two launches, k = 7 the same 860, k = 11 1250 against 1047 (the
// 102 KB intermediate tile leaves too little shared memory for the input stages); 0 turns the kernel off,
// 11 fuses all of them (what tc_fuse_override(3) does for the bit-identity tests)
int tc_fuse128_maxk();
// DTTS_TC_FUSE=0 turns the fused ResBlock pairs off (default on)
int tc_fuse_enabled();
void tc_fuse_override(int v);      // -1: environment default; 0 / 1: force (unit tests compare the two builds of a pass);
                                   // 2: fused pairs but conv_post as its own kernel (bit-identical to mode 0);
                                   // 3: as 2 with every C = 128 pair fused too (tc_fuse128_maxk)
                                   // 4: as 2 with the k = 3 ResBlock of the C = 32 stage as ONE launch (rb_block.cu)
// conv_post folded into the last fused pair (default on with the fused pairs; DTTS_TC_FOLD_POST=0 or override 2: off)
int tc_fold_post_enabled();

// output channels per N block of an interleaved transposed convolution: the largest multiple of 8 dividing C_out with
// stride * cb <= 256 and stride * cb a multiple of 32 (0: unsupported)
static inline int tc_il_block(int C_out, int stride) {
  if (stride < 2 || (stride & 1)) return 0;
  for (int cb = 256 / stride / 8 * 8; cb >= 8; cb -= 8)
    if (C_out % cb == 0 && (stride * cb) % 32 == 0) return cb;
  return 0;
}

// Launch plan / launcher.
cudaError_t tc_conv_init();
cudaError_t launch_tc_conv(TcConvParams p, int B, cudaStream_t stream);
// Fills the derived fields (KC-dependent tiling, shared-memory stages, TMEM columns) of p from the weights + geometry;
// a_planes = planes of the INPUT activation buffers.
void tc_conv_plan(TcConvParams* p, const TcConvW& w, int nq, int a_planes);

// Weight packing: reference layout ([C_out][C_in][K], or [C_in][C_out][K] for ConvTranspose1d) -> blobs.
cudaError_t tc_pack_weights(const float* w_ref, tc16* out, int C_out, int C_in, int K, int transposed,
                            int stride, int N, int KC, int planes, int fmt, int stack, cudaStream_t s, int il_cb = 0,
                            int pair = 0, float wscale = 1.f);
// fp32 strided tensor x[b*bs + c*cs + t*ts] -> operand planes of leaky(x, slope) (valid rows only)
cudaError_t tc_to_planes(const float* x, long bs, long cs, long ts, int B, int C, int T, float slope,
                         tc16* hi, tc16* lo, int rows, int pad, int fmt, cudaStream_t s);
// *fits = 1 when fp16(w * 2^10) stays finite with headroom (max |w| < 32); synchronises the stream (create time only).
// A layer whose weights do not fit keeps two fp16 planes.
cudaError_t tc_lo8_weights_fit(const float* w_ref, size_t n, cudaStream_t s, int* fits);
// e5m2 lo plane of the weights for TcConvW::lo8 (layout in TcConvW)
cudaError_t tc_pack_weights_lo8(const float* w_ref, uint8_t* out, int C_out, int C_in, int K, int N, int KC, int fmt,
                                int pair, cudaStream_t s);
// TcConvW::wstream of a packed pair-mode convolution with C_in = C_out = 128 (rb_pair128.cu)
cudaError_t rb_pair128_pack_stream(const TcConvW& w, uint8_t* out, cudaStream_t s);
// same, and zero-fills every row outside [pad, pad+T) (one launch for a freshly re-shaped scratch buffer)
// C_total / c_off: the C source channels become channels [c_off, c_off + C) of planes that hold C_total channels
// s2d_H > 0: space-to-depth source, plane channel ch = x channel ch % s2d_H at time t * ts + ch / s2d_H
cudaError_t tc_to_planes_full(const float* x, long bs, long cs, long ts, int B, int C, int T, float slope, tc16* hi,
                              tc16* lo, int rows, int pad, int fmt, cudaStream_t s, int C_total = 0, int c_off = 0,
                              int s2d_H = 0);
// zero the halo rows [0,pad) and [pad+T, rows) of every slab of a plane pair
cudaError_t tc_zero_halo(tc16* hi, tc16* lo, int n_slabs_total, int rows, int pad, int T,
                         cudaStream_t s);
// fp32 stream <-> [B][C][T]
cudaError_t tc_stream_to_nct(const float* st, float* out, int B, int C, int T, cudaStream_t s);
cudaError_t tc_nct_to_stream(const float* in, float* st, int B, int C, int T, cudaStream_t s);
// operand planes (hi + lo) -> fp32 [B][C][T]  (tests)
cudaError_t tc_planes_to_nct(const tc16* hi, const tc16* lo, float* out, int B, int C, int T,
                             int rows, int pad, int fmt, cudaStream_t s);
// second half of the folded conv_post (RbPairParams::post_part): wav[b,t] = tanh(bias + sum_j part[b][j][t + j - 3]);
// samples t >= lens[b] * len_mul are written as 0
cudaError_t tc_conv_post_finish(const float* part, const float* bias, float* wav, int B, int T, cudaStream_t s,
                                const int* lens = nullptr, int len_mul = 0);
// conv_post: y[b,t] = tanh(bias + sum_{c,j} w[c][j] * leaky(x[b,c,t+j-pad], slope)) from an fp32 stream
// lens (optional, device int32 [B]): samples t >= lens[b] * len_mul of item b are written as 0 without being computed
cudaError_t tc_conv_post(const float* st, const float* w, const float* bias, float* wav, int B, int C, int T, int K,
                         float slope, cudaStream_t s, const int* lens = nullptr, int len_mul = 0);

}  // namespace dtts
