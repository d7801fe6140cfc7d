// The FVAE prior flow (reverse direction) as ONE launch: every residual coupling layer of the block -- pre (1x1), a WaveNet
// of n_layers gated dilation-1 convolutions conditioned on g, post (1x1), mean-only update -- for one 128-row tile of one
// utterance per CTA, activations resident in shared memory, accumulators in TMEM, the weights streamed through a ring
// of bulk copies.  Replaces ~80 dependent launches of 8-20 us (tc_conv_kernel / wn_gate_planes_kernel /
// pointwise_small_kernel per layer) that ran the same arithmetic one layer per launch.
//
// Reference: ResidualCouplingBlock.forward(reverse=True) / ResidualCouplingLayer.forward (modules/commons/
// glow_modules.py:108-128,157-163), WN.forward (modules/commons/wavenet.py:54-78), called from
// FVAE_semantics.forward(infer=True) (modules/dict_tts/fvae_semantics.py:96-101).
#pragma once
#include "tc_conv.cuh"

namespace dtts {

// Weights of the whole flow block in EXECUTION order (the reverse pass runs the last coupling layer first).
struct FlowFusedW {
  uint8_t* stream = nullptr;   // [coupling][layer][stage] 32 KB stages: {bf16 hi plane, bf16 lo plane} of [8 slabs][128][8];
                               // stages of a layer: cond chunks (64 channels of g each), taps 0..2 of in_layers, res_skip
  float* par = nullptr;        // [coupling][par_stride] fp32: biases (cond + in, prefix sums of res / skip), pre / post
  int par_stride = 0;
  int n_flows = 0, n_layers = 0, n_chunks = 0, H = 0;
  uint32_t odd_mask = 0;       // bit e: coupling e (execution order) sees the latent channel-flipped
  size_t stream_bytes() const { return (size_t)n_flows * n_layers * (n_chunks + 4) * 32768; }
  bool ready() const { return stream && par; }
};
// geometry the kernel is written for: hidden 64, latent 16, kernel 3, g channels a multiple of 64 and <= 192
int flow_fused_supported(int H, int flow_hidden, int latent, int flow_kernel, int n_layers, int n_flows);
int flow_fused_par_floats(int n_layers);

// One WaveNet layer of one coupling layer -> its (n_chunks + 4) weight stages.  Reference layouts: in_w [128][64][3],
// cond_w rows [128 * layer, 128 * layer + 128) of [2 * 64 * n_layers][H], rs_w [rs_rows][64] (rs_rows = 128, or 64 for
// the last layer, whose outputs all go to the skip sum).
cudaError_t flow_fused_pack_layer(const float* in_w, const float* cond_w_rows, const float* rs_w, int rs_rows, int H,
                                  uint8_t* out, cudaStream_t s);
struct FlowParSrc {
  const float* in_b[8];
  const float* rs_b[8];
  const float* cond_b;                              // [2 * 64 * n_layers]
  const float *pre_w, *pre_b, *post_w, *post_b;     // ConvW packing ([C_in][C_out]) with the Flip already folded in
  int n_layers;
};
cudaError_t flow_fused_pack_par(const FlowParSrc& src, float* out, cudaStream_t s);

struct FlowFusedParams {
  const tc16* g_hi;            // conditioning g_sqz as bf16 hi / lo operand planes [B][H/8][g_rows][8], row = g_pad + t
  const tc16* g_lo;
  long g_bs;
  int g_rows, g_pad;
  const float* z_in;           // [B][16][T]
  float* z_out;                // [B][16][T]; NOT z_in (a tile reads its neighbours' rows of z_in as halo)
  int T, B;
};
cudaError_t launch_flow_fused(const FlowFusedW& w, const FlowFusedParams& p, cudaStream_t stream);

// DTTS_AC_FUSE=0 keeps the acoustic model on the one-launch-per-layer form (default: fused); the override is the unit
// tests' switch (dtts_debug_set_acoustic_fuse: -1 environment default, 0 off, 1 on)
int ac_fuse_enabled();
void ac_fuse_override(int v);

}  // namespace dtts
