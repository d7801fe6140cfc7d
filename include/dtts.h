/* libdtts -- C ABI of the B200-native Dict-TTS inference engine (text -> mel -> waveform forward pass).
 *
 * Drop-in boundary: everything the reference computes inside
 *   PortaSpeech_dict.forward(infer=True)        /root/reference/modules/dict_tts/model.py:36-62
 *   HifiGAN.spec2wav -> HifiGanGenerator.forward /root/reference/vocoders/hifigan.py:54-62,
 *                                                /root/reference/modules/hifigan/hifigan.py:126-142
 * is reached through the entry points below.  The reference is pure Python/PyTorch (no FFI of its own); the
 * binding a maintainer adds is the ctypes stub in INTEGRATION.md (shipped as dict_tts_b200/binding.py).
 *
 * Conventions
 *  - plain C types only; every pointer named *_dev is DEVICE memory owned by the caller (PyTorch allocates it);
 *    the library allocates device memory only at *_create (its re-laid-out copy of the weights) and frees it
 *    at *_destroy.  Workspaces are caller-provided, sized by the *_workspace_bytes queries.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and no call synchronises, except
 *    where stated.  Handles are not re-entrant; one handle per (process, device).
 *  - every call returns 0 on success or a negative dtts_status; dtts_last_error() gives the message.
 *  - float tensors are fp32, index tensors int64 (PyTorch's LongTensor), row-major, contiguous.
 */
#ifndef DTTS_H_
#define DTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTTS_ABI_VERSION 5

typedef enum dtts_status {
  DTTS_OK = 0,
  DTTS_ERR_BAD_ARG = -1,
  DTTS_ERR_BAD_SHAPE = -2,
  DTTS_ERR_MISSING_WEIGHT = -3,
  DTTS_ERR_WORKSPACE_TOO_SMALL = -4,
  DTTS_ERR_UNSUPPORTED_ARCH = -5, /* device is not sm_100 */
  DTTS_ERR_CUDA = -6,
  DTTS_ERR_ALIGNMENT = -7
} dtts_status;

/* One entry of the weight table: `name` is the reference checkpoint key after weight-norm folding
 * (e.g. "fvae.decoder.wn.in_layers.0.weight"), `offset` counts floats from the arena base. */
typedef struct dtts_weight_entry {
  const char* name;
  uint64_t offset;
  uint64_t numel;
} dtts_weight_entry;

/* Hyper-parameters of the acoustic model (egs/datasets/audio/biaobei/dict_tts.yaml, resolved). */
typedef struct dtts_acoustic_desc {
  int32_t hidden, n_heads, enc_layers, ffn_kernel, ffn_filter, dict_dim;
  int32_t word_size, pinyin_size;
  int32_t dur_layers, dur_kernel, dur_chans;
  int32_t frames_multiple, latent;
  int32_t dec_layers, dec_kernel;
  int32_t flow_hidden, flow_kernel, flow_blocks, flow_layers;
  int32_t n_mel;
  int32_t language_zh; /* 1: apply add_pron_rule (layers/utils.py:109-115) */
  int32_t precision;   /* dense convolutions (encoder QKV/O/FFN, S2PA projections, duration predictor, WaveNet stacks):
                          0 = fp32 FMA pipe (exact), 1 = tcgen05 with bf16 hi/lo split operands (3 MMAs, fp32-class) */
  int32_t s2pa_route;  /* S2PAAttention (layers/dict_encoder.py:32-66) over the gloss tokens:
                          0 = folded streaming pass: logits = keys.(W_k^T q), context = W_o W_v (sum_l w_l values_l) --
                              algebraically identical, 0.3 % of the FLOPs, one HBM pass over keys/values (default);
                          1 = as the reference computes it: k = W_k keys, v = W_v values for EVERY gloss token as one
                              [B*Tw*Lk, dict_dim] x [dict_dim, 2*hidden] GEMM on tcgen05 (needs precision = 1), then
                              per-character scores / softmax / weighted sum over the projected rows.  Not available
                              with the dictionary bank (dtts_text_encode_bank). */
  int32_t model;       /* DTTS_MODEL_DICT (0): PortaSpeech_dict, the fields above.  DTTS_MODEL_PORTASPEECH (1): the non-dict
                          sibling (modules/portaspeech/model.py:132-366): phoneme encoder with
                          relative-position attention + FFT-block word encoder + word-to-phoneme attention in front of
                          the SAME duration predictor / length regulator / FVAE decoder; dict_dim, word_size, pinyin_size,
                          language_zh and s2pa_route are ignored, the three fields below are used */
  int32_t ph_size;          /* phoneme vocabulary (rows of ph_encoder.emb.weight)                                       */
  int32_t word_enc_layers;  /* FFT blocks of the word encoder (hparams word_enc_layers)                                 */
  int32_t rel_window;       /* relative-position window of the phoneme encoder's attention (4 upstream)                 */
} dtts_acoustic_desc;
#define DTTS_MODEL_DICT 0
#define DTTS_MODEL_PORTASPEECH 1

/* HiFi-GAN V1 generator description (egs/egs_bases/tts/vocoder/hifigan.yaml:3-10). */
#define DTTS_MAX_UPS 8
#define DTTS_MAX_RB 4
typedef struct dtts_vocoder_desc {
  int32_t n_mel, init_ch, n_ups, n_rb;
  int32_t up_rates[DTTS_MAX_UPS], up_kernels[DTTS_MAX_UPS];
  int32_t rb_kernels[DTTS_MAX_RB];
  int32_t rb_dilations[DTTS_MAX_RB][3];
  int32_t precision; /* 0 = fp32 FMA pipe; tcgen05 modes: 1 = bf16 hi/lo x hi/lo (3 MMAs, fp32-class accuracy),
                        2 = bf16 (1 MMA), 3 = fp16 activations x fp16 hi/lo weights (2 MMAs), 4 = fp16 (1 MMA),
                        5 = as 3, but layers with C_out >= 128 use one weight plane (1 MMA there; hi|lo stacked elsewhere),
                        6 = as 3, with the lo-plane correction of the C_out >= 128 ResBlock convolutions as an FP8 MMA
                            (e5m2 x e5m2, K = 32 per instruction: 1.5 instead of 2 units of tensor work, same accuracy) */
} dtts_vocoder_desc;

typedef struct dtts_acoustic dtts_acoustic; /* opaque */
typedef struct dtts_vocoder dtts_vocoder;   /* opaque */

int dtts_abi_version(void);
/* Message of the last failing call on this thread (never NULL). */
const char* dtts_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Acoustic model: replaces PortaSpeech_dict (modules/dict_tts/model.py:14-122) after
 * PortaSpeechFlowTask.test_start's weight-norm removal (tasks/tts/ps_flow.py:257-268).
 * ---------------------------------------------------------------------------------------------------------- */
int dtts_acoustic_create(const dtts_acoustic_desc* desc, const float* arena_dev, uint64_t arena_floats,
                         const dtts_weight_entry* table, int32_t n_entries, void* stream, dtts_acoustic** out);
int dtts_acoustic_destroy(dtts_acoustic* h);

/* Text side: DictEncoder.forward + duration predictor (model.py:84-97 -> dict_encoder.py:130-172,
 * portaspeech/model.py:58-66, model.py:64-82).  Shapes: B utterances, Tw word tokens, Lk gloss tokens per
 * character, Lp pinyin slots per character. */
typedef struct dtts_text_in {
  const int64_t* word_tokens_dev;   /* [B,Tw]                                                       */
  const int64_t* pron_modified_dev; /* [B,Tw] or NULL                                               */
  const float* keys_dev;            /* [B,Tw,Lk,dict_dim]                                           */
  const float* values_dev;          /* [B,Tw,Lk,dict_dim] (may alias keys_dev)                      */
  const float* key_map_dev;         /* [B,Tw,Lk] float, 0 = masked (dataset_utils.py:287-288)       */
  const int64_t* pinyin_dev;        /* [B,Tw,Lp]                                                    */
  const int64_t* pinyin_map_dev;    /* [B,Tw,Lp]                                                    */
  int32_t B, Tw, Lk, Lp;
} dtts_text_in;

typedef struct dtts_text_out {
  float* word_encoder_out_dev; /* [B,Tw,hidden]   ret['word_encoder_out']                    */
  float* dict_attn_dev;        /* [B,1,Lk,Tw]     ret['dict_attn']                           */
  float* pron_attn_dev;        /* [B,Tw,Lp]       ret['pron_attn']                           */
  float* dur_dev;              /* [B,Tw]          ret['dur'] (log-scale, softplus output)    */
  int64_t* dur_int_dev;        /* [B,Tw]          clamp(round(exp(dur)-1),0)                 */
  int64_t* ilens_dev;          /* [B]             (1-src_padding).sum(-1)                    */
} dtts_text_out;

uint64_t dtts_text_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tw, int32_t Lk, int32_t Lp);
int dtts_text_encode(dtts_acoustic* h, const dtts_text_in* in, const dtts_text_out* out, void* ws_dev,
                     uint64_t ws_bytes, void* stream);

/* GPU-resident dictionary bank.  keys / values / key_map / pinyin / pinyin_map of a character are a
 * pure function of its dictionary id (tasks/tts/dataset_utils.py:305-330 reads them from dict_embed.{data,idx} for
 * every batch), so the bank is uploaded once and a batch names its characters by id: entry i owns gloss rows
 * [tok_offsets[i], tok_offsets[i+1]) of keys/values/key_map and pronunciation slots [pin_offsets[i], pin_offsets[i+1])
 * of pinyin/pinyin_map.  dict_ids: >= 0 bank entry, -1 the BOS/EOS row the collater builds (keys 0, key_map 1,
 * pinyin 0, pinyin_map 1; dataset_utils.py:286-296), -2 padding.  Lk / Lp: padded widths of this batch (>= the longest
 * entry used; they are the Lk / Lp of the outputs dict_attn / pron_attn).  Results are identical to dtts_text_encode on
 * the tensors the collater would have built. */
typedef struct dtts_dict_bank {
  const float* keys_dev;          /* [rows, dict_dim]                       */
  const float* values_dev;        /* [rows, dict_dim] (may alias keys_dev)  */
  const float* key_map_dev;       /* [rows]                                 */
  const int64_t* tok_offsets_dev; /* [n_entries + 1]                        */
  const int64_t* pinyin_dev;      /* [slots]                                */
  const int64_t* pinyin_map_dev;  /* [slots]                                */
  const int64_t* pin_offsets_dev; /* [n_entries + 1]                        */
  int32_t n_entries;
} dtts_dict_bank;

typedef struct dtts_text_in_bank {
  const int64_t* word_tokens_dev;   /* [B,Tw]         */
  const int64_t* pron_modified_dev; /* [B,Tw] or NULL */
  const int64_t* dict_ids_dev;      /* [B,Tw]         */
  int32_t B, Tw, Lk, Lp;
} dtts_text_in_bank;

uint64_t dtts_text_bank_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tw, int32_t Lk, int32_t Lp);
int dtts_text_encode_bank(dtts_acoustic* h, const dtts_dict_bank* bank, const dtts_text_in_bank* in,
                          const dtts_text_out* out, void* ws_dev, uint64_t ws_bytes, void* stream);
/* Deferred status of the dictionary-bank gather.  dtts_text_encode_bank does not synchronise, so a dict_id >=
 * bank.n_entries (DTTS_ERR_BAD_ARG; the character was encoded as an all-zero row) or a bank entry wider than the call's
 * Lk / Lp (DTTS_ERR_BAD_SHAPE; it was truncated) is recorded in a pinned status word when the gather kernel has run.  It
 * is reported -- once -- by the next dtts_length_regulate_scan on the handle (the path's synchronising call) or by this
 * query: sync != 0 synchronises `stream` first, sync == 0 only looks at what has already landed.  The reference's own
 * behaviour for a bad index is a Python IndexError in the collater (tasks/tts/dataset_utils.py:312-330). */
int dtts_acoustic_status(dtts_acoustic* h, void* stream, int32_t sync);

/* Length regulator (modules/fastspeech/tts_modules.py:215-251).  Step 1 scans durations on the device and
 * returns the longest utterance in *t_raw_host -- this call SYNCHRONISES the stream (the one data-dependent shape
 * on the path).  cum_dev is int32 [B,Tw] scratch kept for step 2; totals_dev is int32 [B+1]
 * (per-utterance frame counts, then the batch maximum). */
int dtts_length_regulate_scan(dtts_acoustic* h, const int64_t* dur_int_dev, const int64_t* ilens_dev, int32_t B,
                              int32_t Tw, int32_t* cum_dev, int32_t* totals_dev, int32_t* t_raw_host, void* stream);
/* Step 2: mel2word [B,T] with T = t_raw rounded up to frames_multiple; the padded columns repeat the last
 * column (modules/dict_tts/model.py:98-100). */
int dtts_length_regulate_fill(dtts_acoustic* h, const int32_t* cum_dev, const int64_t* ilens_dev, int32_t B,
                              int32_t Tw, int32_t t_raw, int32_t T, int64_t* mel2word_dev, void* stream);
/* Zero-row pad + gather + x*tgt_nonpadding (model.py:101-107,53): decoder_inp [B,T,hidden], x_mask [B,T]. */
int dtts_expand(dtts_acoustic* h, const float* word_encoder_out_dev, const int64_t* mel2word_dev, int32_t B,
                int32_t Tw, int32_t T, float* decoder_inp_dev, float* g_bct_dev, float* x_mask_dev, void* stream);

/* FVAE_semantics.forward(infer=True) (modules/dict_tts/fvae_semantics.py:84-115): g_bct is decoder_inp in
 * [B,hidden,T] layout (second output of dtts_expand), z_in the N(0,1) prior sample [B,latent,T/4] drawn by the
 * host; writes mel [B,T,n_mel] and z_p [B,latent,T/4].  T must be a multiple of frames_multiple. */
uint64_t dtts_decode_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t T);
int dtts_decode_mel(dtts_acoustic* h, const float* g_bct_dev, const float* z_in_dev, int32_t B, int32_t T,
                    float* mel_dev, float* z_p_dev, void* ws_dev, uint64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * PortaSpeech (non-dict) sibling, SURVEY.md §8f-3: replaces PortaSpeech.run_text_encoder
 * (modules/portaspeech/model.py:239-262) on a handle created with desc.model = DTTS_MODEL_PORTASPEECH.  The duration
 * scan / fill (dtts_length_regulate_*) and dtts_decode_mel are the calls of the dict model.
 *   dtts_ps_text_encode: TextEncoder (model.py:69-129: embedding, ConvReluNorm pre-net, post-LN encoder with
 *     relative-position attention, rel_transformer_encoder.py:26-247) * src_nonpadding; group_hidden_by_segs
 *     (portaspeech/utils.py:3-16) + FFT-block word encoder (fastspeech/tts_modules.py:458-566,
 *     commons/common_layers.py:624-673); phoneme-level durations summed per word (model.py:317-340).
 *   dtts_ps_attend: in-word sinusoidal positions (model.py:359-363), gather by mel2word, enc_pos_proj / dec_query_proj /
 *     dec_res_proj and the one-head word-to-phoneme attention whose mask lets a frame see only the phonemes of its own
 *     word (model.py:304-315), * tgt_nonpadding: the decoder input.
 * Shapes: B utterances, Tp phonemes, Tw words (= word_len.max()), T mel frames (a multiple of frames_multiple).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct dtts_ps_text_in {
  const int64_t* txt_tokens_dev; /* [B,Tp] phoneme ids, 0 = padding                      */
  const int64_t* ph2word_dev;    /* [B,Tp] 1-based word index of every phoneme, 0 = pad  */
  int32_t B, Tp, Tw;
} dtts_ps_text_in;

typedef struct dtts_ps_text_out {
  float* ph_encoder_out_dev;   /* [B,Tp,hidden]  ret['ph_encoder_out']                                     */
  float* word_encoder_out_dev; /* [B,Tw,hidden]  ret['word_encoder_out']                                   */
  float* dur_dev;              /* [B,Tw]         ret['dur']: per-word sum of the phoneme-level predictions */
  int64_t* dur_int_dev;        /* [B,Tw]         clamp(round(exp(dur)-1),0)                                */
  int64_t* ilens_dev;          /* [B]            (1-src_padding).sum(-1) over PHONEMES, as add_dur passes it */
} dtts_ps_text_out;

uint64_t dtts_ps_text_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tp, int32_t Tw);
int dtts_ps_text_encode(dtts_acoustic* h, const dtts_ps_text_in* in, const dtts_ps_text_out* out, void* ws_dev,
                        uint64_t ws_bytes, void* stream);
uint64_t dtts_ps_attend_workspace_bytes(const dtts_acoustic* h, int32_t B, int32_t Tp, int32_t Tw, int32_t T);
/* attn_dev (optional): [B,T,Tp] attention weights, ret['attn'].  decoder_inp [B,T,hidden], g_bct [B,hidden,T] (the
 * layout dtts_decode_mel reads), x_mask [B,T] = mel2word > 0. */
int dtts_ps_attend(dtts_acoustic* h, const float* ph_encoder_out_dev, const float* word_encoder_out_dev,
                   const int64_t* ph2word_dev, const int64_t* mel2word_dev, int32_t B, int32_t Tp, int32_t Tw, int32_t T,
                   float* attn_dev, float* decoder_inp_dev, float* g_bct_dev, float* x_mask_dev, void* ws_dev,
                   uint64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Vocoder: replaces HifiGanGenerator.forward behind BaseVocoder.spec2wav (vocoders/hifigan.py:54-62).
 * mel [B,T,n_mel] -> wav [B, T*hop].
 * ---------------------------------------------------------------------------------------------------------- */
int dtts_vocoder_create(const dtts_vocoder_desc* desc, const float* arena_dev, uint64_t arena_floats,
                        const dtts_weight_entry* table, int32_t n_entries, void* stream, dtts_vocoder** out);
int dtts_vocoder_destroy(dtts_vocoder* h);
uint64_t dtts_vocode_workspace_bytes(const dtts_vocoder* h, int32_t B, int32_t T);
int dtts_vocode(dtts_vocoder* h, const float* mel_dev, int32_t B, int32_t T, float* wav_dev, void* ws_dev,
                uint64_t ws_bytes, void* stream);
/* The same with the valid length of every item (extension; SURVEY.md §8b: spec2wav_batch(mel, lengths)).
 * lens_dev: int32 [B], valid mel frames per item (0 <= lens[b] <= T); more than 512 items run as consecutive
 * sub-batches of 512 on the same workspace.  wav[b, t] for t < lens[b]*hop is bit
 * for bit what dtts_vocode writes (the frames after lens[b] still feed the receptive field of the last valid samples,
 * exactly as in the full-length call); samples past lens[b]*hop are written as 0 and the rows only they depend on are
 * not computed -- DictTTSTask.after_infer never looks at them (B = 1 upstream; here the task trims to the valid length). */
int dtts_vocode_lens(dtts_vocoder* h, const float* mel_dev, const int32_t* lens_dev, int32_t B, int32_t T,
                     float* wav_dev, void* ws_dev, uint64_t ws_bytes, void* stream);
/* Kernels launched by this handle since creation (for bench.py's gpu_launches claim). */
uint64_t dtts_vocoder_launch_count(const dtts_vocoder* h);
uint64_t dtts_acoustic_launch_count(const dtts_acoustic* h);

/* ------------------------------------------------------------------------------------------------------------
 * after_infer on the device (replaces the host work of DictTTSTask.after_infer,
 * tasks/tts/dict_tts.py:227-311 and utils/audio.py:11-16).
 * dtts_wav_to_pcm16: pcm = (int16)(wav * 32767) truncated toward zero (numpy astype semantics); n_samples % 4 == 0.
 * dtts_pron_tokens:  pairs[b,t,:] = pinyin[b,t, i : i+2] with i = first argmax of pron_attn[b,t,:] (-1 past the end);
 *                    pinyin is the explicit [B,Tw,Lp] tensor, or NULL with (bank, dict_ids) naming the characters.
 * ---------------------------------------------------------------------------------------------------------- */
int dtts_wav_to_pcm16(const float* wav_dev, uint64_t n_samples, int16_t* pcm_dev, void* stream);
int dtts_pron_tokens(const float* pron_attn_dev, const int64_t* pinyin_dev, const dtts_dict_bank* bank,
                     const int64_t* dict_ids_dev, int32_t B, int32_t Tw, int32_t Lp, int64_t* pairs_dev, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Unit-test hook: the generic fp32 convolution kernel on raw tensors (torch.nn.functional.conv1d /
 * conv_transpose1d semantics, weights in PyTorch layout).  x [B,C_in,T_in] -> out [B,C_out,T_out].
 * ---------------------------------------------------------------------------------------------------------- */
int dtts_debug_conv1d(const float* x_dev, const float* w_dev, const float* bias_dev, float* out_dev, int32_t B,
                      int32_t C_in, int32_t T_in, int32_t C_out, int32_t K, int32_t stride, int32_t padding,
                      int32_t dilation, int32_t transposed, float pre_slope, float* scratch_w_dev, void* stream);

/* Unit-test hook for the tcgen05 tensor-core convolution of the HiFi-GAN stack (stride-1 Conv1d, or
 * ConvTranspose1d with K % stride == 0 and padding == (K-stride)/2).  x [B,C_in,T_in] -> out [B,C_out,T_out] =
 * post * (conv(leaky(x, pre_slope)) + bias + res); out_act (optional) = leaky(out, act_slope) as it is handed to the next
 * layer (hi [+ lo] 16-bit planes, reconstructed to fp32).  precision: 1..4 as dtts_vocoder_desc.precision.
 * scratch_dev: caller-provided device scratch. */
int dtts_debug_tc_conv1d(const float* x_dev, const float* w_dev, const float* bias_dev, const float* res_dev,
                         float* out_dev, float* out_act_dev, int32_t B, int32_t C_in, int32_t T_in, int32_t C_out,
                         int32_t K, int32_t stride, int32_t padding, int32_t dilation, int32_t transposed,
                         float pre_slope, float post, float act_slope, int32_t precision, void* scratch_dev,
                         uint64_t scratch_bytes, void* stream);

/* Unit-test hook: the fused ResBlock pair kernel (conv1 -> leaky -> conv2 -> + residual of the C = 32 stage in one
 * launch, intermediate in shared memory) is bit-identical to the two launches it replaces; this switch lets a test run
 * both.  mode 0: two launches per pair; 2: fused pairs, conv_post as its own kernel (bit-identical to mode 0); 1: fused
 * pairs with conv_post folded into the last one (the default; other summation order inside conv_post, ~1e-7);
 * 3: as 2, and every ResBlock pair of the C = 128 stage fused as well (by default only its k = 3 pairs are: the others
 * measured no faster than two launches, DTTS_TC_FUSE128_MAXK); 4: as 2, and the whole k = 3 ResBlock of the C = 32 stage
 * (three pairs) as ONE launch, bit-identical again (part of the default); -1: back to the default / the DTTS_TC_FUSE,
 * DTTS_TC_FOLD_POST, DTTS_TC_FUSE_BLOCK environment switches. */
int dtts_debug_set_tc_fuse(int32_t mode);

/* Unit-test hook: the fused kernels of the acoustic model (flow_fused_kernel: the whole reverse prior flow of
 * FVAE_semantics.forward(infer=True), modules/dict_tts/fvae_semantics.py:96-101, as one launch instead of one launch per
 * layer) compute the same arithmetic as the per-layer path in another summation order; this switch lets a test run both.
 * mode 0: one launch per layer; 1: fused; -1: back to the default (fused) / the DTTS_AC_FUSE environment switch. */
int dtts_debug_set_acoustic_fuse(int32_t mode);

#ifdef __cplusplus
}
#endif
#endif /* DTTS_H_ */
