// after_infer on the device: what DictTTSTask.after_infer does on the host after the vocoder
// (tasks/tts/dict_tts.py:227-311, utils/audio.py:11-16) -- float waveform -> int16 PCM, and argmax over the
// pronunciation attention -> the two pinyin token ids of each character (the `pinyin_tokens` column of meta.csv that
// scripts/get_pron_error.py scores) -- so that only int16 samples and [B,Tw,2] ids cross the bus.
#include "engine.cuh"

namespace dtts {

// utils/audio.py:15-16: (wav * 32767).astype(np.int16) -- numpy's float -> int16 cast truncates toward zero
__global__ void wav_to_pcm16_kernel(const float* __restrict__ wav, int16_t* __restrict__ pcm, size_t n4) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(wav)[i];
  short4 o;
  o.x = (short)__float2int_rz(v.x * 32767.f);
  o.y = (short)__float2int_rz(v.y * 32767.f);
  o.z = (short)__float2int_rz(v.z * 32767.f);
  o.w = (short)__float2int_rz(v.w * 32767.f);
  reinterpret_cast<short4*>(pcm)[i] = o;
}

// dict_tts.py:295-304: idx = pron_attn[b,t].max(-1)[1] (first maximum); tokens = pinyin[b,t][idx : idx+2]
// (-1 where the slice runs past Lp).  pinyin: explicit [B,Tw,Lp] tensor, or the bank (padded with 0 like the collater).
__global__ void pron_argmax_kernel(const float* __restrict__ pron_attn, const int64_t* __restrict__ pinyin,
                                   const int64_t* __restrict__ dict_ids, const int64_t* __restrict__ pin_off,
                                   const int64_t* __restrict__ bank_pinyin, int n_entries, int n, int Lp,
                                   int64_t* __restrict__ pairs) {
  const int bt = blockIdx.x * blockDim.x + threadIdx.x;
  if (bt >= n) return;
  const float* a = pron_attn + (size_t)bt * Lp;
  int best = 0;
  float bv = a[0];
  for (int p = 1; p < Lp; ++p)
    if (a[p] > bv) { bv = a[p]; best = p; }
  for (int k = 0; k < 2; ++k) {
    const int p = best + k;
    int64_t tok = -1;
    if (p < Lp) {
      if (pinyin) {
        tok = pinyin[(size_t)bt * Lp + p];
      } else {
        tok = 0;
        const int64_t id = dict_ids[bt];
        if (id >= 0 && id < n_entries) {
          const int64_t p0 = pin_off[id], np = pin_off[id + 1] - p0;
          if (p < np) tok = bank_pinyin[p0 + p];
        }
      }
    }
    pairs[(size_t)bt * 2 + k] = tok;
  }
}

}  // namespace dtts

using namespace dtts;

extern "C" int dtts_wav_to_pcm16(const float* wav_dev, uint64_t n_samples, int16_t* pcm_dev, void* stream) {
  if (!wav_dev || !pcm_dev) return fail(DTTS_ERR_BAD_ARG, "dtts_wav_to_pcm16: null argument");
  if (n_samples % 4 || ((uintptr_t)wav_dev & 15) || ((uintptr_t)pcm_dev & 7))
    return fail(DTTS_ERR_ALIGNMENT, "dtts_wav_to_pcm16: sample count must be a multiple of 4 and buffers aligned");
  if (!n_samples) return DTTS_OK;
  const size_t n4 = n_samples / 4;
  wav_to_pcm16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wav_dev, pcm_dev, n4);
  DTTS_CUDA(cudaGetLastError());
  return DTTS_OK;
}

extern "C" int dtts_pron_tokens(const float* pron_attn_dev, const int64_t* pinyin_dev, const dtts_dict_bank* bank,
                                const int64_t* dict_ids_dev, int32_t B, int32_t Tw, int32_t Lp, int64_t* pairs_dev,
                                void* stream) {
  if (!pron_attn_dev || !pairs_dev) return fail(DTTS_ERR_BAD_ARG, "dtts_pron_tokens: null argument");
  if (!pinyin_dev && (!bank || !dict_ids_dev || !bank->pin_offsets_dev || !bank->pinyin_dev))
    return fail(DTTS_ERR_BAD_ARG, "dtts_pron_tokens: need either the pinyin tensor or the bank + dict_ids");
  if (B <= 0 || Tw <= 0 || Lp <= 0) return fail(DTTS_ERR_BAD_SHAPE, "dtts_pron_tokens: empty shape");
  const int n = B * Tw;
  pron_argmax_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      pron_attn_dev, pinyin_dev, dict_ids_dev, bank ? bank->pin_offsets_dev : nullptr, bank ? bank->pinyin_dev : nullptr,
      bank ? bank->n_entries : 0, n, Lp, pairs_dev);
  DTTS_CUDA(cudaGetLastError());
  return DTTS_OK;
}
