"""GPU parity at BASELINE.json's full sizes (-m gpu).  The oracle cannot run a whole cfg-2 batch in seconds, so these
tests use size-independent properties (an utterance computed inside the batch == the same utterance computed alone,
bit for bit; the expand is an index copy; mel2word is sorted) plus oracle spot checks on single utterances:

  cfg 1  one 64-character utterance, ~1 280 frames, text -> mel -> wav against the oracle end to end
  cfg 2  batch 60, <= 22 word tokens, L_k <= 96, 400 frames (the bench workload)
  cfg 4  vocoder only, 256 segments of 32 frames
"""
import pytest
import torch

from dict_tts_b200 import synth
from dict_tts_b200.config import AcousticConfig, VocoderConfig
from dict_tts_b200.weights import fold_weight_norm
from oracle import dtts_oracle as O
from tests.cases import ACOUSTIC_SEED, TOL_MEL_MAXABS, TOL_WAV_RMS, VOCODER_SEED

pytestmark = pytest.mark.gpu

DICT_KEYS = ("keys", "values", "key_map", "pinyin", "pinyin_map")


def _forward(eng, batch, rows=None):
    sel = (lambda t: t) if rows is None else (lambda t: t[rows].contiguous())
    return eng.forward((sel(batch["word_tokens"]),), sel(batch["pron_modified"]),
                       dict_msg=tuple(sel(batch[k]) for k in DICT_KEYS), mel2word=sel(batch["mel2word"]),
                       z_p=sel(batch["z_p"]))


@pytest.fixture(scope="module")
def engines():
    from dict_tts_b200.engine import DictTTSEngine, HifiGanEngine
    asd, vsd = synth.make_acoustic_state_dict(ACOUSTIC_SEED), synth.make_vocoder_state_dict(VOCODER_SEED)
    eng, voc = DictTTSEngine(asd), HifiGanEngine(vsd)
    yield eng, voc, fold_weight_norm(asd), fold_weight_norm(vsd)
    eng.close()
    voc.close()


def test_cfg2_batch_properties_and_oracle_spot_check(engines):
    eng, voc, W, Wv = engines
    batch = synth.make_batch(seed=1234, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96)
    out = _forward(eng, batch)
    wav = voc(out["mel_out"])
    torch.cuda.synchronize()
    B, T = batch["mel2word"].shape
    assert out["mel_out"].shape == (B, T, 80) and wav.shape == (B, T * 256)
    assert torch.isfinite(out["mel_out"]).all() and torch.isfinite(wav).all() and wav.abs().max() <= 1.0
    # the supplied alignment comes back unchanged; it is sorted inside the valid region and zero after it
    m2w = out["mel2word"].cpu()
    assert torch.equal(m2w, batch["mel2word"])
    for b in range(B):
        n = int((m2w[b] > 0).sum())
        assert (m2w[b, 1:n] >= m2w[b, :n - 1]).all() and (m2w[b, n:] == 0).all()
    # the expand is an index copy of encoder rows
    x, nonpad = O.expand_by_mel2word(out["word_encoder_out"].cpu(), m2w)
    assert torch.equal(out["decoder_inp"].cpu(), x) and torch.equal(out["x_mask"].cpu(), nonpad)
    # dict_attn is a distribution over the gloss tokens of every real character; pron_attn sums to 1 over its slots
    da = out["dict_attn"].cpu()[:, 0]                                 # [B, Lk, Tw]
    real = batch["word_tokens"] > 1
    assert torch.allclose(da.sum(1)[real], torch.ones(int(real.sum())), atol=1e-5)
    masked = (batch["key_map"] == 0) & real[:, :, None]
    assert da.transpose(1, 2)[masked].abs().max() == 0                # -1e9 logits -> exactly zero weight
    # an utterance inside the batch == the same utterance alone (same padded widths), bit for bit, through the vocoder
    for b in (0, 17, 59):
        solo = _forward(eng, batch, [b])
        for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "mel_out"):
            assert torch.equal(solo[k][0], out[k][b]), (k, b)
        assert torch.equal(voc(solo["mel_out"])[0], wav[b]), b
    # oracle on one utterance of the batch (text -> mel -> wav), north-star tolerances
    b = 41
    sub = {k: (v[b:b + 1].contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}
    with torch.no_grad():
        ref = O.acoustic_forward(W, AcousticConfig(), sub, sub["mel2word"], sub["z_p"])
        ref_wav = O.hifigan_forward(Wv, VocoderConfig(), ref["mel_out"])
    assert (out["mel_out"][b].cpu() - ref["mel_out"][0]).abs().max() < TOL_MEL_MAXABS
    assert (out["dict_attn"][b].cpu() - ref["dict_attn"][0]).abs().max() < 1e-5
    assert (wav[b].cpu() - ref_wav[0]).pow(2).mean().sqrt() < TOL_WAV_RMS


def test_cfg2_predicted_durations_match_oracle_bit_exactly(engines):
    """The data-dependent path at batch 60: durations -> mel2word on the device against the oracle's length regulator."""
    eng, _, W, _ = engines
    batch = synth.make_batch(seed=77, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96)
    t = eng.text_encode(batch["word_tokens"], batch["pron_modified"], *[batch[k] for k in DICT_KEYS])
    m2w = eng.length_regulate(t["dur_int"], t["ilens"]).cpu()
    want = O.length_regulate(t["dur_int"].cpu(), t["ilens"].cpu())
    want = want[0] if isinstance(want, (tuple, list)) else want
    assert m2w.dtype == torch.int64 and torch.equal(m2w, want)
    # and the integer durations themselves against the oracle's duration predictor on four utterances
    rows = [0, 13, 30, 59]
    sub = {k: batch[k][rows].contiguous() for k in ("word_tokens", "pron_modified") + DICT_KEYS}
    with torch.no_grad():
        enc, _, _, _ = O.text_encode(W, AcousticConfig(), sub["word_tokens"], sub["pron_modified"],
                                     *[sub[k] for k in DICT_KEYS])
        dur, _ = O.duration_predictor(W, AcousticConfig(), enc)
    assert torch.equal(O.durations_to_int(dur), t["dur_int"].cpu()[rows])


def test_cfg1_single_long_utterance_against_oracle(engines):
    eng, voc, W, Wv = engines
    batch = synth.make_batch(seed=5, B=1, min_chars=64, max_chars=64, max_frames=1280, Lk_cap=96)
    assert batch["word_tokens"].shape == (1, 66)
    out = _forward(eng, batch)
    wav = voc(out["mel_out"])
    with torch.no_grad():
        ref = O.acoustic_forward(W, AcousticConfig(), batch, batch["mel2word"], batch["z_p"])
        ref_wav = O.hifigan_forward(Wv, VocoderConfig(), ref["mel_out"])
    assert wav.shape == (1, 1280 * 256)
    assert (out["mel_out"].cpu() - ref["mel_out"]).abs().max() < TOL_MEL_MAXABS
    assert (out["pron_attn"].cpu() - ref["pron_attn"]).abs().max() < 1e-5
    assert (wav.cpu() - ref_wav).pow(2).mean().sqrt() < TOL_WAV_RMS


def test_cfg4_vocoder_256_segments(engines):
    _, voc, _, Wv = engines
    mel = synth.make_mel(9, 256, 32)
    wav = voc(mel)
    assert wav.shape == (256, 32 * 256) and torch.isfinite(wav).all()
    for b in (0, 100, 255):                                           # a segment inside the batch == the segment alone
        assert torch.equal(voc(mel[b:b + 1])[0], wav[b]), b
    with torch.no_grad():
        ref = O.hifigan_forward(Wv, VocoderConfig(), mel[100:102])
    assert (wav[100:102].cpu() - ref).pow(2).mean().sqrt() < TOL_WAV_RMS
