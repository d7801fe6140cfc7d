"""GPU parity suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle and the
reference-generated golden fixtures.  Tolerances are the north star's (tests/cases.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dict_tts_b200 import binding, synth
from dict_tts_b200.config import AcousticConfig, VocoderConfig
from dict_tts_b200.weights import fold_weight_norm
from oracle import dtts_oracle as O
from tests.cases import (ACOUSTIC_CASES, ACOUSTIC_SEED, TOL_MEL_MAXABS, TOL_WAV_RMS, VOCODER_CASES, VOCODER_SEED)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[(0, 0), (1, 0), (1, 1)], ids=["fp32", "tcgen05", "tcgen05-s2pa_gemm"])
def acoustic(request):
    """Both acoustic precisions: 0 = every convolution on the fp32 FMA pipe, 1 = dense convolutions on tcgen05 with
    bf16 hi/lo split operands (the default of the engine); and, on tcgen05, both S2PA routes: 0 = folded streaming pass,
    1 = K/V projection of every gloss token as one GEMM (dtts_acoustic_desc.s2pa_route)."""
    from dict_tts_b200.engine import DictTTSEngine
    sd = synth.make_acoustic_state_dict(ACOUSTIC_SEED)
    eng = DictTTSEngine(sd, precision=request.param[0], s2pa_route=request.param[1])
    yield eng, fold_weight_norm(sd)
    eng.close()


@pytest.fixture(scope="module")
def vocoder():
    from dict_tts_b200.engine import HifiGanEngine
    sd = synth.make_vocoder_state_dict(VOCODER_SEED)
    eng = HifiGanEngine(sd, precision=0)            # exact fp32 path; the tcgen05 modes are in test_gpu_tensorcore.py
    yield eng, fold_weight_norm(sd)
    eng.close()


def _run_acoustic(eng, batch, predicted, z):
    return eng.forward((batch["word_tokens"], batch["word_tokens"]), batch["pron_modified"], (None, None, None),
                       ph2word=None, word_len=batch["word_lengths"].max(),
                       dict_msg=(batch["keys"], batch["values"], batch["key_map"], batch["pinyin"], batch["pinyin_map"]),
                       infer=True, forward_post_glow=False, two_stage=True,
                       mel2word=None if predicted else batch["mel2word"], z_p=z)


# --------------------------------------------------------------------------------------------------
# generic convolution kernel vs torch (CPU fp32) on the shapes the model uses
# --------------------------------------------------------------------------------------------------
CONV_SHAPES = [
    # B, C_in, T_in, C_out, K, stride, pad, dil, transposed, pre_slope
    (2, 80, 37, 512, 7, 1, 3, 1, 0, 1.0),       # conv_pre
    (2, 512, 19, 256, 16, 8, 4, 1, 1, 0.1),     # ups.0
    (1, 64, 300, 32, 4, 2, 1, 1, 1, 0.1),       # ups.3
    (2, 128, 700, 128, 11, 1, 25, 5, 0, 0.1),   # resblock k11 d5
    (1, 32, 3000, 32, 3, 1, 3, 3, 0, 0.1),      # last stage
    (1, 32, 2500, 1, 7, 1, 3, 1, 0, 0.01),      # conv_post (no tanh here)
    (3, 192, 22, 768, 5, 1, 2, 1, 0, 1.0),      # encoder FFN
    (3, 192, 100, 192, 8, 4, 2, 1, 0, 1.0),     # g_pre_net
    (2, 16, 25, 192, 4, 4, 0, 1, 1, 1.0),       # FVAE decoder pre_net
    (2, 64, 25, 8, 1, 1, 0, 1, 0, 1.0),         # flow post
    (1, 192, 1, 192, 1, 1, 0, 1, 0, 1.0),       # single position
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
def test_conv1d_kernel_matches_torch(shape):
    B, Ci, Ti, Co, K, s, p, d, tr, slope = shape
    lib = binding.load()
    g = torch.Generator().manual_seed(sum(shape[:9]))
    x = torch.randn(B, Ci, Ti, generator=g)
    w = torch.randn((Ci, Co, K) if tr else (Co, Ci, K), generator=g) / (Ci * K) ** 0.5
    b = torch.randn(Co, generator=g)
    xin = F.leaky_relu(x, slope) if slope != 1.0 else x
    ref = F.conv_transpose1d(xin, w, b, stride=s, padding=p) if tr else F.conv1d(xin, w, b, stride=s, padding=p,
                                                                                  dilation=d)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = torch.full(ref.shape, float("nan"), device="cuda")
    scratch = torch.empty(w.numel(), device="cuda")
    rc = lib.dtts_debug_conv1d(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), B, Ci, Ti, Co, K, s, p, d,
                               tr, C.c_float(slope), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    binding.check(rc, "debug_conv1d")
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    assert err < 2e-5 * max(1.0, ref.abs().max().item()), err


# --------------------------------------------------------------------------------------------------
# acoustic model: stage by stage vs golden (reference outputs) and oracle
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(ACOUSTIC_CASES))
def test_acoustic_matches_reference_golden(name, golden_dir, acoustic):
    eng, _ = acoustic
    kw, predicted = ACOUSTIC_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    batch = synth.make_batch(**kw)
    out = _run_acoustic(eng, batch, predicted, torch.from_numpy(gold["z_in"]))
    torch.cuda.synchronize()
    assert np.array_equal(out["mel2word"].cpu().numpy(), gold["mel2word"]), "mel2word must be bit-exact"
    tol = dict(word_encoder_out=2e-4, dict_attn=1e-5, pron_attn=1e-5, dur=1e-4, z_p=1e-4, mel_out=TOL_MEL_MAXABS)
    for k, t in tol.items():
        err = np.abs(out[k].cpu().numpy() - gold[k]).max()
        assert err < t, (k, err)
    # the gather is an index copy: decoder_inp must equal word_encoder_out rows bit for bit
    enc = out["word_encoder_out"].cpu()
    x, nonpad = O.expand_by_mel2word(enc, out["mel2word"].cpu())
    assert torch.equal(out["decoder_inp"].cpu(), x)
    assert torch.equal(out["x_mask"].cpu(), nonpad)


def test_length_regulator_bit_exact_edge_cases(acoustic):
    eng, _ = acoustic
    g = torch.Generator().manual_seed(3)
    B, Tw = 7, 37
    dur = torch.randint(0, 9, (B, Tw), generator=g)
    ilens = torch.tensor([37, 1, 5, 36, 20, 3, 33])
    dur[1] = 0            # all-zero row -> ones
    dur[2, :5] = 0        # all zero inside ilen, garbage after
    dur[5, 1] = 0
    want = O.length_regulate(dur, ilens, 4)
    got = eng.length_regulate(dur.cuda(), ilens.cuda())
    assert torch.equal(got.cpu(), want)
    # long sequence, single utterance
    dur = torch.randint(0, 30, (1, 1500), generator=g)
    want = O.length_regulate(dur, torch.tensor([1500]), 4)
    got = eng.length_regulate(dur.cuda(), torch.tensor([1500]).cuda())
    assert torch.equal(got.cpu(), want)


def test_decoder_only_matches_oracle(acoustic):
    eng, W = acoustic
    cfg = AcousticConfig()
    g = torch.Generator().manual_seed(9)
    B, T = 2, 52
    dec_in = torch.randn(B, T, cfg.hidden, generator=g) * 0.5
    z = synth.draw_z(B, cfg.latent, T // 4, 77)
    with torch.no_grad():
        want, zp = O.decode_mel(W, cfg, dec_in, z)
    mel, z_out = eng.decode_mel(dec_in.transpose(1, 2).contiguous().cuda(), z)
    assert (z_out.cpu() - zp).abs().max() < 1e-4
    assert (mel.cpu() - want).abs().max() < TOL_MEL_MAXABS


# --------------------------------------------------------------------------------------------------
# vocoder
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(VOCODER_CASES))
def test_vocoder_matches_reference_golden(name, golden_dir, vocoder):
    eng, _ = vocoder
    kw = VOCODER_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))["wav"]
    wav = eng(synth.make_mel(kw["seed"], kw["B"], kw["T"])).cpu().numpy()
    rms = float(np.sqrt(np.mean((wav - gold) ** 2)))
    assert rms < TOL_WAV_RMS, rms
    assert np.abs(wav - gold).max() < 1e-3


def test_vocoder_batch_equals_single(vocoder):
    """Utterances are independent: batching must not change any sample."""
    eng, _ = vocoder
    mel = synth.make_mel(5, 3, 40)
    full = eng(mel)
    for b in range(3):
        one = eng(mel[b:b + 1])
        assert torch.equal(one[0], full[b])


def test_bad_arguments_raise(acoustic, vocoder):
    eng, _ = acoustic
    with pytest.raises(RuntimeError):
        eng.decode_mel(torch.zeros(1, 192, 6, device="cuda"), torch.zeros(1, 16, 1))      # T % 4 != 0 -> BAD_SHAPE
    v, _ = vocoder
    with pytest.raises(ValueError):
        v(torch.zeros(1, 10, 79))


def test_s2pa_gemm_route_agrees_with_folded_route_and_oracle():
    """dtts_acoustic_desc.s2pa_route = 1 (k = W_k keys, v = W_v values for every gloss token on tcgen05, then scores /
    softmax / weighted sum -- dict_encoder.py:40-58 as written) against route 0 and the oracle, with values != keys
    (two staging passes + two projection launches) and with values aliasing keys (one GEMM for k | v)."""
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.engine import DictTTSEngine
    sd = synth.make_acoustic_state_dict(ACOUSTIC_SEED)
    W = fold_weight_norm(sd)
    folded = DictTTSEngine(sd, precision=1, s2pa_route=0)
    gemm = DictTTSEngine(sd, precision=1, s2pa_route=1)
    batch = synth.make_batch(seed=91, B=3, min_chars=2, max_chars=8, max_frames=48, Lk_cap=72, pron_modified_p=0.1)
    g = torch.Generator().manual_seed(5)
    other = batch["values"] + 0.1 * torch.randn(batch["values"].shape, generator=g) * (batch["values"] != 0)
    for values in (other, None):
        vals = batch["keys"] if values is None else values
        a = folded.text_encode(batch["word_tokens"], batch["pron_modified"], batch["keys"], values, batch["key_map"],
                               batch["pinyin"], batch["pinyin_map"])
        b = gemm.text_encode(batch["word_tokens"], batch["pron_modified"], batch["keys"], values, batch["key_map"],
                             batch["pinyin"], batch["pinyin_map"])
        with torch.no_grad():
            enc, dict_attn, pron_attn, _ = O.text_encode(W, AcousticConfig(), batch["word_tokens"], batch["pron_modified"],
                                                         batch["keys"], vals, batch["key_map"], batch["pinyin"],
                                                         batch["pinyin_map"])
        ref = dict(word_encoder_out=enc, dict_attn=dict_attn, pron_attn=pron_attn)
        for k, tol in (("word_encoder_out", 2e-4), ("dict_attn", 1e-5), ("pron_attn", 1e-5), ("dur", 1e-4)):
            assert (a[k] - b[k]).abs().max().item() < tol, (k, "route 0 vs 1")
            if k in ref:
                assert (b[k].cpu() - ref[k]).abs().max().item() < tol, (k, "route 1 vs oracle")
        assert torch.equal(a["ilens"], b["ilens"])
    bank, ids = DictBank.from_batch(batch)
    gemm.set_dict_bank(bank)
    with pytest.raises(RuntimeError, match="dictionary bank"):
        gemm.text_encode_bank(batch["word_tokens"], batch["pron_modified"], ids)
    folded.close()
    gemm.close()
    with pytest.raises(RuntimeError, match="s2pa_route"):
        DictTTSEngine(sd, precision=0, s2pa_route=1)


SWEEP = [
    # seed, B, min_chars, max_chars, max_frames, Lk_cap, note
    (301, 1, 1, 1, 9, 8, "one character, T = 9 (padded to 12 by repeating the last column)"),
    (302, 2, 1, 3, 18, 12, "tiny gloss lists, T % 4 == 2"),
    (303, 5, 2, 11, 61, 33, "odd everything"),
    (304, 3, 20, 40, 200, 48, "Tw = 42: still the shared-memory attention"),
    (305, 2, 70, 90, 400, 24, "Tw = 92: shared-memory attention with 128 KB of q | k | v"),
    (307, 1, 131, 140, 560, 8, "Tw = 142 > 128: the general attention kernel, LayerNorm as its own launches"),
    (306, 7, 1, 6, 28, 96, "many short utterances, long gloss lists"),
]


@pytest.mark.parametrize("case", SWEEP, ids=lambda c: "seed%d" % c[0])
def test_acoustic_random_shapes_against_oracle(case, acoustic):
    """Shapes the fixtures do not hold, every stage against the oracle run on the same batch (supplied alignment), and the
    predicted-duration path bit-exactly.  pron_modified = None must equal an all-zero pron_modified."""
    eng, W = acoustic
    seed, B, cmin, cmax, frames, lk, _ = case
    T4 = (frames + 3) // 4 * 4
    batch = synth.make_batch(seed=seed, B=B, min_chars=cmin, max_chars=cmax, max_frames=T4, Lk_cap=lk,
                             pron_modified_p=0.05)
    batch["mel2word"] = batch["mel2word"][:, :frames].contiguous()      # T % 4 != 0: both sides repeat the last column
    z = synth.draw_z(B, AcousticConfig().latent, T4 // 4, seed)
    out = _run_acoustic(eng, batch, False, z)
    with torch.no_grad():
        ref = O.acoustic_forward(W, AcousticConfig(), batch, batch["mel2word"], z)
    assert out["mel_out"].shape == (B, T4, 80)
    assert torch.equal(out["mel2word"].cpu(), ref["mel2word"])
    for k, tol in (("word_encoder_out", 2e-4), ("dict_attn", 1e-5), ("pron_attn", 1e-5), ("dur", 1e-4),
                   ("mel_out", TOL_MEL_MAXABS)):
        err = (out[k].cpu() - ref[k]).abs().max().item()
        assert err < tol, (k, err)
    x, nonpad = O.expand_by_mel2word(out["word_encoder_out"].cpu(), out["mel2word"].cpu())
    assert torch.equal(out["decoder_inp"].cpu(), x) and torch.equal(out["x_mask"].cpu(), nonpad)
    # predicted durations: integer path, bit-exact against the oracle's own prediction from ITS encoder output
    t = eng.text_encode(batch["word_tokens"], batch["pron_modified"], batch["keys"], batch["values"], batch["key_map"],
                        batch["pinyin"], batch["pinyin_map"])
    m2w = eng.length_regulate(t["dur_int"], t["ilens"]).cpu()
    assert torch.equal(m2w, O.length_regulate(t["dur_int"].cpu(), t["ilens"].cpu()))
    with torch.no_grad():
        dur_ref, pad_ref = O.duration_predictor(W, AcousticConfig(), ref["word_encoder_out"] *
                                                (batch["word_tokens"] != 0).float().unsqueeze(-1))
    assert torch.equal(t["dur_int"].cpu(), O.durations_to_int(dur_ref))
    assert torch.equal(t["ilens"].cpu(), (1 - pad_ref.long()).sum(-1))
    # no tone-sandhi overrides: None == zeros
    zero = torch.zeros_like(batch["pron_modified"])
    a = eng.text_encode(batch["word_tokens"], None, batch["keys"], batch["values"], batch["key_map"],
                        batch["pinyin"], batch["pinyin_map"])
    b = eng.text_encode(batch["word_tokens"], zero, batch["keys"], batch["values"], batch["key_map"],
                        batch["pinyin"], batch["pinyin_map"])
    for k in ("word_encoder_out", "pron_attn", "dur_int"):
        assert torch.equal(a[k], b[k]), k


# --------------------------------------------------------------------------------------------------
# fused prior flow (flow_fused_kernel): one launch for every coupling layer, against the per-layer path and the oracle
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,T", [(2, 4), (3, 100), (2, 448), (2, 452), (1, 1280), (60, 400)])
def test_fused_prior_flow_matches_per_layer_path_and_oracle(B, T):
    """T/4 latent frames: 1 (single row), 25, 112 (largest one-tile length), 113 (first two-tile length), 320 (four
    tiles, the cfg-1 utterance) and the cfg-2 batch; tiles overlap by 2 x 16 rows and store only their core rows."""
    from dict_tts_b200.engine import DictTTSEngine
    lib = binding.load()
    sd = synth.make_acoustic_state_dict(ACOUSTIC_SEED)
    eng = DictTTSEngine(sd, precision=1)
    W = fold_weight_norm(sd)
    cfg = AcousticConfig()
    g = torch.Generator().manual_seed(1000 + T)
    x = torch.randn(B, T, cfg.hidden, generator=g) * 0.5
    z = torch.randn(B, cfg.latent, T // 4, generator=g)
    g_bct = x.transpose(1, 2).contiguous().cuda()
    try:
        assert lib.dtts_debug_set_acoustic_fuse(1) == 0
        n0 = eng.launches
        mel_f, zp_f = eng.decode_mel(g_bct, z.cuda())
        n_fused = eng.launches - n0
        assert lib.dtts_debug_set_acoustic_fuse(0) == 0
        n0 = eng.launches
        mel_u, zp_u = eng.decode_mel(g_bct, z.cuda())
        n_unfused = eng.launches - n0
    finally:
        lib.dtts_debug_set_acoustic_fuse(-1)
    torch.cuda.synchronize()
    assert n_fused <= n_unfused - 70, (n_fused, n_unfused)
    with torch.no_grad():
        mel_r, zp_r = O.decode_mel(W, cfg, x, z)
    assert torch.isfinite(zp_f).all()
    assert (zp_f.cpu() - zp_r).abs().max().item() < 1e-4
    assert (zp_f - zp_u).abs().max().item() < 2e-5
    assert (mel_f.cpu() - mel_r).abs().max().item() < TOL_MEL_MAXABS
    eng.close()


@pytest.mark.parametrize("name", sorted(ACOUSTIC_CASES))
def test_fused_text_encoder_epilogues_match_per_layer_path(name):
    """LayerNorm fused into the O / FFN2 convolution epilogues (one standalone LayerNorm per encoder instead of nine)
    against the per-layer launches: same durations bit for bit, activations within fp32 reordering noise."""
    from dict_tts_b200.engine import DictTTSEngine
    lib = binding.load()
    eng = DictTTSEngine(synth.make_acoustic_state_dict(ACOUSTIC_SEED), precision=1)
    kw, _ = ACOUSTIC_CASES[name]
    b = synth.make_batch(**kw)
    args = (b["word_tokens"], b["pron_modified"], b["keys"], b["values"], b["key_map"], b["pinyin"], b["pinyin_map"])
    try:
        assert lib.dtts_debug_set_acoustic_fuse(1) == 0
        n0 = eng.launches
        f = {k: v.clone() for k, v in eng.text_encode(*args).items() if torch.is_tensor(v)}
        n_fused = eng.launches - n0
        assert lib.dtts_debug_set_acoustic_fuse(0) == 0
        n0 = eng.launches
        u = {k: v.clone() for k, v in eng.text_encode(*args).items() if torch.is_tensor(v)}
        n_unfused = eng.launches - n0
    finally:
        lib.dtts_debug_set_acoustic_fuse(-1)
    torch.cuda.synchronize()
    assert n_fused <= n_unfused - 16, (n_fused, n_unfused)
    assert torch.equal(f["dur_int"], u["dur_int"])
    # half the fixture tolerances (the fused build also multiplies the S2PA projection pairs W_k^T W_q and W_o W_v once at
    # create time, so the two builds differ by more than a summation order)
    errs = {k: (f[k] - u[k]).abs().max().item() for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur")}
    tol = dict(word_encoder_out=1e-4, dict_attn=5e-6, pron_attn=5e-6, dur=5e-5)
    assert all(torch.isfinite(f[k]).all() for k in errs)
    assert all(errs[k] < tol[k] for k in errs), errs
    eng.close()
