"""CPU suite: the oracle against the reference-generated golden fixtures, host logic, and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from dict_tts_b200 import binding, synth
from dict_tts_b200.config import AcousticConfig, VocoderConfig
from dict_tts_b200.weights import drop_dead, fold_weight_norm, pack_arena
from oracle import dtts_oracle as O
from tests.cases import ACOUSTIC_CASES, ACOUSTIC_SEED, VOCODER_CASES, VOCODER_SEED

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def acoustic_weights():
    return fold_weight_norm(synth.make_acoustic_state_dict(ACOUSTIC_SEED))


@pytest.fixture(scope="module")
def vocoder_weights():
    return fold_weight_norm(synth.make_vocoder_state_dict(VOCODER_SEED))


@pytest.mark.parametrize("name", sorted(ACOUSTIC_CASES))
def test_oracle_matches_reference_acoustic(name, golden_dir, acoustic_weights):
    kw, predicted = ACOUSTIC_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    batch = synth.make_batch(**kw)
    cfg = AcousticConfig()
    z = torch.from_numpy(gold["z_in"])
    with torch.no_grad():
        out = O.acoustic_forward(acoustic_weights, cfg, batch, None if predicted else batch["mel2word"], z)
    assert np.array_equal(out["mel2word"].numpy(), gold["mel2word"])          # integer path: bit-exact
    for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "decoder_inp", "z_p", "mel_out"):
        err = np.abs(out[k].numpy() - gold[k]).max()
        assert err < 2e-5, (k, err)


@pytest.mark.parametrize("name", sorted(VOCODER_CASES))
def test_oracle_matches_reference_vocoder(name, golden_dir, vocoder_weights):
    kw = VOCODER_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    mel = synth.make_mel(kw["seed"], kw["B"], kw["T"])
    with torch.no_grad():
        wav = O.hifigan_forward(vocoder_weights, VocoderConfig(), mel)
    assert wav.shape == (kw["B"], kw["T"] * 256)
    assert np.abs(wav.numpy() - gold["wav"]).max() < 2e-5


def test_length_regulator_edge_cases():
    # all-zero durations -> filled with ones (tts_modules.py:248-250); zero-length words are skipped
    dur = torch.tensor([[0, 0, 0, 0], [2, 0, 3, 9], [1, 1, 0, 0]])
    ilens = torch.tensor([3, 3, 2])
    m = O.length_regulate(dur, ilens, 4)
    assert m.shape == (3, 8)
    assert m[0].tolist() == [1, 2, 3, 0, 0, 0, 0, 0]
    assert m[1].tolist() == [1, 1, 3, 3, 3, 3, 3, 3]      # T_raw = 5; padded columns repeat the last column (word 3)
    assert m[2].tolist() == [1, 2, 0, 0, 0, 0, 0, 0]


def test_durations_round_half_to_even():
    d = torch.log(torch.tensor([1.5, 2.5, 3.5, 0.2]) + 1)
    assert O.durations_to_int(d).tolist() == [2, 2, 4, 0]


def test_fold_and_pack_arena():
    sd = synth.make_vocoder_state_dict(VOCODER_SEED)
    W = fold_weight_norm(sd)
    assert not any(k.endswith(("weight_g", "weight_v")) for k in W)
    v, g = sd["conv_pre.weight_v"], sd["conv_pre.weight_g"]
    ref = g * v / v.flatten(1).norm(dim=1).view(-1, 1, 1)
    assert torch.allclose(W["conv_pre.weight"], ref, atol=1e-6)
    arena, table = pack_arena(W)
    for name, off, n in table:
        assert off % 16 == 0
        assert torch.equal(arena[off:off + n], W[name].reshape(-1))
    ac = drop_dead(fold_weight_norm(synth.make_acoustic_state_dict(ACOUSTIC_SEED)))
    assert "enc_pos_proj.weight" not in ac and "dict_encoder.S2PA_module.emb.weight" not in ac
    assert "dict_encoder.S2PA_module.word_emb.weight" in ac


def test_batch_follows_collater_contract():
    b = synth.make_batch(seed=5, B=4, min_chars=2, max_chars=6, max_frames=32, Lk_cap=48)
    wt, km, pm, py = b["word_tokens"], b["key_map"], b["pinyin_map"], b["pinyin"]
    n = b["word_lengths"]
    for i in range(4):
        L = int(n[i])
        assert wt[i, 0] == 1 and wt[i, L - 1] == 1 and (wt[i, L:] == 0).all()
        assert (km[i, 0] == 1).all() and (km[i, L - 1] == 1).all()           # BOS/EOS rows (dataset_utils.py:286-296)
        assert (pm[i, 0] == 1).all() and (py[i, 0] == 0).all()
        assert (b["keys"][i, 0] == 0).all()
        assert (b["mel2word"][i] > 0).sum() == b["mel_lengths"][i]
        assert int(b["mel2word"][i].max()) <= L
    assert b["z_p"].shape == (4, 16, 8)


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "dtts.h")).read()
    declared = set(re.findall(r"\b(dtts_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(binding.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dtts.h but not exported"
    assert declared == set(binding.SYMBOLS), declared ^ set(binding.SYMBOLS)
    assert binding.load().dtts_abi_version() == binding.ABI_VERSION


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dict_tts_b200.engine import DictTTSEngine, HifiGanEngine
    with pytest.raises(RuntimeError):
        DictTTSEngine({})
    with pytest.raises(RuntimeError):
        HifiGanEngine({})


def test_abi_error_convention_without_a_gpu():
    """Every entry point returns a negative dtts_status and leaves a message for dtts_last_error() instead of touching
    the device when its arguments are unusable (include/dtts.h: status codes; SURVEY.md §8b error convention).  None of
    these calls reaches a kernel launch, so they run on the CPU box too."""
    lib = binding.load()
    BAD_ARG, BAD_SHAPE = -1, -2

    def last():
        return lib.dtts_last_error().decode()

    null = ctypes.c_void_p()
    out = ctypes.c_void_p()
    assert lib.dtts_acoustic_create(None, None, 0, None, 0, None, ctypes.byref(out)) == BAD_ARG and "null" in last()
    assert lib.dtts_vocoder_create(None, None, 0, None, 0, None, ctypes.byref(out)) == BAD_ARG
    d = binding.AcousticDesc(192, 2, 4, 5, 768, 768, 8000, 185, 3, 5, 128, 4, 16, 4, 5, 64, 3, 4, 4, 80, 1, 7, 0)
    assert lib.dtts_acoustic_create(ctypes.byref(d), None, 0, None, 0, None, ctypes.byref(out)) == BAD_ARG
    assert "precision" in last()
    d.precision, d.s2pa_route = 0, 1
    assert lib.dtts_acoustic_create(ctypes.byref(d), None, 0, None, 0, None, ctypes.byref(out)) == BAD_ARG
    assert "s2pa_route" in last()
    d.s2pa_route, d.frames_multiple = 0, 3
    assert lib.dtts_acoustic_create(ctypes.byref(d), None, 0, None, 0, None, ctypes.byref(out)) == BAD_SHAPE
    v = binding.VocoderDesc()
    v.n_mel, v.init_ch, v.n_ups, v.n_rb, v.precision = 80, 512, 1, 1, 9
    v.up_rates[0], v.up_kernels[0] = 8, 16
    assert lib.dtts_vocoder_create(ctypes.byref(v), None, 0, None, 0, None, ctypes.byref(out)) == BAD_ARG
    assert "precision" in last()
    v.precision, v.up_kernels[0] = 3, 15                      # kernel not a multiple of the rate
    assert lib.dtts_vocoder_create(ctypes.byref(v), None, 0, None, 0, None, ctypes.byref(out)) == BAD_SHAPE
    assert lib.dtts_vocode(None, None, 1, 1, None, None, 0, None) == BAD_ARG
    assert lib.dtts_vocode_lens(None, None, None, 1, 1, None, None, 0, None) == BAD_ARG and "lengths" in last()
    assert lib.dtts_text_encode(None, None, None, None, 0, None) == BAD_ARG
    assert lib.dtts_vocode_workspace_bytes(None, 1, 1) == 0 and lib.dtts_text_workspace_bytes(None, 1, 1, 1, 1) == 0
    assert lib.dtts_acoustic_destroy(None) == 0 and lib.dtts_vocoder_destroy(None) == 0
    with pytest.raises(RuntimeError, match="status -1"):
        binding.check(BAD_ARG, "probe")
