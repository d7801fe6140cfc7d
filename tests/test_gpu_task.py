"""GPU suite (-m gpu) for the drop-in seams: the ``--infer`` entry point on a fabricated experiment tree in the
reference's on-disk formats, and the ``BaseVocoder.spec2wav`` plugin class."""
import csv
import os

import numpy as np
import pytest
import torch

from dict_tts_b200 import hparams as hp_mod, synth
from tests import fake_exp
from dict_tts_b200.config import AcousticConfig, VocoderConfig
from dict_tts_b200.data import DictTTSTestSet
from dict_tts_b200.weights import fold_weight_norm
from oracle import dtts_oracle as O
from tests.cases import TOL_MEL_MAXABS, TOL_WAV_RMS, VOCODER_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exp(tmp_path_factory):
    return fake_exp.write(str(tmp_path_factory.mktemp("exp")), n_items=5)


def _read_wav(path):
    from scipy.io import wavfile
    sr, pcm = wavfile.read(path)
    return sr, pcm


@pytest.mark.parametrize("max_sentences", [1, 3])
def test_infer_entry_point_writes_reference_outputs(exp, max_sentences):
    from dict_tts_b200 import run
    cwd = os.getcwd()
    os.chdir(exp["root"])
    try:
        torch.manual_seed(1234)
        results = run.main(["--exp_name", exp["exp"], "--infer", "--hparams",
                            f"b200_max_sentences={max_sentences},gen_dir_name=bs{max_sentences}"])
    finally:
        os.chdir(cwd)
    gen = os.path.join(exp["work_dir"], f"generated_3000_bs{max_sentences}")
    with open(os.path.join(gen, "meta.csv")) as f:
        rows = list(csv.DictReader(f))
    assert len(rows) == len(results) == exp["n_items"]
    assert sorted(r["item_name"] for r in rows) == sorted(f"fake_{i:03d}" for i in range(exp["n_items"]))
    for r in rows:
        sr, pcm = _read_wav(os.path.join(gen, "wavs", r["wav_fn_pred"] + ".wav"))
        assert sr == 22050 and pcm.dtype == np.int16 and len(pcm) > 0 and len(pcm) % 256 == 0
        n_chars = len(r["text"])
        assert len(r["pinyin_tokens"].split()) == 2 * n_chars          # two pinyin tokens per character


def test_task_step_matches_oracle(exp):
    """One batched test_step of the standalone task == oracle forward on the same collated batch (predicted durations)."""
    from dict_tts_b200.task import B200DictTTSTask
    hp = hp_mod.set_hparams("", exp["exp"], "", root=exp["root"], global_hparams=False)
    hp["work_dir"] = exp["work_dir"]
    task = B200DictTTSTask(hp)
    task.build_model()
    batch = next(DictTTSTestSet(hp).batches(3))
    B, Tw = batch["word_tokens"].shape
    W = fold_weight_norm(synth.make_acoustic_state_dict(1234))
    torch.manual_seed(7)
    out = task.run_model(batch)
    T = out["mel_out"].shape[1]
    torch.manual_seed(7)
    z = torch.distributions.Normal(0, 1).sample([B, 16, T // 4])
    with torch.no_grad():
        ref = O.acoustic_forward(W, AcousticConfig(), batch, None, z)
    assert torch.equal(out["mel2word"].cpu(), ref["mel2word"])
    assert (out["mel_out"].cpu() - ref["mel_out"]).abs().max() < TOL_MEL_MAXABS
    assert (out["pron_attn"].cpu() - ref["pron_attn"]).abs().max() < 1e-5


@pytest.mark.parametrize("name", sorted(VOCODER_CASES))
def test_vocoder_plugin_spec2wav(name, golden_dir):
    from dict_tts_b200.plugin import B200HifiGAN
    voc = B200HifiGAN(state_dict=synth.make_vocoder_state_dict(4321), config=None)
    kw = VOCODER_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))["wav"]
    mel = synth.make_mel(kw["seed"], kw["B"], kw["T"])
    wav = voc.spec2wav(mel[0].numpy())                    # ndarray [T,80] in, ndarray [T*256] out (hifigan.py:54-62)
    assert isinstance(wav, np.ndarray) and wav.dtype == np.float32 and wav.shape == (kw["T"] * 256,)
    assert float(np.sqrt(np.mean((wav - gold[0]) ** 2))) < TOL_WAV_RMS
    batch = voc.spec2wav_batch(mel.cuda())
    assert batch.is_cuda and batch.shape == (kw["B"], kw["T"] * 256)
    assert float(np.sqrt(np.mean((batch.cpu().numpy() - gold) ** 2))) < TOL_WAV_RMS
    with pytest.raises(ValueError):
        voc.spec2wav(mel.numpy())                         # a batch is not one utterance


def test_vocoder_plugin_from_checkpoint_dir(exp):
    from dict_tts_b200.plugin import B200HifiGAN
    hp_mod.hparams.clear()
    hp_mod.hparams.update(exp["hparams"])
    voc = B200HifiGAN()
    mel = synth.make_mel(3, 1, 20)
    wav = voc.spec2wav(mel[0])
    with torch.no_grad():
        ref = O.hifigan_forward(fold_weight_norm(synth.make_vocoder_state_dict(4321)), VocoderConfig(), mel)
    assert float((torch.from_numpy(wav) - ref[0]).pow(2).mean().sqrt()) < TOL_WAV_RMS


def test_pipeline_stream_equals_serial_calls():
    """TextToWav.synthesize_stream (upload of batch i+1 overlapped with the compute of batch i) returns, batch by batch,
    exactly what the one-shot synthesize() returns, and both agree with the oracle."""
    from dict_tts_b200.pipeline import TextToWav
    pipe = TextToWav(synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321))
    batches = [synth.make_batch(seed=40 + i, B=3, min_chars=3, max_chars=6, max_frames=40, Lk_cap=32) for i in range(4)]
    serial = [pipe.synthesize(b).clone() for b in batches]
    streamed = [w.clone() for w in pipe.synthesize_stream(iter(batches))]
    assert len(streamed) == len(serial)
    for a, b in zip(serial, streamed):
        assert torch.equal(a, b)
    # the opt-in variant that runs the acoustic model of batch i+1 on its own stream next to the vocoder of batch i
    overlapped = [w.clone() for w in pipe.synthesize_stream(iter(batches), overlap_acoustic=True)]
    assert len(overlapped) == len(serial)
    for a, b in zip(serial, overlapped):
        assert torch.equal(a, b)
    W = fold_weight_norm(synth.make_acoustic_state_dict(1234))
    Wv = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    with torch.no_grad():
        ref = O.acoustic_forward(W, AcousticConfig(), batches[2], batches[2]["mel2word"], batches[2]["z_p"])
        ref_wav = O.hifigan_forward(Wv, VocoderConfig(), ref["mel_out"])
    # the pipeline vocodes up to each utterance's valid length (zeros after it): compare the valid samples
    valid = (batches[2]["mel2word"] > 0).sum(-1) * 256
    se, n = 0.0, 0
    for b, v in enumerate(valid.tolist()):
        se += float((streamed[2][b, :v] - ref_wav[b, :v]).pow(2).sum())
        n += v
        assert (streamed[2][b, v:] == 0).all()
    assert (se / n) ** 0.5 < TOL_WAV_RMS
    # ... and with trim_padding off the padded frames are vocoded too, exactly like the batched reference forward
    full = TextToWav(synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321), trim_padding=False)
    whole = full.synthesize(batches[2])
    assert float((whole - ref_wav).pow(2).mean().sqrt()) < TOL_WAV_RMS
    for b, v in enumerate(valid.tolist()):
        assert torch.equal(whole[b, :v], streamed[2][b, :v])
    full.close()
    pipe.close()


def test_dictionary_bank_path_equals_explicit_dict_msg():
    """naming characters by bank id must give exactly what shipping keys/values gives."""
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.engine import DictTTSEngine
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    for seed, kw in ((61, dict(B=4, min_chars=1, max_chars=9, max_frames=64, Lk_cap=64, pron_modified_p=0.1)),
                     (62, dict(B=2, min_chars=6, max_chars=6, max_frames=40, Lk_cap=96))):
        batch = synth.make_batch(seed=seed, **kw)
        bank, ids = DictBank.from_batch(batch)
        eng.set_dict_bank(bank)
        Lk, Lp = batch["key_map"].shape[2], batch["pinyin"].shape[2]
        ref = eng.text_encode(batch["word_tokens"], batch["pron_modified"], batch["keys"], batch["values"],
                              batch["key_map"], batch["pinyin"], batch["pinyin_map"])
        got = eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], ids, Lk, Lp)
        for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "dur_int", "ilens"):
            assert torch.equal(ref[k], got[k]), k
        # forward() with dict_ids only (widths derived from the bank) agrees with the oracle on the collated tensors
        out = eng.forward((batch["word_tokens"],), batch["pron_modified"], dict_ids=ids, mel2word=batch["mel2word"],
                          z_p=batch["z_p"])
        lk, lp = bank.batch_dims(ids)
        exp = dict(batch, **bank.collate(ids, lk, lp))
        W = fold_weight_norm(synth.make_acoustic_state_dict(1234))
        with torch.no_grad():
            want = O.acoustic_forward(W, AcousticConfig(), exp, batch["mel2word"], batch["z_p"])
        assert (out["mel_out"].cpu() - want["mel_out"]).abs().max() < TOL_MEL_MAXABS
        assert (out["dict_attn"].cpu() - want["dict_attn"]).abs().max() < 1e-5
    eng.close()


def test_bank_gather_reports_bad_ids_and_widths():
    """dtts_text_encode_bank does not synchronise: an id outside the bank / an entry wider than the call's Lk, Lp is
    recorded on the device and reported once -- by dtts_acoustic_status or at the sync of dtts_length_regulate_scan --
    instead of silently becoming a zero / truncated row."""
    import ctypes as C
    from dict_tts_b200 import binding
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.engine import DictTTSEngine
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    lib = binding.load()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    batch = synth.make_batch(seed=61, B=2, min_chars=4, max_chars=6, max_frames=40, Lk_cap=48)
    bank, ids = DictBank.from_batch(batch)
    eng.set_dict_bank(bank)
    Lk, Lp = bank.batch_dims(ids)
    good = eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], ids, Lk, Lp)
    assert lib.dtts_acoustic_status(eng.handle, stream, 1) == 0
    bad = ids.clone()
    bad[0, 1] = bank.n_entries + 5                       # explicit widths: the host-side validation is skipped
    eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], bad, Lk, Lp)
    assert lib.dtts_acoustic_status(eng.handle, stream, 1) == binding.DTTS_ERR_BAD_ARG
    assert b"n_entries" in lib.dtts_last_error()
    assert lib.dtts_acoustic_status(eng.handle, stream, 1) == 0          # reported once
    if Lk > 1:
        out = eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], ids, Lk - 1, Lp)
        with pytest.raises(RuntimeError, match="longer than"):          # surfaces at the length regulator's sync
            eng.length_regulate(out["dur_int"], out["ilens"])
    again = eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], ids, Lk, Lp)
    assert lib.dtts_acoustic_status(eng.handle, stream, 1) == 0
    assert torch.equal(again["word_encoder_out"], good["word_encoder_out"])
    with pytest.raises(ValueError):                                      # without widths the ids are validated on the host
        eng.text_encode_bank(batch["word_tokens"], batch["pron_modified"], bad)
    eng.close()


def test_pipeline_with_bank_equals_explicit():
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.pipeline import TextToWav
    pipe = TextToWav(synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321))
    batch = synth.make_batch(seed=70, B=3, min_chars=3, max_chars=6, max_frames=40, Lk_cap=32)
    explicit = pipe.synthesize(batch).clone()
    bank, ids = DictBank.from_batch(batch)
    pipe.acoustic.set_dict_bank(bank)
    slim = {k: v for k, v in batch.items() if k not in ("keys", "values", "key_map", "pinyin", "pinyin_map")}
    slim["dict_ids"] = ids
    # same widths as the explicit batch so that the outputs are comparable bit for bit
    via_bank = pipe.synthesize(slim).clone()
    assert torch.allclose(explicit, via_bank, atol=2e-6)
    streamed = [w.clone() for w in pipe.synthesize_stream(iter([slim, slim]))]
    assert torch.equal(streamed[0], via_bank) and torch.equal(streamed[1], via_bank)
    pipe.close()


def test_device_after_infer_matches_host_semantics():
    """int16 conversion and pinyin-token selection on the device == the host code of the reference
    (utils/audio.py:15-16, tasks/tts/dict_tts.py:295-304), with the explicit pinyin tensor and with the bank."""
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.engine import DictTTSEngine
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    g = torch.Generator().manual_seed(5)
    wav = (torch.rand(3, 1024, generator=g) * 2 - 1) * 0.999
    wav[0, :4] = torch.tensor([0.99999, -0.99999, 0.0, 3.05e-5])
    pcm = eng.pcm16(wav.cuda()).cpu().numpy()
    assert pcm.dtype == np.int16 and np.array_equal(pcm, (wav.numpy() * 32767).astype(np.int16))
    batch = synth.make_batch(seed=81, B=3, min_chars=2, max_chars=7, max_frames=32, Lk_cap=48)
    B, Tw, Lp = batch["pinyin"].shape
    pron_attn = torch.rand(B, Tw, Lp, generator=g)
    pron_attn[0, 1, :] = 0.25                                    # ties -> first maximum
    pron_attn[1, 2, Lp - 1] = 2.0                                # maximum in the last slot -> one token only
    want = torch.full((B, Tw, 2), -1, dtype=torch.long)
    idx = pron_attn.max(-1)[1]
    for b in range(B):
        for t in range(Tw):
            sl = batch["pinyin"][b, t][idx[b, t]:idx[b, t] + 2]
            want[b, t, :len(sl)] = sl
    got = eng.pron_tokens(pron_attn, pinyin=batch["pinyin"]).cpu()
    assert torch.equal(got, want)
    bank, ids = DictBank.from_batch(batch)
    eng.set_dict_bank(bank)
    got_bank = eng.pron_tokens(pron_attn, dict_ids=ids).cpu()
    assert torch.equal(got_bank, want)
    eng.close()


def test_ragged_batch_equals_padded_batch(exp):
    """a batch collated ragged (batch-local bank of its distinct characters + dict_ids naming the rows
    the reference collater builds) gives bit for bit what the reference-shaped padded batch gives -- including the
    collater's quirk that the appended (key_map 1) row sits in column Tw-1 of every utterance."""
    from dict_tts_b200.engine import DictTTSEngine
    from dict_tts_b200.pipeline import TextToWav
    padded = next(DictTTSTestSet(exp["hparams"]).batches(max_sentences=4))
    ds = DictTTSTestSet(exp["hparams"])
    ds.ragged = True
    ragged = next(ds.batches(max_sentences=4))
    assert len(set(padded["word_lengths"].tolist())) > 1            # unequal lengths: the quirk is exercised
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    ref = eng.text_encode(padded["word_tokens"], padded["pron_modified"], padded["keys"], padded["values"],
                          padded["key_map"], padded["pinyin"], padded["pinyin_map"])
    eng.set_dict_bank(ragged["dict_bank"])
    got = eng.text_encode_bank(ragged["word_tokens"], ragged["pron_modified"], ragged["dict_ids"])
    for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "dur_int", "ilens"):
        assert torch.equal(ref[k], got[k]), k
    eng.close()
    # through the public pipeline: the ragged batch carries its own bank (pinned), serial and streamed calls agree
    pipe = TextToWav(synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321))
    torch.manual_seed(7)
    T4 = (padded["mel2word"].shape[1] + 3) // 4 * 4
    z = torch.randn(padded["word_tokens"].shape[0], 16, T4 // 4)
    want = pipe.synthesize(dict(padded, z_p=z)).clone()
    rb = dict(ragged, z_p=z, dict_bank=ragged["dict_bank"].pin_memory())
    got_wav = pipe.synthesize(rb).clone()
    assert torch.equal(want, got_wav)
    streamed = [w.clone() for w in pipe.synthesize_stream(iter([rb, rb]))]
    assert torch.equal(streamed[0], want) and torch.equal(streamed[1], want)
    h2d_ragged = sum(t.numel() * t.element_size() for t in rb["dict_bank"].tensors()[:1]) + rb["dict_ids"].numel() * 8
    assert h2d_ragged < padded["keys"].numel() * 4
    pipe.close()


def test_infer_entry_point_dict_modes_agree(exp):
    """The three ways the dictionary features reach the engine -- GPU-resident bank (default), ragged per-batch bank,
    the reference collater's padded tensors -- write identical waveforms and pinyin tokens; so does the reference's
    own rank dealing (deal=reference) up to the order of the rows."""
    from dict_tts_b200 import run
    outs = {}
    cwd = os.getcwd()
    os.chdir(exp["root"])
    try:
        for mode, extra in (("bank", ""), ("ragged", ""), ("padded", ""), ("bank", ",b200_deal=reference")):
            tag = mode + ("_ref" if extra else "")
            torch.manual_seed(1234)
            run.main(["--exp_name", exp["exp"], "--infer", "--hparams",
                      f"b200_max_sentences=3,gen_dir_name=m_{tag},b200_dict_mode={mode}{extra}"])
            gen = os.path.join(exp["work_dir"], f"generated_3000_m_{tag}")
            with open(os.path.join(gen, "meta.csv")) as f:
                rows = {r["item_name"]: r for r in csv.DictReader(f)}
            outs[tag] = {n: (r["pinyin_tokens"], _read_wav(os.path.join(gen, "wavs", r["wav_fn_pred"] + ".wav"))[1])
                         for n, r in rows.items()}
    finally:
        os.chdir(cwd)
    assert len(outs["bank"]) == exp["n_items"]
    for tag in ("ragged", "padded"):
        assert outs[tag].keys() == outs["bank"].keys()
        for n, (tok, pcm) in outs["bank"].items():
            assert outs[tag][n][0] == tok, (tag, n)
            assert np.array_equal(outs[tag][n][1], pcm), (tag, n)
    # reference dealing groups the utterances differently (natural order, not longest-first): the noise z and the
    # padding differ, so only the discrete outputs are comparable
    assert outs["bank_ref"].keys() == outs["bank"].keys()
    for n, (tok, _) in outs["bank"].items():
        assert outs["bank_ref"][n][0] == tok
