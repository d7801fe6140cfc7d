"""GPU parity suite of the PortaSpeech (non-dict) sibling, SURVEY.md §8f-3 (-m gpu): the CUDA path through the C ABI
(dtts_ps_text_encode / dtts_ps_attend + the dict model's length regulator and dtts_decode_mel) against

* tests/golden/ps_small.npz -- weights, inputs and the outputs of the UNMODIFIED reference model at a reduced size
  (hidden 48: outside what the tcgen05 kernel tiles, so it pins the fp32 FMA build, precision 0);
* tests/golden/ps_full.npz -- reference outputs at the shipped Biaobei size for the seeded synthetic checkpoint (both
  builds: precision 0 and the tcgen05 3-MMA split, precision 1);
* oracle/ps_oracle.py on shapes the fixtures do not hold (phoneme sequences longer than one attention tile, one word
  per utterance, unsorted lengths).
Tolerances: mel <= 1e-3 max-abs (north star), intermediate stages <= 2e-4, integer mel2word bit-exact."""
import os
import types

import numpy as np
import pytest
import torch

from dict_tts_b200 import synth
from dict_tts_b200.config import PortaSpeechConfig
from dict_tts_b200.weights import fold_weight_norm
from oracle import ps_oracle as P
from tests.cases import PS_FULL_BATCH, PS_FULL_PH_SIZE, PS_FULL_WEIGHT_SEED, PS_STAGES, TOL_MEL_MAXABS

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_STAGE = 2e-4


def _check(out, gold, tag):
    assert np.array_equal(out["mel2word"].cpu().numpy(), gold[f"{tag}_mel2word"])          # integer path: bit-exact
    for k in PS_STAGES:
        got = out[k].cpu().numpy()
        if k == "x_mask":
            continue
        err = np.abs(got - gold[f"{tag}_{k}"]).max()
        assert err < (TOL_MEL_MAXABS if k == "mel_out" else TOL_STAGE), (tag, k, err)


@pytest.mark.parametrize("tag", ["given", "pred"])
def test_ps_small_fixture_fp32_build(tag):
    from dict_tts_b200.engine import PortaSpeechEngine
    d = np.load(os.path.join(ROOT, "tests", "golden", "ps_small.npz"))
    W = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("W/")}
    dims = {k: int(v) for k, v in (x.split("=") for x in d["dims"])}
    cfg = PortaSpeechConfig(hidden=dims["hidden"], n_heads=dims["n_heads"], enc_layers=dims["enc_layers"],
                            ffn_kernel=dims["ffn_kernel"], ffn_filter=4 * dims["hidden"], dur_layers=dims["dur_layers"],
                            dur_kernel=dims["dur_kernel"], latent=dims["latent"], dec_layers=dims["dec_layers"],
                            dec_kernel=dims["dec_kernel"], flow_hidden=dims["flow_hidden"], flow_kernel=dims["flow_kernel"],
                            flow_blocks=dims["flow_blocks"], flow_layers=dims["flow_layers"], n_mel=dims["n_mel"],
                            ph_size=W["ph_encoder.emb.weight"].shape[0], word_enc_layers=dims["word_enc_layers"])
    eng = PortaSpeechEngine(W, cfg, precision=0)
    txt, ph2word = torch.from_numpy(d["txt_tokens"]), torch.from_numpy(d["ph2word"])
    m2w = torch.from_numpy(d["mel2word"]) if tag == "given" else None
    out = eng.forward(txt, ph2word, int(d["word_len"].max()), mel2word=m2w, z_p=torch.from_numpy(d[f"{tag}_z_in"]))
    _check(out, d, tag)
    eng.close()


@pytest.fixture(scope="module")
def full():
    cfg = PortaSpeechConfig(ph_size=PS_FULL_PH_SIZE)
    sd = synth.make_ps_state_dict(PS_FULL_WEIGHT_SEED, cfg)
    return cfg, sd, synth.make_ps_batch(ph_size=PS_FULL_PH_SIZE, **PS_FULL_BATCH), \
        np.load(os.path.join(ROOT, "tests", "golden", "ps_full.npz"))


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("tag", ["given", "pred"])
def test_ps_full_size_fixture(full, precision, tag):
    from dict_tts_b200.engine import PortaSpeechEngine
    cfg, sd, b, gold = full
    eng = PortaSpeechEngine(sd, cfg, precision=precision)
    out = eng.forward(b["txt_tokens"], b["ph2word"], int(b["word_lengths"].max()),
                      mel2word=b["mel2word"] if tag == "given" else None, z_p=torch.from_numpy(gold[f"{tag}_z_in"]))
    _check(out, gold, tag)
    # frames attend only to the phonemes of their own word
    w = out["attn"].cpu()
    m2w = out["mel2word"].cpu()
    same = m2w[:, :, None] == b["ph2word"][:, None, :]
    valid = (m2w > 0)[:, :, None]
    assert float((w * (~same) * valid).abs().max()) < 1e-6
    eng.close()


@pytest.mark.parametrize("precision", [0, 1])
@pytest.mark.parametrize("kw", [
    dict(seed=41, B=2, min_words=30, max_words=40, max_ph_per_word=4, max_frames=400),     # Tp > 64: several attention tiles
    dict(seed=42, B=3, min_words=1, max_words=1, max_ph_per_word=3, max_frames=8),         # one word per utterance
    dict(seed=43, B=5, min_words=2, max_words=12, max_ph_per_word=2, max_frames=96),
])
def test_ps_shapes_against_oracle(full, precision, kw):
    from dict_tts_b200.engine import PortaSpeechEngine
    cfg, sd, _, _ = full
    W = fold_weight_norm(sd)
    b = synth.make_ps_batch(ph_size=PS_FULL_PH_SIZE, **kw)
    eng = PortaSpeechEngine(sd, cfg, precision=precision)
    for m2w in (b["mel2word"], None):
        if m2w is None:
            with torch.no_grad():
                want = P.ps_forward(W, cfg, b["txt_tokens"], b["ph2word"], b["word_lengths"].max(), None,
                                    None)                 # durations first: the prior sample needs T
            T4 = want["mel2word"].shape[1] // cfg.frames_multiple
            z = synth.draw_z(b["txt_tokens"].shape[0], cfg.latent, T4, 5)
        else:
            z = b["z_p"]
        with torch.no_grad():
            want = P.ps_forward(W, cfg, b["txt_tokens"], b["ph2word"], b["word_lengths"].max(), m2w, z)
        out = eng.forward(b["txt_tokens"], b["ph2word"], int(b["word_lengths"].max()), mel2word=m2w, z_p=z)
        assert torch.equal(out["mel2word"].cpu(), want["mel2word"])
        for k in PS_STAGES:
            err = float((out[k].cpu() - want[k]).abs().max())
            assert err < (TOL_MEL_MAXABS if k == "mel_out" else TOL_STAGE), (kw["seed"], precision, k, err)
    eng.close()


def test_ps_handle_refuses_dict_entry_points_and_vice_versa(full):
    import ctypes as C
    from dict_tts_b200 import binding
    from dict_tts_b200.engine import DictTTSEngine, PortaSpeechEngine
    cfg, sd, b, _ = full
    ps = PortaSpeechEngine(sd, cfg)
    batch = synth.make_batch(seed=3, B=2, min_chars=3, max_chars=5, max_frames=32, Lk_cap=32)
    with pytest.raises(RuntimeError, match="PortaSpeech"):
        ps.text_encode(batch["word_tokens"], batch["pron_modified"], batch["keys"], batch["values"], batch["key_map"],
                       batch["pinyin"], batch["pinyin_map"])
    dict_eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    lib = binding.load()
    assert lib.dtts_ps_text_workspace_bytes(dict_eng.handle, 2, 8, 4) == 0
    tin, tout = binding.PsTextIn(None, None, 2, 8, 4), binding.PsTextOut(None, None, None, None, None)
    assert lib.dtts_ps_text_encode(dict_eng.handle, C.byref(tin), C.byref(tout), C.c_void_p(1), 0, None) == binding.DTTS_ERR_BAD_ARG
    ps.close()
    dict_eng.close()


def test_portaspeech_infer_entry_point(tmp_path):
    """``python -m dict_tts_b200.run --exp_name <ps exp> --infer``: the reference's own task name in config.yaml maps onto
    B200PortaSpeechTask; checkpoint discovery, word-level collate, predicted durations, vocoder, wavs and meta.csv."""
    import csv
    from scipy.io import wavfile
    from dict_tts_b200 import run
    from dict_tts_b200.data import PortaSpeechTestSet
    from dict_tts_b200 import hparams as hp_mod
    from tests import fake_exp
    exp = fake_exp.write_ps(str(tmp_path), n_items=5)
    cwd = os.getcwd()
    os.chdir(exp["root"])
    try:
        torch.manual_seed(1234)
        results = run.main(["--exp_name", exp["exp"], "--infer", "--hparams", "b200_max_sentences=3,gen_dir_name=ps"])
    finally:
        os.chdir(cwd)
    gen = os.path.join(exp["work_dir"], "generated_2500_ps")
    with open(os.path.join(gen, "meta.csv")) as f:
        rows = list(csv.DictReader(f))
    assert len(rows) == len(results) == 5
    assert [r["item_name"] for r in rows] == [f"ps_{i:03d}" for i in range(5)]           # dataset order
    for r in rows:
        sr, pcm = wavfile.read(os.path.join(gen, "wavs", r["wav_fn_pred"] + ".wav"))
        assert sr == 22050 and pcm.dtype == np.int16 and len(pcm) > 0 and len(pcm) % 256 == 0
        assert all(tok.startswith("ph") for tok in r["ph_tokens"].split())
    # one test_step of the task == the oracle forward on the same collated batch (predicted durations)
    from dict_tts_b200.task import B200PortaSpeechTask
    hp = hp_mod.set_hparams("", exp["exp"], "", root=exp["root"], global_hparams=False)
    task = B200PortaSpeechTask(hp)
    task.build_model()
    ds = PortaSpeechTestSet(hp)
    batch = next(ds.batches(3))
    torch.manual_seed(5)
    out = task.run_model(batch)
    W = fold_weight_norm(synth.make_ps_state_dict(2468, PortaSpeechConfig(ph_size=80)))
    with torch.no_grad():
        want = P.ps_forward(W, task.model.cfg, batch["txt_tokens"], batch["ph2word"], batch["word_lengths"].max(), None,
                            out["z_p"].cpu() * 0 + synth.draw_z(3, 16, out["mel_out"].shape[1] // 4, 1))
    assert torch.equal(out["mel2word"].cpu(), want["mel2word"])
    assert float((out["word_encoder_out"].cpu() - want["word_encoder_out"]).abs().max()) < TOL_STAGE
    task.model.close()
