"""CPU suite for the NEXT scope row (the PortaSpeech non-dict sibling): oracle/ps_oracle.py against
tests/golden/ps_small.npz, which holds weights, inputs and the outputs of the UNMODIFIED reference model
(modules/portaspeech/model.py, reduced size, `use_post_glow=False` -- `modules/glow` is not in the reference checkout),
written by oracle/make_golden_ps.py.  No product code runs this path yet; the test pins the oracle it will be built against."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import ps_oracle as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGES = ("ph_encoder_out", "word_encoder_out", "dur", "attn", "decoder_inp", "z_p", "mel_out")


@pytest.fixture(scope="module")
def gold():
    d = np.load(os.path.join(ROOT, "tests", "golden", "ps_small.npz"))
    W = {k[2:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("W/")}
    cfg = types.SimpleNamespace(**{k: int(v) for k, v in (x.split("=") for x in d["dims"])})
    return d, W, cfg


@pytest.mark.parametrize("tag", ["given", "pred"])
def test_ps_oracle_matches_reference_outputs(gold, tag):
    d, W, cfg = gold
    txt, ph2word = torch.from_numpy(d["txt_tokens"]), torch.from_numpy(d["ph2word"])
    word_len = torch.from_numpy(d["word_len"]).max()
    m2w = torch.from_numpy(d["mel2word"]) if tag == "given" else None
    with torch.no_grad():
        out = P.ps_forward(W, cfg, txt, ph2word, word_len, m2w, torch.from_numpy(d[f"{tag}_z_in"]))
    assert np.array_equal(out["mel2word"].numpy(), d[f"{tag}_mel2word"])          # integer path: bit-exact
    for k in STAGES:
        err = np.abs(out[k].numpy() - d[f"{tag}_{k}"]).max()
        assert err < 2e-5, (k, err)
    # frames only attend to the phonemes of their own word
    w = out["attn"]
    valid = (out["mel2word"] > 0)[:, :, None]                                    # padded frames match the padded phonemes
    same_word = out["mel2word"][:, :, None] == ph2word[:, None, :]
    assert float((w * (~same_word) * valid).abs().max()) < 1e-6
    assert torch.allclose((w * same_word).sum(-1)[valid[..., 0]], torch.ones(int(valid.sum())), atol=1e-5)


def test_relative_position_terms_matter(gold):
    d, W, cfg = gold
    txt = torch.from_numpy(d["txt_tokens"])
    with torch.no_grad():
        a = P.ph_encode(W, cfg, txt)
        b = P.ph_encode({k: (torch.zeros_like(v) if "emb_rel" in k else v) for k, v in W.items()}, cfg, txt)
    assert float((a - b).abs().max()) > 0.05


def test_group_hidden_by_segs_is_a_segment_mean():
    h = torch.arange(24, dtype=torch.float32).view(1, 6, 4)
    seg = torch.tensor([[1, 1, 2, 3, 3, 0]])
    g = P.group_hidden_by_segs(h, seg, 4)
    assert g.shape == (1, 4, 4)
    assert torch.equal(g[0, 0], h[0, :2].mean(0)) and torch.equal(g[0, 1], h[0, 2]) and torch.equal(g[0, 2], h[0, 3:5].mean(0))
    assert (g[0, 3] == 0).all()                                                  # a word without phonemes


@pytest.mark.parametrize("tag", ["given", "pred"])
def test_ps_oracle_matches_full_size_reference_outputs(tag):
    """tests/golden/ps_full.npz: outputs of the unmodified reference model at the shipped Biaobei size for the seeded
    synthetic checkpoint (weights and inputs are regenerated from their seeds)."""
    from dict_tts_b200 import synth
    from dict_tts_b200.config import PortaSpeechConfig
    from dict_tts_b200.weights import fold_weight_norm
    from tests.cases import PS_FULL_BATCH, PS_FULL_PH_SIZE, PS_FULL_WEIGHT_SEED, PS_STAGES
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ps_full.npz"))
    cfg = PortaSpeechConfig(ph_size=PS_FULL_PH_SIZE)
    W = fold_weight_norm(synth.make_ps_state_dict(PS_FULL_WEIGHT_SEED, cfg))
    b = synth.make_ps_batch(ph_size=PS_FULL_PH_SIZE, **PS_FULL_BATCH)
    with torch.no_grad():
        out = P.ps_forward(W, cfg, b["txt_tokens"], b["ph2word"], b["word_lengths"].max(),
                           b["mel2word"] if tag == "given" else None, torch.from_numpy(gold[f"{tag}_z_in"]))
    assert np.array_equal(out["mel2word"].numpy(), gold[f"{tag}_mel2word"])
    for k in PS_STAGES:
        assert np.abs(out[k].numpy() - gold[f"{tag}_{k}"]).max() < 2e-5, k
