"""TEST INFRASTRUCTURE ONLY -- run in the build container (needs /root/reference).

Builds the UNMODIFIED reference PortaSpeech (modules/portaspeech/model.py) at a reduced size, removes weight-norm the way
PortaSpeechFlowTask.test_start does (tasks/tts/ps_flow.py:257-268), runs its inference forward on seeded inputs, asserts
oracle/ps_oracle.py reproduces every stage, and writes tests/golden/ps_small.npz (weights + inputs + REFERENCE outputs;
the model is small so that the weights fit in the fixture).

    python -m oracle.make_golden_ps
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ps_oracle as P  # noqa: E402
from oracle import ref_loader  # noqa: E402

DIMS = dict(hidden=48, n_heads=2, enc_layers=2, word_enc_layers=2, ffn_kernel=5, dur_layers=3, dur_kernel=5, latent=16,
            dec_layers=2, dec_kernel=5, flow_hidden=16, flow_kernel=3, flow_blocks=4, flow_layers=2, frames_multiple=4,
            n_mel=80)
HP = ("use_post_glow=False,hidden_size=48,enc_layers=2,word_enc_layers=2,fvae_enc_dec_hidden=48,fvae_dec_n_layers=2,"
      "fvae_enc_n_layers=2,prior_glow_hidden=16,prior_glow_n_blocks=2")


def main():
    ref_root = ref_loader.REF_ROOT
    for name in ("chardet", "librosa"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cwd = os.getcwd()
    os.chdir(ref_root)
    sys.path.insert(0, ref_root)
    try:
        from utils.hparams import set_hparams
        set_hparams(config="egs/datasets/audio/biaobei/ps_flow.yaml", exp_name="", hparams_str=HP, print_hparams=False)
        from utils.text_encoder import TokenTextEncoder
        from modules.portaspeech.model import PortaSpeech
        torch.manual_seed(77)
        enc = TokenTextEncoder(None, vocab_list=[f"p{i}" for i in range(40)], replace_oov="<UNK>")
        model = PortaSpeech(enc).eval()
    finally:
        os.chdir(cwd)
    # the pre-net projection and a few biases are zero-initialised: perturb everything a little so nothing is vacuous
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for prm in model.parameters():
            prm.add_(0.05 * torch.randn(prm.shape, generator=g))

    def strip(m):
        try:
            torch.nn.utils.remove_weight_norm(m)
        except ValueError:
            pass
    model.apply(strip)
    W = {k: v.detach().clone() for k, v in model.state_dict().items() if not k.startswith("fvae.encoder.")}

    B, Tp = 3, 14
    txt = torch.randint(3, 40, (B, Tp), generator=g)
    ph2word = torch.tensor([[1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 5, 5, 6, 6],
                            [1, 2, 2, 3, 3, 3, 4, 4, 4, 4, 0, 0, 0, 0],
                            [1, 1, 1, 2, 3, 3, 0, 0, 0, 0, 0, 0, 0, 0]])
    txt = txt * (ph2word > 0)
    word_len = ph2word.max(-1)[0]
    durs = torch.randint(2, 7, (B, int(word_len.max())), generator=g) * (torch.arange(int(word_len.max()))[None] < word_len[:, None])
    T = int(durs.sum(-1).max())
    mel2word = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        m = torch.repeat_interleave(torch.arange(1, durs.shape[1] + 1), durs[b])
        mel2word[b, :len(m)] = m
    cfg = types.SimpleNamespace(**DIMS)
    out = {}
    for tag, m2w in (("given", mel2word), ("pred", None)):
        torch.manual_seed(9)
        with torch.no_grad():
            ref = model(txt, ph2word, word_len.max(), mel2word=m2w, infer=True, forward_post_glow=False, two_stage=True)
        T4 = ref["mel_out"].shape[1]
        torch.manual_seed(9)
        z = torch.distributions.Normal(0, 1).sample([B, DIMS["latent"], T4 // 4])
        with torch.no_grad():
            mine = P.ps_forward(W, cfg, txt, ph2word, word_len.max(), m2w, z)
        for k in ("ph_encoder_out", "word_encoder_out", "dur", "attn", "decoder_inp", "z_p", "mel_out"):
            err = (mine[k] - ref[k]).abs().max().item()
            print(f"{tag:6s} {k:18s} max-abs diff oracle vs reference {err:.2e}")
            assert err < 2e-5, (tag, k, err)
        if m2w is None:
            assert mine["mel2word"].shape[1] == T4
        out.update({f"{tag}_{k}": ref[k].numpy() for k in ("ph_encoder_out", "word_encoder_out", "dur", "attn",
                                                            "decoder_inp", "z_p", "mel_out")})
        out[f"{tag}_z_in"] = z.numpy()
        out[f"{tag}_mel2word"] = mine["mel2word"].numpy()
    out.update(txt_tokens=txt.numpy(), ph2word=ph2word.numpy(), word_len=word_len.numpy(), mel2word=mel2word.numpy())
    out.update({"W/" + k: v.numpy() for k, v in W.items()})
    out["dims"] = np.array([f"{k}={v}" for k, v in DIMS.items()])
    path = os.path.join(ROOT, "tests", "golden", "ps_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, round(os.path.getsize(path) / 1e6, 2), "MB")
    full_size(PortaSpeech, TokenTextEncoder, set_hparams, ref_root)


FULL_BATCH = dict(seed=31, B=4, min_words=3, max_words=9, max_ph_per_word=4, max_frames=64)
FULL_WEIGHT_SEED = 2468
FULL_PH_SIZE = 80


def full_size(PortaSpeech, TokenTextEncoder, set_hparams, ref_root):
    """tests/golden/ps_full.npz: the reference model at the shipped Biaobei size (hidden 192, 4 + 4 layers) holding the
    seeded synthetic checkpoint of dict_tts_b200.synth.make_ps_state_dict -- outputs only, weights and inputs are
    regenerated from their seeds at test time."""
    from dict_tts_b200 import synth
    from dict_tts_b200.config import PortaSpeechConfig
    from dict_tts_b200.weights import fold_weight_norm
    cwd = os.getcwd()
    os.chdir(ref_root)
    try:
        set_hparams(config="egs/datasets/audio/biaobei/ps_flow.yaml", exp_name="", hparams_str="use_post_glow=False",
                    print_hparams=False)
        enc = TokenTextEncoder(None, vocab_list=[f"p{i}" for i in range(FULL_PH_SIZE - 3)], replace_oov="<UNK>")
        assert len(enc) == FULL_PH_SIZE
        model = PortaSpeech(enc).eval()
    finally:
        os.chdir(cwd)
    cfg = PortaSpeechConfig(ph_size=FULL_PH_SIZE)
    sd = synth.make_ps_state_dict(FULL_WEIGHT_SEED, cfg)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.startswith("fvae.encoder.") for m in missing), (missing, unexpected)

    def strip(m):
        try:
            torch.nn.utils.remove_weight_norm(m)
        except ValueError:
            pass
    model.apply(strip)
    W = fold_weight_norm(sd)
    b = synth.make_ps_batch(ph_size=FULL_PH_SIZE, **FULL_BATCH)
    out = {}
    for tag, m2w in (("given", b["mel2word"]), ("pred", None)):
        torch.manual_seed(9)
        with torch.no_grad():
            ref = model(b["txt_tokens"], b["ph2word"], b["word_lengths"].max(), mel2word=m2w, infer=True,
                        forward_post_glow=False, two_stage=True)
        T4 = ref["mel_out"].shape[1]
        torch.manual_seed(9)
        z = torch.distributions.Normal(0, 1).sample([b["txt_tokens"].shape[0], cfg.latent, T4 // 4])
        with torch.no_grad():
            mine = P.ps_forward(W, cfg, b["txt_tokens"], b["ph2word"], b["word_lengths"].max(), m2w, z)
        for k in ("ph_encoder_out", "word_encoder_out", "dur", "attn", "decoder_inp", "z_p", "mel_out"):
            err = (mine[k] - ref[k]).abs().max().item()
            print(f"full {tag:6s} {k:18s} max-abs diff oracle vs reference {err:.2e}")
            assert err < 2e-5, (tag, k, err)
        out.update({f"{tag}_{k}": ref[k].numpy() for k in ("ph_encoder_out", "word_encoder_out", "dur", "attn",
                                                            "decoder_inp", "z_p", "mel_out")})
        out[f"{tag}_z_in"] = z.numpy()
        out[f"{tag}_mel2word"] = mine["mel2word"].numpy()
    path = os.path.join(ROOT, "tests", "golden", "ps_full.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, round(os.path.getsize(path) / 1e6, 2), "MB")


if __name__ == "__main__":
    main()
