"""TEST INFRASTRUCTURE ONLY -- run in the build container (needs /root/reference).

1. Loads the synthetic checkpoints (dict_tts_b200/synth.py) into the UNMODIFIED reference modules and runs
   them on seeded synthetic batches.
2. Asserts oracle/dtts_oracle.py reproduces every stage.
3. Writes tests/golden/*.npz from the REFERENCE outputs (inputs are regenerated from seeds at test time).

    python -m oracle.make_golden
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.config import AcousticConfig, VocoderConfig  # noqa: E402
from dict_tts_b200.weights import fold_weight_norm  # noqa: E402
from oracle import dtts_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, make_batch kwargs, predicted durations?)
ACOUSTIC_CASES = [
    ("ac_small", dict(seed=11, B=3, min_chars=3, max_chars=7, max_frames=48, Lk_cap=40), False),
    ("ac_ragged", dict(seed=12, B=4, min_chars=1, max_chars=9, max_frames=64, Lk_cap=64, pron_modified_p=0.1), False),
    ("ac_preddur", dict(seed=13, B=3, min_chars=2, max_chars=6, max_frames=40, Lk_cap=32), True),
    ("ac_single", dict(seed=14, B=1, min_chars=12, max_chars=12, max_frames=120, Lk_cap=96), False),
]
VOCODER_CASES = [("voc_small", dict(seed=21, B=2, T=24)), ("voc_single", dict(seed=22, B=1, T=57))]
# "trained-like" dynamic range: the same generator with its internal activations scaled by `hot`
# (synth.make_vocoder_state_dict): 1e2 -> stage-1 activations of 1e2..1e3; 2e4 -> past the fp16 range
HOT_VOCODER_CASES = [("voc_hot", dict(seed=23, B=1, T=40), 100.0), ("voc_overflow", dict(seed=24, B=1, T=24), 20000.0)]


def build_reference_models():
    sd = synth.make_acoustic_state_dict(1234)
    model, voc = ref_loader.build_models(sd, synth.make_vocoder_state_dict(4321))
    return model, voc, sd


def run_reference_acoustic(model, batch, predicted):
    z = batch["z_p"]
    # the reference samples z_p itself from the global CPU generator; feed ours by seeding identically
    import torch.distributions as D
    orig = D.Normal.sample

    def fixed(self, shape=torch.Size()):
        if list(shape) == list(z.shape) or predicted:
            if predicted:
                torch.manual_seed(991)
                return orig(self, shape)
            return z.clone()
        return orig(self, shape)
    D.Normal.sample = fixed
    try:
        with torch.no_grad():
            out = model((batch["word_tokens"], batch["word_tokens"]), batch["pron_modified"], (None, None, None),
                        ph2word=None, word_len=batch["word_lengths"].max(),
                        dict_msg=(batch["keys"], batch["values"], batch["key_map"], batch["pinyin"], batch["pinyin_map"]),
                        infer=True, forward_post_glow=False, spk_embed=None, two_stage=True,
                        mel2word=None if predicted else batch["mel2word"])
    finally:
        D.Normal.sample = orig
    return out


def maxabs(a, b):
    return float((a - b).abs().max())


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    model, voc, sd = build_reference_models()
    cfg = AcousticConfig()
    W = fold_weight_norm(sd)
    for name, kw, predicted in ACOUSTIC_CASES:
        batch = synth.make_batch(**kw)
        ref = run_reference_acoustic(model, batch, predicted)
        T = ref["mel_out"].shape[1]
        if predicted:
            torch.manual_seed(991)
            z = torch.distributions.Normal(0, 1).sample([kw["B"], cfg.latent, T // cfg.frames_multiple])
        else:
            z = batch["z_p"]
        with torch.no_grad():
            mine = O.acoustic_forward(W, cfg, batch, None if predicted else batch["mel2word"], z)
        # integer parts must be exact
        ref_m2w = None
        if predicted:
            dur_int = O.durations_to_int(ref["dur"])
            ilens = (batch["word_tokens"] != 0).sum(-1)
            ref_m2w = O.length_regulate(dur_int, ilens, cfg.frames_multiple)
            assert torch.equal(ref_m2w, mine["mel2word"]), "mel2word mismatch"
            assert ref_m2w.shape[1] == T
        errs = {k: maxabs(ref[k], mine[k]) for k in
                ("word_encoder_out", "dict_attn", "pron_attn", "dur", "decoder_inp", "x_mask", "z_p", "mel_out")}
        print(name, "T=%d" % T, {k: "%.2e" % v for k, v in errs.items()},
              "mel|max|=%.3f" % float(ref["mel_out"].abs().max()))
        assert max(errs.values()) < 2e-5, errs
        save = {k: ref[k].numpy() for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "decoder_inp",
                                             "z_p", "mel_out")}
        save["mel2word"] = (mine["mel2word"] if predicted else batch["mel2word"]).numpy()
        save["z_in"] = z.numpy()
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **save)
    vcfg = VocoderConfig()
    Wv = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    for name, kw in VOCODER_CASES:
        mel = synth.make_mel(kw["seed"], kw["B"], kw["T"])
        with torch.no_grad():
            ref = voc(mel.transpose(1, 2)).squeeze(1)
            mine = O.hifigan_forward(Wv, vcfg, mel)
        e = maxabs(ref, mine)
        print(name, "wav err %.2e" % e, "rms %.3f max %.3f" % (float(ref.pow(2).mean().sqrt()), float(ref.abs().max())))
        assert e < 2e-5
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), wav=ref.numpy())
    R = ref_loader.load()
    for name, kw, hot in HOT_VOCODER_CASES:
        sd_hot = synth.make_vocoder_state_dict(4321, hot=hot)
        vh = R["hifigan_cls"](R["voc_cfg"]).eval()
        vh.load_state_dict(sd_hot, strict=True)
        vh.remove_weight_norm()
        mel = synth.make_mel(kw["seed"], kw["B"], kw["T"])
        with torch.no_grad():
            ref = vh(mel.transpose(1, 2)).squeeze(1)
            mine = O.hifigan_forward(fold_weight_norm(sd_hot), vcfg, mel)
        e = maxabs(ref, mine)
        print(name, "hot=%g wav err %.2e" % (hot, e), "rms %.3f max %.3f" % (float(ref.pow(2).mean().sqrt()), float(ref.abs().max())))
        assert e < 5e-5                        # fp32 summation-order noise grows with the internal scale
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), wav=ref.numpy())
    # spec2wav semantics (vocoders/hifigan.py:54-62): one utterance [T,80] -> flat wav
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    main()
