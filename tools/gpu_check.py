"""Developer diagnostic (GPU box): per-stage errors of the CUDA path vs the oracle, without asserting."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dict_tts_b200 import synth  # noqa
from dict_tts_b200.config import AcousticConfig, VocoderConfig  # noqa
from dict_tts_b200.engine import DictTTSEngine, HifiGanEngine  # noqa
from dict_tts_b200.weights import fold_weight_norm  # noqa
from oracle import dtts_oracle as O  # noqa
from tests.cases import ACOUSTIC_CASES, VOCODER_CASES  # noqa


def main():
    print(torch.cuda.get_device_name(0))
    sd = synth.make_acoustic_state_dict(1234)
    W = fold_weight_norm(sd)
    cfg = AcousticConfig()
    eng = DictTTSEngine(sd)
    for name, (kw, predicted) in ACOUSTIC_CASES.items():
        gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        batch = synth.make_batch(**kw)
        z = torch.from_numpy(gold["z_in"])
        try:
            out = eng.forward((batch["word_tokens"],), batch["pron_modified"],
                              dict_msg=(batch["keys"], batch["values"], batch["key_map"], batch["pinyin"],
                                        batch["pinyin_map"]),
                              mel2word=None if predicted else batch["mel2word"], z_p=z)
            torch.cuda.synchronize()
        except Exception as e:  # noqa
            print(name, "FAILED", repr(e))
            continue
        m2w_ok = out["mel2word"].shape == gold["mel2word"].shape and np.array_equal(out["mel2word"].cpu().numpy(),
                                                                                     gold["mel2word"])
        errs = {}
        for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur", "decoder_inp", "z_p", "mel_out"):
            a = out[k].cpu().numpy()
            errs[k] = "shape" if a.shape != gold[k].shape else "%.2e" % np.abs(a - gold[k]).max()
        print(name, "mel2word_exact=%s" % m2w_ok, errs)
    vsd = synth.make_vocoder_state_dict(4321)
    voc = HifiGanEngine(vsd)
    for name, kw in VOCODER_CASES.items():
        gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))["wav"]
        wav = voc(synth.make_mel(kw["seed"], kw["B"], kw["T"])).cpu().numpy()
        print(name, "rms %.2e max %.2e" % (np.sqrt(np.mean((wav - gold) ** 2)), np.abs(wav - gold).max()))
    # quick timing at cfg-2 size
    mel = synth.make_mel(1, 60, 400).cuda()
    for _ in range(2):
        voc(mel)
    torch.cuda.synchronize()
    t0 = time.time()
    voc(mel)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("vocoder cfg2 (60x400 frames): %.1f ms -> %.1f TFLOP/s fp32, %.0fx real-time" %
          (dt * 1e3, 24000 * 614.1e6 / dt / 1e12, 24000 * 256 / 22050 / dt))


if __name__ == "__main__":
    main()
