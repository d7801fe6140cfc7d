"""Runs the vocoder alone at the cfg-2 shape (B=60, T=400) -- the command profiled with ncu (profiles/)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.engine import HifiGanEngine  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", type=int, default=6)
    ap.add_argument("--B", type=int, default=60)
    ap.add_argument("--T", type=int, default=400)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--lens", action="store_true",
                    help="vocode up to the valid lengths of the bench batch (dtts_vocode_lens) instead of all T frames")
    a = ap.parse_args()
    eng = HifiGanEngine(synth.make_vocoder_state_dict(4321), precision=a.precision)
    mel = synth.make_mel(7, a.B, a.T).cuda()
    lens = None
    if a.lens:        # the valid lengths of the cfg-2 bench batch (300-400 of 400 frames), scaled to T
        ml = synth.make_batch(seed=1234, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=8)["mel_lengths"]
        lens = (ml[torch.arange(a.B) % 60].float() * a.T / 400).round().int().cuda()
    for _ in range(a.iters):
        eng(mel, lens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng(mel, lens)
    e1.record()
    torch.cuda.synchronize()
    print("vocoder precision %d B %d T %d%s: %.3f ms" % (a.precision, a.B, a.T, " (valid lengths, %d of %d frames)" % (
        int(lens.sum()), a.B * a.T) if a.lens else "", e0.elapsed_time(e1)))
