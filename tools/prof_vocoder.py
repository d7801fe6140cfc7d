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
    ap.add_argument("--precision", type=int, default=1)
    ap.add_argument("--B", type=int, default=60)
    ap.add_argument("--T", type=int, default=400)
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    eng = HifiGanEngine(synth.make_vocoder_state_dict(4321), precision=a.precision)
    mel = synth.make_mel(7, a.B, a.T).cuda()
    for _ in range(a.iters):
        eng(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng(mel)
    e1.record()
    torch.cuda.synchronize()
    print("vocoder precision %d B %d T %d: %.3f ms" % (a.precision, a.B, a.T, e0.elapsed_time(e1)))
