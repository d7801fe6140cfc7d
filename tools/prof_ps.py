"""Runs the PortaSpeech (non-dict) acoustic model alone at the bench shape (B=60, 12-20 words, T=400) -- for ncu."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.config import PortaSpeechConfig  # noqa: E402
from dict_tts_b200.engine import PortaSpeechEngine  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    pcfg = PortaSpeechConfig()
    peng = PortaSpeechEngine(synth.make_ps_state_dict(2468, pcfg), pcfg, dev, precision=1)
    pb = synth.make_ps_batch(seed=77, B=60, min_words=12, max_words=20, max_ph_per_word=4, max_frames=400, ph_size=pcfg.ph_size)
    pd = {k: v.to(dev) for k, v in pb.items()}
    wl = int(pb["word_lengths"].max())

    def once():
        return peng.forward(pd["txt_tokens"], pd["ph2word"], wl, mel2word=pd["mel2word"], z_p=pd["z_p"])
    for _ in range(a.iters):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    once()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("portaspeech acoustic (B=60, %d phonemes, T=400): %.3f ms" % (int((pb["txt_tokens"] > 0).sum()), e0.elapsed_time(e1)))
