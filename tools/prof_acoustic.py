"""Runs the acoustic model alone at the cfg-2 shape (B=60, Tw<=22, Lk<=96, T=400) -- the command profiled with ncu."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.engine import DictTTSEngine  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--s2pa-route", type=int, default=0, help="1: K/V projection GEMM on tcgen05 (dtts.h s2pa_route)")
    ap.add_argument("--alias", action="store_true", help="values IS keys (one tensor), as the binarized data has it")
    a = ap.parse_args()
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234), s2pa_route=a.s2pa_route)
    b = synth.make_batch(seed=1234, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96, alias_values=a.alias)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.items()}
    if a.alias:
        dev["values"] = dev["keys"]
    print("valid gloss tokens:", int((b["key_map"] != 0).sum()), "of", b["key_map"].numel())

    def once():
        return eng.forward((dev["word_tokens"],), dev["pron_modified"],
                           dict_msg=(dev["keys"], dev["values"], dev["key_map"], dev["pinyin"], dev["pinyin_map"]),
                           mel2word=dev["mel2word"], z_p=dev["z_p"])
    for _ in range(a.iters):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    once()
    e1.record()
    torch.cuda.synchronize()
    print("acoustic cfg2 (s2pa_route %d): %.3f ms" % (a.s2pa_route, e0.elapsed_time(e1)))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    eng.text_encode(dev["word_tokens"], dev["pron_modified"], dev["keys"], dev["values"], dev["key_map"], dev["pinyin"],
                    dev["pinyin_map"])
    t1.record()
    torch.cuda.synchronize()
    print("text_encode cfg2 (s2pa_route %d): %.3f ms" % (a.s2pa_route, t0.elapsed_time(t1)))
