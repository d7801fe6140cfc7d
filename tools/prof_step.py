"""One bench step (cfg 2: text_encode -> length regulator -> decode_mel -> vocode, inputs resident in HBM) between
cudaProfilerStart/Stop -- the command profiled with ncu for profiles/*launches*.csv:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv \
        python tools/prof_step.py [--bank]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.pipeline import TextToWav  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--bank", action="store_true", help="characters named by dictionary-bank id (SURVEY.md 8f-1)")
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--vocoder-precision", type=int, default=6)
    a = ap.parse_args()
    pipe = TextToWav(synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321),
                     vocoder_precision=a.vocoder_precision)
    batch = synth.make_batch(seed=1234, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96)
    if a.bank:
        from dict_tts_b200.bank import DictBank
        bank, ids = DictBank.from_batch(batch)
        pipe.acoustic.set_dict_bank(bank)
        batch = {k: v for k, v in batch.items() if k not in ("keys", "values", "key_map", "pinyin", "pinyin_map")}
        batch["dict_ids"] = ids
    dev = pipe.to_device(batch)
    for _ in range(a.warm):
        pipe.run_device(dev)
    torch.cuda.synchronize()
    n0 = pipe.launches
    torch.cuda.cudart().cudaProfilerStart()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pipe.run_device(dev)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("one step: %d launches, %.3f ms (not a bench value when run under ncu)" % (pipe.launches - n0, e0.elapsed_time(e1)))
