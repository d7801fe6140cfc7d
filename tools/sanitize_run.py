"""One small pass of every CUDA path for compute-sanitizer (tools/sanitize.sh): acoustic model (explicit tensors, bank,
predicted durations, both S2PA routes), vocoder full-length and ragged.  Environment switches (DTTS_TC_PAIR,
DTTS_TC_CLUSTER) are read once per process, so the script is run once per variant."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.bank import DictBank  # noqa: E402
from dict_tts_b200.engine import DictTTSEngine, HifiGanEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vocoder-precision", type=int, default=6)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--skip-acoustic", action="store_true")
    ap.add_argument("--skip-vocoder", action="store_true")
    ap.add_argument("--long", action="store_true", help="also a 2-tile prior flow (T/4 = 120) and the per-layer acoustic build")
    ap.add_argument("--fuse-all", action="store_true",
                    help="dtts_debug_set_tc_fuse(3): every C = 128 ResBlock pair on rb_pair128_kernel (default: k = 3 only)")
    args = ap.parse_args()
    if args.fuse_all:
        from dict_tts_b200 import binding
        assert binding.load().dtts_debug_set_tc_fuse(3) == 0
    if not args.skip_acoustic:
        batch = synth.make_batch(seed=3, B=2, min_chars=3, max_chars=5, max_frames=32, Lk_cap=32)
        for route in (0, 1):
            eng = DictTTSEngine(synth.make_acoustic_state_dict(1234), s2pa_route=route)
            dm = (batch["keys"], batch["values"], batch["key_map"], batch["pinyin"], batch["pinyin_map"])
            eng.forward((batch["word_tokens"],), batch["pron_modified"], dict_msg=dm, mel2word=batch["mel2word"], z_p=batch["z_p"])
            out = eng.forward((batch["word_tokens"],), batch["pron_modified"], dict_msg=dm)          # predicted durations
            if route == 0:
                bank, ids = DictBank.from_batch(batch)
                eng.set_dict_bank(bank)
                eng.forward((batch["word_tokens"],), batch["pron_modified"], dict_ids=ids, mel2word=batch["mel2word"],
                            z_p=batch["z_p"])
                eng.pron_tokens(out["pron_attn"], pinyin=batch["pinyin"])
            if args.long and route == 0:
                from dict_tts_b200 import binding
                g = torch.Generator().manual_seed(1)
                gb = (torch.randn(2, 192, 480, generator=g) * 0.5).cuda()
                zz = torch.randn(2, 16, 120, generator=g).cuda()
                eng.decode_mel(gb, zz)                                   # flow_fused_kernel with two tiles per utterance
                binding.load().dtts_debug_set_acoustic_fuse(0)
                eng.decode_mel(gb, zz)                                   # the per-layer build of the same pass
                eng.forward((batch["word_tokens"],), batch["pron_modified"], dict_msg=dm)
                binding.load().dtts_debug_set_acoustic_fuse(-1)
            torch.cuda.synchronize()
            eng.close()
    if args.skip_vocoder:
        print("sanitize_run ok (acoustic only)")
        return
    voc = HifiGanEngine(synth.make_vocoder_state_dict(4321), precision=args.vocoder_precision)
    mel = synth.make_mel(5, 3, args.frames)
    w = voc(mel)
    w2 = voc(mel, torch.tensor([args.frames, args.frames // 2, 1]))
    pcm_ok = bool(torch.isfinite(w).all()) and bool(torch.isfinite(w2).all())
    torch.cuda.synchronize()
    voc.close()
    print("sanitize_run ok", pcm_ok)


if __name__ == "__main__":
    main()
