"""decode_mel alone (g_pre_net -> prior flow -> FVAE decoder) at the cfg-2 shape, fused acoustic kernels on and off
(dtts_debug_set_acoustic_fuse); CUDA-event times outside any profiler.  --ncu: one pass only (the command profiled)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dict_tts_b200 import binding, synth  # noqa: E402
from dict_tts_b200.engine import DictTTSEngine  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=60)
    ap.add_argument("--T", type=int, default=400)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ncu", action="store_true")
    a = ap.parse_args()
    lib = binding.load()
    eng = DictTTSEngine(synth.make_acoustic_state_dict(1234))
    g = torch.Generator().manual_seed(7)
    g_bct = (torch.randn(a.B, 192, a.T, generator=g) * 0.5).cuda()
    z = torch.randn(a.B, 16, a.T // 4, generator=g).cuda()
    if a.ncu:
        eng.decode_mel(g_bct, z)
        torch.cuda.synchronize()
        sys.exit(0)
    for mode in (0, 1, 0, 1):
        lib.dtts_debug_set_acoustic_fuse(mode)
        for _ in range(3):
            eng.decode_mel(g_bct, z)
        n0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            eng.decode_mel(g_bct, z)
        e1.record()
        torch.cuda.synchronize()
        print("decode_mel B=%d T=%d fuse=%d: %.3f ms, %d launches" % (a.B, a.T, mode, e0.elapsed_time(e1) / a.iters,
                                                                    (eng.launches - n0) // a.iters))
    lib.dtts_debug_set_acoustic_fuse(-1)
