"""CPU emulation of the tensor-core operand formats of the HiFi-GAN stack (no GPU needed).

Every convolution input (after its leaky-ReLU) and every weight is rounded the way a given vocoder precision mode feeds
the tcgen05 MMA (products exact, fp32 accumulation); conv_post stays fp32 like the CUDA path.  Prints the waveform RMS
error against the exact-fp32 oracle for each mode -- the evidence behind DESIGN.md's "vocoder precision" table.
Test infrastructure: uses oracle/ only as the fp32 checker.

    python tools/precision_study.py [--T 48] [--B 2] [--seeds 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.config import VocoderConfig  # noqa: E402
from dict_tts_b200.weights import fold_weight_norm  # noqa: E402
from oracle import dtts_oracle as O  # noqa: E402


def rnd(x, dt):
    return x.to(dt).to(torch.float32)


def split(x, dt, planes):
    """hi (+ lo) planes of x in dtype dt, returned as their fp32 sum (what the MMAs see in total)."""
    hi = rnd(x, dt)
    if planes == 1:
        return hi
    return hi + rnd(x - hi, dt)


def emulate(W, vcfg, mel, a_dt, a_planes, w_dt, w_planes):
    """hifigan_forward with quantised conv operands (with split operands the CUDA kernel issues hi*hi + hi*lo + lo*hi,
    i.e. everything but lo*lo, ~2^-2p relative).  w_planes may be a function of the layer's C_out."""
    qa = lambda t: split(t, a_dt, a_planes)      # noqa: E731

    def qw(t, c_out=None):
        planes = w_planes(c_out if c_out is not None else t.shape[0]) if callable(w_planes) else w_planes
        return split(t, w_dt, planes)
    x = F.conv1d(qa(mel.transpose(1, 2)), qw(W["conv_pre.weight"]), W["conv_pre.bias"], padding=3)
    nk = len(vcfg.rb_kernels)
    for i, (u, k) in enumerate(zip(vcfg.up_rates, vcfg.up_kernels)):
        x = F.conv_transpose1d(qa(F.leaky_relu(x, 0.1)), qw(W[f"ups.{i}.weight"], W[f"ups.{i}.weight"].shape[1]), W[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        xs = None
        for j, (kr, dils) in enumerate(zip(vcfg.rb_kernels, vcfg.rb_dilations)):
            r = f"resblocks.{i * nk + j}"
            y = x
            for m, d in enumerate(dils):
                t = F.conv1d(qa(F.leaky_relu(y, 0.1)), qw(W[f"{r}.convs1.{m}.weight"]), W[f"{r}.convs1.{m}.bias"],
                             dilation=d, padding=(kr * d - d) // 2)
                t = F.conv1d(qa(F.leaky_relu(t, 0.1)), qw(W[f"{r}.convs2.{m}.weight"]), W[f"{r}.convs2.{m}.bias"],
                             padding=(kr - 1) // 2)
                y = t + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)
    return torch.tanh(F.conv1d(x, W["conv_post.weight"], W["conv_post.bias"], padding=3)).squeeze(1)


MODES = [
    # name, a_dtype, a_planes, w_dtype, w_planes, MMAs per product
    ("bf16 x bf16            (1 MMA)", torch.bfloat16, 1, torch.bfloat16, 1),
    ("fp16 x fp16            (1 MMA)", torch.float16, 1, torch.float16, 1),
    ("fp16 x fp16 hi+lo      (2 MMA)", torch.float16, 1, torch.float16, 2),
    ("fp16 hi+lo x fp16      (2 MMA)", torch.float16, 2, torch.float16, 1),
    ("fp16 x fp16 hi+lo only C_out<=64", torch.float16, 1, torch.float16, lambda c: 2 if c <= 64 else 1),
    ("fp16 x fp16 hi+lo only C_out<=128", torch.float16, 1, torch.float16, lambda c: 2 if c <= 128 else 1),
    ("fp16 x fp16 hi+lo only C_out>=128", torch.float16, 1, torch.float16, lambda c: 2 if c >= 128 else 1),
    ("bf16 hi+lo x bf16 hi+lo (3 MMA)", torch.bfloat16, 2, torch.bfloat16, 2),
    ("fp16 hi+lo x fp16 hi+lo (3 MMA)", torch.float16, 2, torch.float16, 2),
]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=48)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--seeds", type=int, default=3)
    a = ap.parse_args()
    torch.set_grad_enabled(False)
    vcfg = VocoderConfig()
    print("%-34s %s" % ("mode", "wav RMS error vs fp32 per weight seed (tolerance 1e-4)"))
    rows = {m[0]: [] for m in MODES}
    amax = []
    for s in range(a.seeds):
        W = fold_weight_norm(synth.make_vocoder_state_dict(4321 + s))
        mel = synth.make_mel(100 + s, a.B, a.T)
        ref = O.hifigan_forward(W, vcfg, mel)
        amax.append(float(ref.abs().max()))
        for name, adt, ap_, wdt, wp in MODES:
            out = emulate(W, vcfg, mel, adt, ap_, wdt, wp)
            rows[name].append(float((out - ref).pow(2).mean().sqrt()))
    for name, v in rows.items():
        print("%-34s %s" % (name, "  ".join("%.2e" % e for e in v)))
    print("reference |wav| max per seed:", "  ".join("%.3f" % v for v in amax))
