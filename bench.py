"""Benchmark of the Dict-TTS text -> mel -> waveform forward pass (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--quick]

One "step" = one pass of the hot path over one Biaobei-shaped synthetic batch -- cfg 2 of BASELINE.json (batch 60,
<= 22 word tokens, L_k <= 96 gloss tokens, 400 mel frames, full text -> mel -> HiFi-GAN); it is the headline line.
With N > 1 the driver launches this file under torchrun: every rank owns one GPU and one independent batch (weak
scaling), rank 0 broadcasts the weight arenas over NCCL once at load and there is no hot-path collective.
Rank 0 prints ONE JSON line.  At N = 1 the same line also carries (key "configs") the other BASELINE.json
configurations -- cfg 1 (single 64-character utterance, with the reference's CPU path timed on all threads and on one,
min and median), cfg 3 (bf16 tensor-core mode, with its measured error) and cfg 4 (vocoder only,
256 x 32 frames, its own roofline) -- the fp32-class precision mode of cfg 2 ("fp32_class"), the stage rooflines
("roofline_stages") and the PyTorch-eager-on-GPU library baselines.

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the UNMODIFIED reference
modules through oracle/ref_loader.py when a reference tree is present (/root/reference in the build container,
baseline/_ref on a pod that has one; cpu_baseline.kind = "reference"), otherwise the oracle port (kind = "port").  It
runs the way the reference's inference runs: one utterance at a time (tasks/tts/dict_tts.py:229), un-padded.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.config import HOP_SIZE, SAMPLE_RATE, AcousticConfig, VocoderConfig  # noqa: E402
from dict_tts_b200.weights import drop_dead, fold_weight_norm, pack_arena  # noqa: E402

# ---- algorithmic work model (checked against torch FlopCounterMode on the reference) ----
VOCODER_FLOP_PER_FRAME = 614.1e6
VOCODER_BYTES_PER_FRAME = 1344            # fully fused: 80 * 4 in + 256 * 4 out
DECODER_FLOP_PER_FRAME = 4.69e6           # WN decoder 4.09 + prior flow 0.45 + g_pre_net 0.15 MFLOP
DECODER_BYTES_PER_FRAME = 1104            # fully fused: g 768 + z 16 + mel 320
ENCODER_FLOP_PER_TOKEN = 16.6e6           # 8 encoder layers
S2PA_BYTES_PER_GLOSS_TOKEN = 6144         # keys + values rows, fp32 (3072 when values alias keys)
LR_BYTES_PER_FRAME = 1536
# ncu dram__bytes_read + dram__bytes_write of one valid-length vocode pass at cfg 2 (profiles/, see traffic_source)
TC_CONV_DRAM_BYTES_PER_STEP = 61.1e9
TC_CONV_TRAFFIC_SOURCE = ("profiles/r02j_vocoder_lens_dram_agg.txt: ncu dram__bytes_read+write summed over the 52 tensor-core "
                          "launches of one valid-length vocode pass (35 tc_conv_kernel + 3 rb_pair128_kernel + 6 rb_pair64_kernel "
                          "+ 6 rb_pair32_kernel + 2 rb_block_kernel) = 61.1 GB = 1.17 GB per launch (61.3 GB with the conv_post "
                          "finishing kernel; before the k = 3 ResBlocks of the two narrow stages became one launch each: 69.1 GB "
                          "over 56; round 1, before the fused ResBlock pairs: 82.9 GB over 77 launches). Byte models: SURVEY 8d "
                          "fully-fused algorithmic 1344 B/frame = 27.8 MB per step; this design's per-layer model (16 B per "
                          "element and ResBlock pair, 12 B where the pair is fused, 10 B for a block-fused ResBlock) = 61 GB "
                          "per step -- the measured traffic equals the per-layer model, i.e. ~2200x the fully-fused figure: "
                          "that factor is what a layer-by-layer design costs, not re-reads")
WORKLOADS = {
    "cfg1": dict(B=1, min_chars=64, max_chars=64, max_frames=1280, Lk_cap=96),
    "cfg2": dict(B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96),
}
CFG4 = dict(B=256, T=32)
# tensor-core launches of one vocode pass: conv_pre, 4 transposed convolutions, 30 ResBlock convolutions of stages 1-2, the 3
# fused k = 3 pairs of stage 2 (rb_pair128.cu), the 12 fused ResBlock pairs of stages 3-4 (rb_pair.cu) and the k = 3 ResBlocks of
# stages 3-4 as one launch each (rb_block.cu); 72 ResBlock convolutions = 77 launches before the fusion
N_TC_CONV_LAUNCHES = 1 + 4 + 30 + 3 + 12 + 2
CPU_SAMPLE_UTTS = 30      # ~10 k of the batch's 20.7 k frames: 10-20 s of host work for the two timed passes
VOC_DTYPE = {0: "f32", 1: "f32-class (bf16 hi/lo x hi/lo on tcgen05: 3 MMAs per product, fp32 accumulate)",
             2: "bf16 (tcgen05, 1 MMA)", 3: "fp16 activations x fp16 hi/lo weights (tcgen05, 2 MMAs, fp32 accumulate)",
             4: "fp16 (tcgen05, 1 MMA)", 5: "fp16 x fp16, hi/lo weights only where C_out < 128 (tcgen05)",
             6: "fp16 activations x fp16 hi/lo weights, lo plane of the C_out >= 128 layers as e5m2 (tcgen05, fp32 accumulate)"}
VOC_KERNEL = {0: "conv1d_f32_kernel (fp32 FMA pipe)", 1: "tc_conv_kernel (tcgen05, bf16 hi/lo x hi/lo: 3 MMAs per product)",
              2: "tc_conv_kernel (tcgen05, bf16: 1 MMA)", 3: "tc_conv_kernel (tcgen05, fp16 x fp16 hi/lo weights: 2 MMAs)",
              4: "tc_conv_kernel (tcgen05, fp16: 1 MMA)", 5: "tc_conv_kernel (tcgen05, fp16; hi/lo weights only where C_out < 128)",
              6: "tc_conv_kernel + rb_pair{32,64,128}_kernel (tcgen05, fp16 x fp16 hi/lo weights: 2 MMAs; lo plane in FP8 where "
                 "C_out >= 128: 1.5; ResBlock pairs of the C <= 64 stages and the k = 3 pairs of the C = 128 stage fused, the k = 3 "
                 "ResBlocks of the C = 32 / 64 stages as one launch each)"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    tflops_burst=p["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, tflops_burst=1650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------
# the reference's CPU path
# ---------------------------------------------------------------------------------------------------------------
def slice_utt(batch, b):
    """Utterance b of a collated batch as the reference's own inference sees it: a batch of ONE (max_sentences = 1,
    tasks/tts/dict_tts.py:229), i.e. without the batch's padding in any dimension."""
    n = int(batch["word_lengths"][b])
    T = int(batch["mel_lengths"][b])
    km = batch["key_map"][b:b + 1, :n]
    pm = batch["pinyin_map"][b:b + 1, :n]
    Lk = max(int((km != 0).any(0).any(0).nonzero().max()) + 1, 1)
    Lp = max(int((pm != 0).any(0).any(0).nonzero().max()) + 1, 1)
    keys = batch["keys"][b:b + 1, :n, :Lk].contiguous()
    values = keys if batch["values"] is batch["keys"] else batch["values"][b:b + 1, :n, :Lk].contiguous()
    T4 = (T + 3) // 4
    return dict(word_tokens=batch["word_tokens"][b:b + 1, :n], pron_modified=batch["pron_modified"][b:b + 1, :n],
                keys=keys, values=values, key_map=km[:, :, :Lk].contiguous(),
                pinyin=batch["pinyin"][b:b + 1, :n, :Lp].contiguous(), pinyin_map=pm[:, :, :Lp].contiguous(),
                word_lengths=batch["word_lengths"][b:b + 1], mel2word=batch["mel2word"][b:b + 1, :T],
                z_p=batch["z_p"][b:b + 1, :, :T4].contiguous(), mel_lengths=batch["mel_lengths"][b:b + 1])


class CpuReference:
    """text -> mel -> wav of one utterance on the host: the unmodified reference modules when a reference tree is
    present (kind 'reference'), else the oracle port (kind 'port').  Test / baseline infrastructure only."""

    def __init__(self, force_port=False):
        from oracle import ref_loader
        self.cfg, self.vcfg = AcousticConfig(), VocoderConfig()
        a_sd, v_sd = synth.make_acoustic_state_dict(1234), synth.make_vocoder_state_dict(4321)
        self.kind = "port"
        force_port = force_port or os.environ.get("DTTS_BENCH_FORCE_PORT", "0") == "1"     # A/B: port vs real reference
        if not force_port and ref_loader.available():
            try:
                self.model, self.voc = ref_loader.build_models(a_sd, v_sd)
                self.kind = "reference"
                self.root = ref_loader.REF_ROOT
            except Exception as e:                       # noqa: BLE001 -- a broken tree must not cost the arm
                print(f"reference tree at {ref_loader.REF_ROOT} not usable ({e!r}); timing the oracle port", file=sys.stderr)
        if self.kind == "port":
            self.W, self.Wv = fold_weight_norm(a_sd), fold_weight_norm(v_sd)

    @torch.no_grad()
    def run_utt(self, u):
        """PortaSpeech_dict.forward(infer=True) with supplied durations (profile_infer, dict_tts.py:194) at B = 1, then
        spec2wav on the un-padded mel (vocoders/hifigan.py:54-62)."""
        T = int(u["mel_lengths"][0])
        if self.kind == "reference":
            import torch.distributions as D
            orig, z = D.Normal.sample, u["z_p"]
            D.Normal.sample = lambda self_, shape=torch.Size(): z.clone()     # the engine arm is given the same prior sample
            try:
                ret = self.model((u["word_tokens"], u["word_tokens"]), u["pron_modified"], (None, None, None),
                                 ph2word=None, word_len=u["word_lengths"].max(),
                                 dict_msg=(u["keys"], u["values"], u["key_map"], u["pinyin"], u["pinyin_map"]),
                                 infer=True, forward_post_glow=False, spk_embed=None, two_stage=True,
                                 mel2word=u["mel2word"])
            finally:
                D.Normal.sample = orig
            mel = ret["mel_out"][0, :T]                                          # after_infer hands spec2wav one [T,80] mel
            return self.voc(mel.T.unsqueeze(0)).view(-1)
        from oracle import dtts_oracle as O
        ret = O.acoustic_forward(self.W, self.cfg, u, u["mel2word"], u["z_p"])
        return O.hifigan_forward(self.Wv, self.vcfg, ret["mel_out"][:, :T]).view(-1)

    def time_passes(self, utts, threads, iters, warm_utts=1):
        """Seconds of `iters` passes over the list of utterances (after `warm_utts` untimed ones)."""
        torch.set_num_threads(threads)
        for u in utts[:warm_utts]:
            self.run_utt(u)
        secs = []
        for _ in range(iters):
            t0 = time.perf_counter()
            for u in utts:
                self.run_utt(u)
            secs.append(time.perf_counter() - t0)
        return secs


def cpu_stats(secs, frames):
    return dict(min_s=min(secs), median_s=statistics.median(secs), iters=len(secs),
                frames_per_s=frames / min(secs), frames_per_s_median=frames / statistics.median(secs),
                x_realtime=frames * HOP_SIZE / SAMPLE_RATE / min(secs))


# ---------------------------------------------------------------------------------------------------------------
# PyTorch eager on the GPU: the "library kernels" bar (cuDNN / cuBLAS), SURVEY.md §8d
# ---------------------------------------------------------------------------------------------------------------
def eager_gpu_run(batch, dev, mode, iters=2):
    """The restatement of the reference's PyTorch forward (oracle/dtts_oracle.py), run as PyTorch eager ON THE GPU, whole
    cfg-2 batch, vocoder batched (kinder than the reference's one utterance at a time).  mode: 'fp32' (TF32 off, what
    the reference's amp: false means), 'tf32' (cuDNN / cuBLAS allowed to use TF32 tensor cores), 'fp16' (autocast).
    Returns (frames/s, ms)."""
    from oracle import dtts_oracle as O
    tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    cfg, vcfg = AcousticConfig(), VocoderConfig()
    W = {k: v.to(dev) for k, v in fold_weight_norm(synth.make_acoustic_state_dict(1234)).items()}
    Wv = {k: v.to(dev) for k, v in fold_weight_norm(synth.make_vocoder_state_dict(4321)).items()}
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    frames = int(batch["mel_lengths"].sum())

    def once():
        with torch.no_grad():
            ret = O.acoustic_forward(W, cfg, b, b["mel2word"], b["z_p"])
            with torch.autocast("cuda", dtype=torch.float16, enabled=(mode == "fp16")):
                return O.hifigan_forward(Wv, vcfg, ret["mel_out"]).float()
    once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return frames / (ms / 1e3), ms


def claim_stdout():
    """stdout carries exactly ONE line (the JSON record): whatever libraries print there while the bench runs (NCCL's
    version banner, for one) is sent to stderr; returns the function that writes the record to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        sys.stdout.flush()
        os.write(real, (line + "\n").encode())
    return emit


def base_config(world):
    return dict(workload="cfg2: Biaobei-shaped batch=60/GPU, Tw<=22, Lk<=96, T=400 frames, full text->mel->HiFi-GAN "
                         "(reference arithmetic: fp32; see `dtype` / `precision` for this arm's operand formats)",
                batch_per_gpu=WORKLOADS["cfg2"]["B"], frames_per_utt=WORKLOADS["cfg2"]["max_frames"],
                parallelism=f"replicas x{world}", supplied_durations=True,
                l2="every step's inputs and activations (>= 0.4 GB of keys, 80 GB of vocoder traffic) exceed the 126 MB L2")


def reference_arm(args, emit, cores, world):
    """bench.py --impl reference: the reference's CPU path, all host threads, K steps of a bounded sample each."""
    batch = synth.make_batch(seed=1234, alias_values=True, **WORKLOADS["cfg2"])
    ref = CpuReference()
    torch.set_num_threads(cores)
    utts_all = [slice_utt(batch, b) for b in range(batch["word_tokens"].shape[0])]
    t0 = time.perf_counter()
    ref.run_utt(utts_all[0])                                     # first call: thread-pool / primitive-cache warm-up
    t0 = time.perf_counter()
    ref.run_utt(utts_all[0])
    t_utt = time.perf_counter() - t0
    K, Wm = max(args.steps, 1), max(args.warmup, 0)
    # bounded sample per step: the whole K + W run stays around two minutes on this host
    n = int(max(1, min(len(utts_all), 120.0 / max(t_utt, 1e-3) / (K + Wm))))
    utts = utts_all[:n]
    frames = sum(int(u["mel_lengths"][0]) for u in utts)
    for _ in range(Wm):
        for u in utts:
            ref.run_utt(u)
    secs = []
    for _ in range(K):
        t0 = time.perf_counter()
        for u in utts:
            ref.run_utt(u)
        secs.append(time.perf_counter() - t0)
    total = sum(secs)
    fps = frames * K / total
    audio = frames * HOP_SIZE / SAMPLE_RATE
    sample = (f"each step = the first {n} utterances ({frames} valid frames) of the cfg-2 batch, one utterance at a time, "
              f"un-padded, as the reference's inference loop runs them; {K} timed steps on {cores} host threads")
    emit(json.dumps(dict(
        impl="reference", metric="mel_frames_per_s", value=fps, unit="frames/s", n_gpus=args.gpus, steps=K,
        warmup=Wm, ms_per_step=total / K * 1e3, higher_is_better=True, scaling="weak",
        vs_baseline=None, dtype="f32", data="synthetic", config=base_config(world), rtf=(total / K) / audio,
        step_s_min=min(secs), step_s_median=statistics.median(secs),
        cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, kind=ref.kind, sample=sample,
                          reference_root=getattr(ref, "root", None)),
        e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline line only: no other configs, no CPU / eager baselines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip timing the PyTorch-eager (cuDNN/cuBLAS) forward of the same batch on the GPU")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip cfg 1 / 3 / 4 and the fp32-class arm")
    ap.add_argument("--vocoder-precision", type=int, default=int(os.environ.get("DTTS_VOCODER_PRECISION", "6")),
                    help="0: fp32 FMA pipe; 1: tcgen05 bf16 hi/lo x hi/lo (3 MMAs, ~1e-6 wav RMS); 2: tcgen05 bf16 (cfg 3, "
                         "outside the tolerance); 3: tcgen05 fp16 x fp16 hi/lo weights (2 MMAs, ~6e-5 wav RMS, "
                         "inside the 1e-4 tolerance); 4: tcgen05 fp16 (1 MMA, ~8.5e-5); 5: as 3 with single-plane weights "
                         "in the C_out >= 128 layers (~7.5e-5); 6 (default): as 3 with the lo plane of the C_out >= 128, k >= 7 layers as an FP8 MMA "
                         "(same error as 3)")
    ap.add_argument("--acoustic-precision", type=int, default=int(os.environ.get("DTTS_ACOUSTIC_PRECISION", "1")),
                    help="0: fp32 FMA pipe; 1 (default): dense convolutions on tcgen05, bf16 hi/lo split (fp32-class)")
    args = ap.parse_args()
    if args.quick:
        args.no_cpu_baseline = args.no_eager_baseline = args.no_extra_configs = True
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    emit = claim_stdout()
    config = base_config(world)

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, emit, cores, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dict_tts_b200.bank import DictBank
    from dict_tts_b200.engine import HifiGanEngine
    from dict_tts_b200.pipeline import TextToWav

    # ---- weights: rank 0 builds the arenas, one NCCL broadcast each ----
    acfg, vcfg = AcousticConfig(), VocoderConfig()
    a_sd = drop_dead(fold_weight_norm(synth.make_acoustic_state_dict(1234)))
    v_sd = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    a_host, a_table = pack_arena(a_sd)
    v_host, v_table = pack_arena(v_sd)
    bcast_ms = None
    if world > 1:
        a_arena = a_host.to(dev) if rank == 0 else torch.empty_like(a_host, device=dev)
        v_arena = v_host.to(dev) if rank == 0 else torch.empty_like(v_host, device=dev)
        warm = torch.zeros(1, device=dev)
        dist.broadcast(warm, 0)                                   # communicator set-up is not the broadcast
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        dist.broadcast(a_arena, 0)
        dist.broadcast(v_arena, 0)
        b1.record()
        torch.cuda.synchronize()
        bcast_ms = b0.elapsed_time(b1)
    else:
        a_arena, v_arena = a_host.to(dev), v_host.to(dev)
    arenas = ((a_arena, a_table), (v_arena, v_table))

    def make_pipe(vp, ap_=args.acoustic_precision, route=0):
        return TextToWav(None, None, acfg, vcfg, dev, arenas=arenas, vocoder_precision=vp, acoustic_precision=ap_,
                         s2pa_route=route)
    pipe = make_pipe(args.vocoder_precision)

    # values IS keys, as in the binarized data (binarizer_zh.py:231-233): one tensor, uploaded once
    batch = synth.make_batch(seed=1234 + rank, alias_values=True, **WORKLOADS["cfg2"])
    frames = int(batch["mel_lengths"].sum())
    padded_frames = batch["mel2word"].shape[0] * batch["mel2word"].shape[1]
    n_tokens = batch["word_tokens"].numel()
    gloss_valid = int((batch["key_map"] != 0).sum())
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items() if k != "values"}
    host["values"] = host["keys"]
    devb = pipe.to_device(host)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream()
    stage_names = ("text_encode", "length_regulate", "decode_mel", "vocode")

    def time_device(p, d, steps, warm):
        """Device-resident arm: (ms per step max over ranks, per-stage ms, launches in the timed region)."""
        for _ in range(max(warm, 3)):
            p.run_device(d)
        barrier()
        marks = []
        l0 = p.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(steps):
            sm = [torch.cuda.Event(enable_timing=True)]
            sm[0].record(stream)

            def record(name, sm=sm):
                e = torch.cuda.Event(enable_timing=True)
                e.record(stream)
                sm.append(e)
            p.run_device(d, record)
            marks.append(sm)
        ev1.record(stream)
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        st = {n: sum(m[i].elapsed_time(m[i + 1]) for m in marks) / steps for i, n in enumerate(stage_names)}
        return ms, st, p.launches - l0

    def time_stream(p, hb, steps, bufs):
        """End-to-end arm: pinned host inputs -> H2D -> engine -> wav D2H every step through the public API
        (synthesize_stream: the upload of step i+1 overlaps the compute of step i)."""
        def gen(n):
            for _ in range(n):
                yield hb
        for _ in p.synthesize_stream(gen(2), bufs):
            pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_out = sum(1 for _ in p.synthesize_stream(gen(steps), bufs))
        e1.record(stream)
        barrier()
        assert n_out == steps
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    def time_serial(p, hb, out, steps):
        for _ in range(2):
            p.synthesize(hb, out)
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        for _ in range(steps):
            p.synthesize(hb, out)
        l1.record(stream)
        barrier()
        return max_over_ranks(l0.elapsed_time(l1)) / steps

    # ---- headline: device-resident arm ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms, stage_ms, launches = time_device(pipe, devb, args.steps, args.warmup)

    # ---- end-to-end arms ----
    B, Tmax = WORKLOADS["cfg2"]["B"], WORKLOADS["cfg2"]["max_frames"]
    wav_bufs = [torch.empty(B, Tmax * HOP_SIZE, dtype=torch.float32, pin_memory=True) for _ in range(2)]
    wav_host = wav_bufs[0]
    d2h = wav_host.numel() * wav_host.element_size()
    # (a) GPU-resident dictionary bank -- the task's default mode: characters are named by id, the bank
    # is uploaded once like the weights, per step only ids / durations / noise cross the bus.  This is the headline e2e.
    bank, dict_ids = DictBank.from_batch(batch)
    pipe.acoustic.set_dict_bank(bank)
    slim = {k: v for k, v in host.items() if k not in ("keys", "values", "key_map", "pinyin", "pinyin_map")}
    slim["dict_ids"] = dict_ids.pin_memory()
    bank_ms = time_stream(pipe, slim, args.steps, wav_bufs)
    bank_serial_ms = time_serial(pipe, slim, wav_host, min(args.steps, 5))
    bank_h2d = sum(slim[k].numel() * slim[k].element_size() for k in
                   ("word_tokens", "pron_modified", "dict_ids", "mel2word", "z_p"))
    # (b) the reference collater's padded [B,Tw,Lk,768] tensors from pinned host memory every step (keys uploaded once
    # for keys and values, as the data holds one tensor for both)
    pad_ms = time_stream(pipe, host, args.steps, wav_bufs)
    pad_serial_ms = time_serial(pipe, host, wav_host, min(args.steps, 5))
    pad_h2d = sum(host[k].numel() * host[k].element_size() for k in
                  ("word_tokens", "pron_modified", "keys", "key_map", "pinyin", "pinyin_map", "mel2word", "z_p"))
    clocks = sampler.stop() if rank == 0 else None

    total_frames = frames
    if world > 1:
        t = torch.tensor([frames], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_frames = int(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    audio_total = total_frames * HOP_SIZE / SAMPLE_RATE

    def voc_roofline(vp, voc_ms, n_frames, n_launch):
        ach = n_frames * VOCODER_FLOP_PER_FRAME / (voc_ms / 1e3) / 1e12
        return dict(bound="tensor", kernel="%s, HiFi-GAN stack, %d launches/step" % (VOC_KERNEL[vp], n_launch),
                    achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"],
                    peak_source=peaks["source"] + " bf16 dense (sustained)", avg_launch_ms=voc_ms / n_launch,
                    flop_per_launch=n_frames * VOCODER_FLOP_PER_FRAME / n_launch)

    n_voc_launch = N_TC_CONV_LAUNCHES
    n_voc_launch_unfused = 1 + 4 + 72                  # precisions without stacked hi | lo weights have no fused pair / block kernels
    # the vocoder is run up to each utterance's valid length (dtts_vocode_lens): algorithmic work = valid frames only
    roofline = voc_roofline(args.vocoder_precision, stage_ms["vocode"], frames, n_voc_launch)
    roofline["traffic"] = TC_CONV_DRAM_BYTES_PER_STEP / n_voc_launch if args.vocoder_precision in (3, 6) else None
    roofline["traffic_algorithmic_fully_fused_per_step"] = frames * VOCODER_BYTES_PER_FRAME
    roofline["traffic_source"] = TC_CONV_TRAFFIC_SOURCE

    def hbm(nbytes, ms):
        a = nbytes / (ms / 1e3) / 1e9
        return dict(achieved=a, peak=peaks["hbm_gbs"], unit="GB/s", frac=a / peaks["hbm_gbs"], bytes=nbytes)

    def tensor(flop, ms):
        a = flop / (ms / 1e3) / 1e12
        return dict(achieved=a, peak=peaks["tflops"], unit="TFLOP/s", frac=a / peaks["tflops"], flop=flop)
    s2pa_bytes = gloss_valid * S2PA_BYTES_PER_GLOSS_TOKEN // 2          # values alias keys: each valid row is read twice from
    roofline_stages = dict(                                              # HBM/L2, but it is ONE tensor: 3072 B algorithmic
        model="SURVEY.md 8d algorithmic work per unit x the units of one step, over the stage's CUDA-event time",
        text_encode=dict(ms=stage_ms["text_encode"], tokens=n_tokens, valid_gloss_tokens=gloss_valid,
                         tensor=tensor(n_tokens * ENCODER_FLOP_PER_TOKEN, stage_ms["text_encode"]),
                         hbm=hbm(s2pa_bytes, stage_ms["text_encode"]),
                         note="latency-bound stage: 8 encoder layers over 1 320 tokens; the S2PA pass alone is timed in profiles/"),
        length_regulate=dict(ms=stage_ms["length_regulate"], hbm=hbm(padded_frames * LR_BYTES_PER_FRAME, stage_ms["length_regulate"])),
        decode_mel=dict(ms=stage_ms["decode_mel"], frames=padded_frames,
                        tensor=tensor(padded_frames * DECODER_FLOP_PER_FRAME, stage_ms["decode_mel"]),
                        hbm=hbm(padded_frames * DECODER_BYTES_PER_FRAME, stage_ms["decode_mel"]),
                        note="fully-fused byte model (g 768 + z 16 + mel 320 B per frame); padded frames are computed, as "
                             "the batched reference computes them"),
        vocode=dict(ms=stage_ms["vocode"], frames=frames, tensor=tensor(frames * VOCODER_FLOP_PER_FRAME, stage_ms["vocode"]),
                    hbm=hbm(frames * VOCODER_BYTES_PER_FRAME, stage_ms["vocode"])))

    line = dict(metric="mel_frames_per_s", value=total_frames / (dev_ms / 1e3), unit="frames/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=dev_ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype=VOC_DTYPE[args.vocoder_precision] + "; acoustic model: "
                + ("bf16 hi/lo x hi/lo on tcgen05 (fp32-class)" if args.acoustic_precision else "f32"),
                data="synthetic", config=config,
                precision=dict(vocoder=args.vocoder_precision, acoustic=args.acoustic_precision,
                               vocoder_frames="valid frames + receptive-field margin (padded tail of the batch skipped)"),
                rtf=(dev_ms / 1e3) / audio_total, x_realtime=audio_total / (dev_ms / 1e3),
                stages_ms=stage_ms, gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                roofline_stages=roofline_stages,
                e2e=dict(value=total_frames / (bank_ms / 1e3), unit="frames/s", h2d_bytes_per_step=int(bank_h2d),
                         d2h_bytes_per_step=int(d2h), ms_per_step=bank_ms, x_realtime=audio_total / (bank_ms / 1e3),
                         api="TextToWav.synthesize_stream on host batches that name their characters by dictionary-bank "
                             "id (SURVEY.md 8f-1, the task's default mode; the bank is resident like the weights); copy of "
                             "step i+1 overlaps compute of step i",
                         serial_ms_per_step=bank_serial_ms, bank_bytes_resident=int(bank.nbytes)),
                e2e_padded=dict(value=total_frames / (pad_ms / 1e3), unit="frames/s", ms_per_step=pad_ms,
                                h2d_bytes_per_step=int(pad_h2d), d2h_bytes_per_step=int(d2h),
                                serial_ms_per_step=pad_serial_ms,
                                note="same call with the reference collater's padded [B,Tw,Lk,768] dictionary features "
                                     "from pinned host memory every step (one tensor for keys and values, as the data has it)"))
    if bcast_ms is not None:
        line["arena_broadcast_ms"] = bcast_ms
        line["arena_broadcast_bytes"] = int((a_host.numel() + v_host.numel()) * 4)

    extras = world == 1
    # ---- cfg 2 in the fp32-class precision mode (vocoder precision 1: bf16 hi/lo x hi/lo, ~1e-6 waveform RMS) ----
    if extras and not args.no_extra_configs:
        try:
            steps2 = max(3, min(args.steps, 10))
            p1 = make_pipe(1)
            ms1, st1, _ = time_device(p1, devb, steps2, 3)
            r1, w1 = p1.run_device(devb)
            ref_mel, ref_wav = r1["mel_out"].clone(), w1.clone()
            n_valid = frames * HOP_SIZE                               # both arms write 0 past each utterance's valid length

            def wav_rms(a, b):
                return float(((a - b).double().pow(2).sum() / n_valid).sqrt())
            line["fp32_class"] = dict(vocoder_precision=1, ms_per_step=ms1, value=frames / (ms1 / 1e3), stages_ms=st1,
                                      roofline=voc_roofline(1, st1["vocode"], frames, n_voc_launch_unfused),
                                      note="same workload with every tensor-core product as a 3-MMA bf16 hi/lo split: "
                                           "waveform within ~1e-6 RMS of the reference's fp32 forward (tests)")
            p1.close()
            # default mode against it, on the bench batch itself
            r6, w6 = pipe.run_device(devb)
            line["precision"]["wav_rms_vs_fp32_class"] = wav_rms(w6, ref_wav)
            line["precision"]["wav_signal_rms"] = wav_rms(ref_wav, torch.zeros_like(ref_wav))
            line["precision"]["mel_maxabs_vs_fp32_class"] = float((r6["mel_out"] - ref_mel).abs().max())
            # ---- cfg 3: bf16 tensor-core mode (single bf16 MMA in the vocoder, dict-attention as the K/V projection GEMM) ----
            p3 = make_pipe(2, 1, 1)
            ms3, st3, _l = time_device(p3, devb, steps2, 3)
            r3, w3 = p3.run_device(devb)
            line.setdefault("configs", {})["cfg3"] = dict(
                workload="cfg3: the cfg-2 batch with single-pass bf16 tensor-core convolutions in the vocoder (1 MMA per "
                         "product) and S2PA as the [B*Tw*Lk,768]x[768,384] K/V projection GEMM on tcgen05 (s2pa_route=1)",
                ms_per_step=ms3, value=frames / (ms3 / 1e3), unit="frames/s", stages_ms=st3,
                roofline=voc_roofline(2, st3["vocode"], frames, n_voc_launch_unfused),
                error=dict(wav_rms_vs_fp32_class=wav_rms(w3, ref_wav),
                           mel_maxabs_vs_fp32_class=float((r3["mel_out"] - ref_mel).abs().max()),
                           tolerance="1e-4 RMS on the waveform: a throughput mode, OUTSIDE the tolerance (DESIGN.md §5)"))
            p3.close()
            del p3, p1
        except Exception as e:                                   # an extra must never cost the headline
            line.setdefault("configs", {})["cfg3"] = dict(unavailable=repr(e)[:300])
        # ---- cfg 4: vocoder only, 256 x 32 frames (8192-sample segments) ----
        try:
            mel4 = synth.make_mel(4, CFG4["B"], CFG4["T"])
            mel4_h = mel4.pin_memory()
            mel4_d = mel4.to(dev)
            voc = pipe.vocoder
            for _ in range(3):
                voc(mel4_d)
            torch.cuda.synchronize()
            n4 = max(10, args.steps)
            l0 = voc.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n4):
                voc(mel4_d)
            e1.record(stream)
            torch.cuda.synchronize()
            ms4 = e0.elapsed_time(e1) / n4
            l4 = (voc.launches - l0) // n4
            out4 = torch.empty(CFG4["B"], CFG4["T"] * HOP_SIZE, pin_memory=True)
            e0.record(stream)
            for _ in range(n4):
                out4.copy_(voc(mel4_h.to(dev, non_blocking=True)), non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize()
            ms4e = e0.elapsed_time(e1) / n4
            fr4 = CFG4["B"] * CFG4["T"]
            r4 = voc_roofline(args.vocoder_precision, ms4, fr4, n_voc_launch)
            r4["hbm_fully_fused"] = hbm(fr4 * VOCODER_BYTES_PER_FRAME, ms4)
            line.setdefault("configs", {})["cfg4"] = dict(
                workload="cfg4: HiFi-GAN only, spec2wav_batch on 256 x 32-frame mels (8192-sample segments), mel ~ U(-6,1.5)",
                ms_per_step=ms4, value=fr4 / (ms4 / 1e3), unit="frames/s", launches_per_step=int(l4),
                x_realtime=fr4 * HOP_SIZE / SAMPLE_RATE / (ms4 / 1e3), roofline=r4,
                e2e=dict(value=fr4 / (ms4e / 1e3), unit="frames/s", ms_per_step=ms4e,
                         h2d_bytes_per_step=int(mel4.numel() * 4), d2h_bytes_per_step=int(out4.numel() * 4)),
                note="short segments: every layer is a fraction of a wave of row tiles; L2-resident between layers")
        except Exception as e:
            line.setdefault("configs", {})["cfg4"] = dict(unavailable=repr(e)[:300])
        # ---- cfg 1: one 64-character utterance, T = 1280 frames: GPU latency + the reference's CPU path ----
        try:
            b1 = synth.make_batch(seed=1234, alias_values=True, **WORKLOADS["cfg1"])
            fr1 = int(b1["mel_lengths"].sum())
            h1 = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b1.items() if k != "values"}
            h1["values"] = h1["keys"]
            d1 = pipe.to_device(h1)
            ms1d, st1d, l1d = time_device(pipe, d1, max(10, args.steps), 3)
            out1 = torch.empty(1, WORKLOADS["cfg1"]["max_frames"] * HOP_SIZE, pin_memory=True)
            ms1e = time_serial(pipe, h1, out1, max(10, args.steps))
            audio1 = fr1 * HOP_SIZE / SAMPLE_RATE
            c1 = dict(workload="cfg1: Biaobei dict_tts.yaml, single 64-character utterance (Tw=66), 1280 mel frames = "
                               "14.9 s of audio, text->mel->wav",
                      frames=fr1, audio_s=audio1,
                      gpu=dict(device_ms=ms1d, stages_ms=st1d, launches=int(l1d) // max(10, args.steps),
                               e2e_latency_ms=ms1e, x_realtime=audio1 / (ms1e / 1e3), frames_per_s=fr1 / (ms1e / 1e3),
                               api="TextToWav.synthesize: pinned host batch -> H2D -> engine -> wav D2H, synchronised"))
            if not args.no_cpu_baseline:
                refc = CpuReference()
                u1 = [slice_utt(b1, 0)]
                allt = refc.time_passes(u1, cores, 5, warm_utts=1)
                onet = refc.time_passes(u1, 1, 3, warm_utts=0)      # one thread: the primitives are warm from the passes above
                torch.set_num_threads(cores)
                c1["cpu"] = dict(kind=refc.kind, protocol="one utterance, B=1, exactly as after_infer calls spec2wav "
                                 "(vocoders/hifigan.py:54-62); 1 warm-up + 5 timed passes on all threads; 3 timed passes on 1 thread",
                                 all_threads=dict(cores=cores, **cpu_stats(allt, fr1)),
                                 one_thread=dict(cores=1, **cpu_stats(onet, fr1)))
                c1["gpu_vs_cpu_all_threads"] = min(allt) / (ms1e / 1e3)
            line.setdefault("configs", {})["cfg1"] = c1
        except Exception as e:
            line.setdefault("configs", {})["cfg1"] = dict(unavailable=repr(e)[:300])

    # ---- the PortaSpeech (non-dict) sibling on the same workload shape (B=60, ~50 phonemes, 400 frames) ----
    if extras and not args.no_extra_configs:
        try:
            from dict_tts_b200.config import PortaSpeechConfig
            from dict_tts_b200.engine import PortaSpeechEngine
            pcfg = PortaSpeechConfig()
            p_sd = synth.make_ps_state_dict(2468, pcfg)
            peng = PortaSpeechEngine(p_sd, pcfg, dev, precision=args.acoustic_precision)
            pb = synth.make_ps_batch(seed=77, B=60, min_words=12, max_words=20, max_ph_per_word=4, max_frames=400,
                                     ph_size=pcfg.ph_size)
            pd = {k: v.to(dev) for k, v in pb.items()}
            pframes = int(pb["mel_lengths"].sum())
            wl = int(pb["word_lengths"].max())

            def ps_step():
                out = peng.forward(pd["txt_tokens"], pd["ph2word"], wl, mel2word=pd["mel2word"], z_p=pd["z_p"])
                return pipe.vocoder(out["mel_out"], (out["mel2word"] > 0).sum(-1))
            for _ in range(3):
                ps_step()
            torch.cuda.synchronize()
            nps = max(10, args.steps)
            l0 = peng.launches
            e0, e1, em = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ac_ms = 0.0
            e0.record(stream)
            for _ in range(nps):
                ps_step()
            e1.record(stream)
            torch.cuda.synchronize()
            ps_ms = e0.elapsed_time(e1) / nps
            e0.record(stream)
            for _ in range(nps):
                peng.forward(pd["txt_tokens"], pd["ph2word"], wl, mel2word=pd["mel2word"], z_p=pd["z_p"])
            em.record(stream)
            torch.cuda.synchronize()
            ac_ms = e0.elapsed_time(em) / nps
            ps_line = dict(workload="PortaSpeech (non-dict) sibling, SURVEY 8f-3: batch=60, 12-20 words (<= 4 phonemes each), "
                                    "T=400 frames, supplied durations, text->mel (dtts_ps_text_encode / dtts_ps_attend / "
                                    "dtts_decode_mel) -> HiFi-GAN",
                           ms_per_step=ps_ms, value=pframes / (ps_ms / 1e3), unit="frames/s", acoustic_ms=ac_ms,
                           phonemes=int((pb["txt_tokens"] > 0).sum()), frames=pframes,
                           acoustic_launches=int(peng.launches - l0) // (2 * nps))
            if not args.no_cpu_baseline:
                from oracle import ps_oracle as PO
                Wp = fold_weight_norm(p_sd)
                Wv_cpu = fold_weight_norm(synth.make_vocoder_state_dict(4321))
                from oracle import dtts_oracle as O
                torch.set_num_threads(cores)
                nu = 6

                def ps_cpu():
                    with torch.no_grad():
                        for b in range(nu):
                            n = int((pb["txt_tokens"][b] > 0).sum())
                            T = int(pb["mel_lengths"][b])
                            T4 = (T + 3) // 4
                            r = PO.ps_forward(Wp, pcfg, pb["txt_tokens"][b:b + 1, :n], pb["ph2word"][b:b + 1, :n],
                                              pb["word_lengths"][b], pb["mel2word"][b:b + 1, :T], pb["z_p"][b:b + 1, :, :T4])
                            O.hifigan_forward(Wv_cpu, vcfg, r["mel_out"][:, :T])
                ps_cpu()
                t0 = time.perf_counter()
                ps_cpu()
                dt = time.perf_counter() - t0
                fr = int(pb["mel_lengths"][:nu].sum())
                ps_line["cpu_baseline"] = dict(value=fr / dt, unit="frames/s", cores=cores, kind="port", seconds=dt,
                                               sample=f"first {nu} utterances ({fr} frames), one at a time, oracle port")
            line.setdefault("configs", {})["portaspeech"] = ps_line
            peng.close()
        except Exception as e:
            line.setdefault("configs", {})["portaspeech"] = dict(unavailable=repr(e)[:300])

    if not args.no_cpu_baseline and world == 1:
        refc = CpuReference()
        utts = [slice_utt(batch, b) for b in range(CPU_SAMPLE_UTTS)]
        fr = sum(int(u["mel_lengths"][0]) for u in utts)
        secs = refc.time_passes(utts, cores, 2, warm_utts=1)
        line["cpu_baseline"] = dict(value=fr / min(secs), unit="frames/s", cores=cores, kind=refc.kind,
                                    sample=f"first {CPU_SAMPLE_UTTS} utterances ({fr} valid frames) of the same batch, one "
                                           f"utterance at a time and un-padded as the reference's inference loop runs them, "
                                           f"{cores} host threads, best of 2 passes", seconds=min(secs),
                                    seconds_median=statistics.median(secs))
    else:
        line["cpu_baseline"] = None
    if not args.no_eager_baseline and world == 1:
        try:
            pipe.close()
            del pipe, devb
            torch.cuda.empty_cache()
            eager = {}
            for mode, what in (("fp32", "fp32, TF32 off (the reference's amp: false)"),
                               ("tf32", "TF32 tensor cores allowed in cuDNN / cuBLAS"),
                               ("fp16", "vocoder (98.6 % of the FLOPs) under torch.autocast(float16): fp16 tensor cores; acoustic "
                                        "model with TF32 (its -1e9 mask fills overflow in fp16)")):
                try:
                    fps, ms = eager_gpu_run(batch, dev, mode)
                    eager[mode] = dict(value=fps, unit="frames/s", ms_per_step=ms, what=what)
                except Exception as e:               # noqa: BLE001
                    eager[mode] = dict(unavailable=repr(e)[:200], what=what)
            line["eager_gpu_baseline"] = dict(kind="port", **{k: v for k, v in eager["fp32"].items() if k != "what"}, modes=eager,
                                              note="oracle restatement as PyTorch eager on the same B200 (cuDNN / cuBLAS), whole "
                                                   "batch, device-resident inputs; the tf32 / fp16 modes are the like-for-like "
                                                   "bars for an arm that uses 16-bit tensor-core operands")
        except Exception as e:                       # a baseline must never cost the bench line
            line["eager_gpu_baseline"] = dict(unavailable=repr(e)[:200])
    emit(json.dumps(line))                   # written before NCCL teardown: a buffered line can be lost at process exit
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
