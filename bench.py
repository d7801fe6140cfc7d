"""Benchmark of the Dict-TTS text -> mel -> waveform forward pass (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one Biaobei-shaped synthetic batch (cfg 2 of BASELINE.json: batch 60,
<= 22 word tokens, L_k <= 96 gloss tokens, 400 mel frames, fp32, full text -> mel -> HiFi-GAN).  With N > 1 the
driver launches this file under torchrun: every rank owns one GPU and one independent batch (weak scaling,
SURVEY.md §8e), rank 0 broadcasts the weight arenas over NCCL once at load and there is no hot-path collective.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
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

# algorithmic work model (checked against torch FlopCounterMode on the reference)
VOCODER_FLOP_PER_FRAME = 614.1e6
TC_CONV_DRAM_BYTES_PER_LAUNCH = 1.076e9   # measured (ncu): 82.9 GB over the 77 launches of a valid-length pass (95.5 GB full length)
WORKLOAD = dict(B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=96)
CPU_SAMPLE_UTTS = 30      # ~10 k of the batch's 20.7 k frames: 10-20 s of host work for the two timed passes


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


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


def cpu_port_run(batch, n_utts, threads, iters):
    """The oracle port (own-words restatement of the reference's PyTorch CPU forward, oracle/dtts_oracle.py) on the
    host cores, on the first n_utts utterances of the same batch.  Returns (frames/s, seconds per pass, frames)."""
    from oracle import dtts_oracle as O
    torch.set_num_threads(threads)
    cfg, vcfg = AcousticConfig(), VocoderConfig()
    W = fold_weight_norm(synth.make_acoustic_state_dict(1234))
    Wv = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    sub = {k: (v[:n_utts].contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}
    frames = int(sub["mel_lengths"].sum())

    def once():
        with torch.no_grad():
            ret = O.acoustic_forward(W, cfg, sub, sub["mel2word"], sub["z_p"])
            wavs = [O.hifigan_forward(Wv, vcfg, ret["mel_out"][b:b + 1]) for b in range(n_utts)]  # spec2wav is B=1
        return wavs
    best = None
    for _ in range(iters):
        t0 = time.perf_counter()
        once()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return frames / best, best, frames


def eager_gpu_run(batch, dev, iters=2):
    """The same restatement of the reference's PyTorch forward (oracle/dtts_oracle.py), run as PyTorch eager ON THE GPU:
    the 'library kernels' bar of SURVEY.md §8d (cuDNN convolutions, cuBLAS GEMMs, element-wise kernels, fp32, TF32 off),
    whole cfg-2 batch, vocoder batched (kinder than the reference's one utterance at a time).  Returns (frames/s, ms)."""
    from oracle import dtts_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, vcfg = AcousticConfig(), VocoderConfig()
    W = {k: v.to(dev) for k, v in fold_weight_norm(synth.make_acoustic_state_dict(1234)).items()}
    Wv = {k: v.to(dev) for k, v in fold_weight_norm(synth.make_vocoder_state_dict(4321)).items()}
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    frames = int(batch["mel_lengths"].sum())

    def once():
        with torch.no_grad():
            ret = O.acoustic_forward(W, cfg, b, b["mel2word"], b["z_p"])
            return O.hifigan_forward(Wv, vcfg, ret["mel_out"])
    once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        once()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true",
                    help="skip timing the PyTorch-eager (cuDNN/cuBLAS) forward of the same batch on the GPU")
    ap.add_argument("--vocoder-precision", type=int, default=int(os.environ.get("DTTS_VOCODER_PRECISION", "6")),
                    help="0: fp32 FMA pipe; 1: tcgen05 bf16 hi/lo x hi/lo (3 MMAs, ~1e-6 wav RMS); 2: tcgen05 bf16 (cfg 3, "
                         "outside the tolerance); 3: tcgen05 fp16 x fp16 hi/lo weights (2 MMAs, ~6e-5 wav RMS, "
                         "inside the 1e-4 tolerance); 4: tcgen05 fp16 (1 MMA, ~8.5e-5); 5: as 3 with single-plane weights "
                         "in the C_out >= 128 layers (~7.5e-5); 6 (default): as 3 with the lo plane of the C_out >= 128, k >= 7 layers as an FP8 MMA "
                         "(same error as 3)")
    ap.add_argument("--acoustic-precision", type=int, default=int(os.environ.get("DTTS_ACOUSTIC_PRECISION", "1")),
                    help="0: fp32 FMA pipe; 1 (default): dense convolutions on tcgen05, bf16 hi/lo split (fp32-class)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    emit = claim_stdout()
    config = dict(workload="cfg2: Biaobei-shaped batch=60/GPU, Tw<=22, Lk<=96, T=400 frames, fp32, text->mel->HiFi-GAN",
                  batch_per_gpu=WORKLOAD["B"], frames_per_utt=WORKLOAD["max_frames"], parallelism=f"replicas x{world}",
                  supplied_durations=True, l2="inputs (779 MB keys+values per step) exceed the 126 MB L2")

    if args.impl == "reference":
        if rank != 0:
            return
        batch = synth.make_batch(seed=1234, **WORKLOAD)
        steps = max(1, min(args.steps, 3))
        if args.warmup > 0:
            cpu_port_run(batch, 1, cores, 1)
        fps, secs, frames = cpu_port_run(batch, CPU_SAMPLE_UTTS, cores, steps)
        audio = frames * HOP_SIZE / SAMPLE_RATE
        sample = f"first {CPU_SAMPLE_UTTS} utterances ({frames} frames) of the cfg-2 batch, best of {steps}"
        emit(json.dumps(dict(
            impl="reference", metric="mel_frames_per_s", value=fps, unit="frames/s", n_gpus=args.gpus, steps=steps,
            warmup=min(args.warmup, 1), ms_per_step=secs * 1e3, higher_is_better=True, scaling="weak",
            vs_baseline=None, dtype="f32", data="synthetic", config=config, rtf=secs / audio,
            cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, kind="port", sample=sample),
            e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from dict_tts_b200.pipeline import TextToWav

    # ---- weights: rank 0 builds the arenas, one NCCL broadcast each ----
    acfg, vcfg = AcousticConfig(), VocoderConfig()
    a_sd = drop_dead(fold_weight_norm(synth.make_acoustic_state_dict(1234)))
    v_sd = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    a_host, a_table = pack_arena(a_sd)
    v_host, v_table = pack_arena(v_sd)
    if world > 1:
        a_arena = a_host.to(dev) if rank == 0 else torch.empty_like(a_host, device=dev)
        v_arena = v_host.to(dev) if rank == 0 else torch.empty_like(v_host, device=dev)
        dist.broadcast(a_arena, 0)
        dist.broadcast(v_arena, 0)
    else:
        a_arena, v_arena = a_host.to(dev), v_host.to(dev)
    pipe = TextToWav(None, None, acfg, vcfg, dev, arenas=((a_arena, a_table), (v_arena, v_table)),
                     vocoder_precision=args.vocoder_precision, acoustic_precision=args.acoustic_precision)

    batch = synth.make_batch(seed=1234 + rank, **WORKLOAD)
    frames = int(batch["mel_lengths"].sum())
    audio_s = frames * HOP_SIZE / SAMPLE_RATE
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in batch.items()}
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

    # ---- device-resident arm ----
    for _ in range(max(args.warmup, 3)):
        pipe.run_device(devb)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    marks = []
    launches0 = pipe.launches
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_marks = [torch.cuda.Event(enable_timing=True)]
        step_marks[0].record(stream)

        def record(name, sm=step_marks):
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            sm.append(e)
        pipe.run_device(devb, record)
        marks.append(step_marks)
    ev1.record(stream)
    barrier()
    launches = pipe.launches - launches0
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    stage_ms = {n: sum(m[i].elapsed_time(m[i + 1]) for m in marks) / args.steps for i, n in enumerate(stage_names)}

    # ---- end-to-end arm: pinned host inputs -> H2D -> engine -> wav D2H, every step, through the public API.
    # synthesize_stream() is the throughput call: the upload of step i+1 overlaps the compute of step i (second stream),
    # every step still copies its 780 MB of inputs from pinned host memory and reads its waveform back.
    wav_bufs = [torch.empty(WORKLOAD["B"], WORKLOAD["max_frames"] * HOP_SIZE, dtype=torch.float32, pin_memory=True)
                for _ in range(2)]
    wav_host = wav_bufs[0]

    def host_batches(n):
        for _ in range(n):
            yield host
    for _ in pipe.synthesize_stream(host_batches(2), wav_bufs):
        pass
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n_out = 0
    for w in pipe.synthesize_stream(host_batches(args.steps), wav_bufs):
        n_out += 1
    e1.record(stream)
    barrier()
    assert n_out == args.steps
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    # latency call (one batch, nothing overlapped): H2D + compute + D2H back to back
    for _ in range(2):
        pipe.synthesize(host, wav_host)
    barrier()
    l0 = torch.cuda.Event(enable_timing=True)
    l1 = torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    for _ in range(min(args.steps, 5)):
        pipe.synthesize(host, wav_host)
    l1.record(stream)
    barrier()
    e2e_serial_ms = max_over_ranks(l0.elapsed_time(l1)) / min(args.steps, 5)
    # ---- same end-to-end call with the GPU-resident dictionary bank: the characters are named by
    # id, keys/values never cross the bus again (the bank -- here one entry per character position of this batch -- is
    # uploaded once, outside the timed region, like the weights)
    from dict_tts_b200.bank import DictBank
    bank, dict_ids = DictBank.from_batch(batch)
    pipe.acoustic.set_dict_bank(bank)
    slim = {k: v for k, v in host.items() if k not in ("keys", "values", "key_map", "pinyin", "pinyin_map")}
    slim["dict_ids"] = dict_ids.pin_memory()

    def slim_batches(n):
        for _ in range(n):
            yield slim
    for _ in pipe.synthesize_stream(slim_batches(2), wav_bufs):
        pass
    barrier()
    b0 = torch.cuda.Event(enable_timing=True)
    b1 = torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for w in pipe.synthesize_stream(slim_batches(args.steps), wav_bufs):
        pass
    b1.record(stream)
    barrier()
    bank_ms = max_over_ranks(b0.elapsed_time(b1)) / args.steps
    bank_h2d = sum(slim[k].numel() * slim[k].element_size() for k in
                   ("word_tokens", "pron_modified", "dict_ids", "mel2word", "z_p"))
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(host[k].numel() * host[k].element_size() for k in
              ("word_tokens", "pron_modified", "keys", "values", "key_map", "pinyin", "pinyin_map", "mel2word", "z_p"))
    d2h = wav_host.numel() * wav_host.element_size()

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
    dtype = {0: "f32", 1: "f32 (vocoder convs: bf16x3 split on tcgen05, fp32 accumulate)",
             2: "bf16 vocoder convs (tcgen05), f32 elsewhere",
             3: "f32 (vocoder convs: fp16 activations x fp16 hi/lo weights on tcgen05, fp32 accumulate)",
             4: "fp16 vocoder convs (tcgen05), f32 elsewhere",
             5: "f32 (vocoder convs: fp16 x fp16 on tcgen05, hi/lo weight planes where C_out < 128, fp32 accumulate)",
             6: "f32 (vocoder convs: fp16 activations x fp16 hi/lo weights on tcgen05, lo-plane correction of the "
                "C_out >= 128 layers as an e5m2 MMA, fp32 accumulate)"}[args.vocoder_precision]
    config["vocoder_precision"] = args.vocoder_precision
    config["acoustic_precision"] = args.acoustic_precision
    voc_s = stage_ms["vocode"] / 1e3
    # the vocoder is run up to each utterance's valid length (dtts_vocode_lens): algorithmic work = valid frames only
    # (the ~1 frame of receptive-field margin it also computes per utterance is not counted)
    padded_frames = frames
    config["vocoder_frames"] = "valid frames + receptive-field margin (padded tail of the batch skipped)"
    achieved = padded_frames * VOCODER_FLOP_PER_FRAME / voc_s / 1e12
    n_voc_launch = 1 + 4 + 72 + (1 if args.vocoder_precision == 0 else 0)   # conv_post is a separate CUDA-core kernel on the TC path
    kname = {0: "conv1d_f32_kernel (fp32 FMA pipe)", 1: "tc_conv_kernel (tcgen05, bf16 hi/lo x hi/lo: 3 MMAs per product)",
             2: "tc_conv_kernel (tcgen05, bf16: 1 MMA)", 3: "tc_conv_kernel (tcgen05, fp16 x fp16 hi/lo weights: 2 MMAs)",
             4: "tc_conv_kernel (tcgen05, fp16: 1 MMA)",
             5: "tc_conv_kernel (tcgen05, fp16; hi/lo weights only where C_out < 128)",
             6: "tc_conv_kernel (tcgen05, fp16 x fp16 hi/lo weights: 2 MMAs; lo plane in FP8 where C_out >= 128: 1.5)"
             }[args.vocoder_precision]
    roofline = dict(bound="tensor", kernel="%s, HiFi-GAN stack, %d launches/step" % (kname, n_voc_launch),
                    achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"],
                    traffic=(TC_CONV_DRAM_BYTES_PER_LAUNCH if args.vocoder_precision in (3, 6) else None),   # same bytes in 3 and 6
                    traffic_source="profiles/r01_vocoder_lens_dram_agg.txt (ncu dram__bytes_read+write, average over the "
                                   "77 tc_conv_kernel launches of one valid-length vocode pass; algorithmic: 1.08 GB)",
                    peak_source=peaks["source"] + " bf16 dense (sustained)",
                    avg_launch_ms=stage_ms["vocode"] / n_voc_launch,
                    flop_per_launch=padded_frames * VOCODER_FLOP_PER_FRAME / n_voc_launch)
    line = dict(metric="mel_frames_per_s", value=total_frames / (dev_ms / 1e3), unit="frames/s", n_gpus=world,
                steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=dev_ms, higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype=dtype, data="synthetic", config=config,
                rtf=(dev_ms / 1e3) / (total_frames * HOP_SIZE / SAMPLE_RATE),
                x_realtime=(total_frames * HOP_SIZE / SAMPLE_RATE) / (dev_ms / 1e3),
                stages_ms=stage_ms, gpu_launches=int(launches), clocks=clocks, roofline=roofline,
                e2e=dict(value=total_frames / (e2e_ms / 1e3), unit="frames/s", h2d_bytes_per_step=int(h2d),
                         d2h_bytes_per_step=int(d2h), ms_per_step=e2e_ms,
                         x_realtime=(total_frames * HOP_SIZE / SAMPLE_RATE) / (e2e_ms / 1e3),
                         api="TextToWav.synthesize_stream (copy of step i+1 overlaps compute of step i)",
                         serial_ms_per_step=e2e_serial_ms),
                e2e_dict_bank=dict(value=total_frames / (bank_ms / 1e3), unit="frames/s", ms_per_step=bank_ms,
                                   h2d_bytes_per_step=int(bank_h2d), d2h_bytes_per_step=int(d2h),
                                   bank_bytes_resident=int(bank.nbytes),
                                   note="same call with characters named by dictionary-bank id (SURVEY.md 8f-1)"))
    if not args.no_cpu_baseline and world == 1:
        fps, secs, fr = cpu_port_run(batch, CPU_SAMPLE_UTTS, cores, 2)
        line["cpu_baseline"] = dict(value=fps, unit="frames/s", cores=cores, kind="port",
                                    sample=f"first {CPU_SAMPLE_UTTS} utterances ({fr} frames) of the same batch, "
                                           f"oracle port on {cores} host threads, best of 2", seconds=secs)
    else:
        line["cpu_baseline"] = None
    if not args.no_eager_baseline and world == 1:
        try:
            pipe.close()
            del pipe, devb
            torch.cuda.empty_cache()
            fps, ms = eager_gpu_run(batch, dev)
            line["eager_gpu_baseline"] = dict(value=fps, unit="frames/s", ms_per_step=ms, kind="port",
                                              note="oracle restatement as PyTorch eager on the same B200 (cuDNN / cuBLAS, "
                                                   "fp32, TF32 off), whole batch, device-resident inputs")
        except Exception as e:                       # a baseline must never cost the bench line
            line["eager_gpu_baseline"] = dict(unavailable=repr(e)[:200])
    emit(json.dumps(line))                   # written before NCCL teardown: a buffered line can be lost at process exit
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
