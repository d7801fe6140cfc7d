"""Text -> mel -> waveform pipeline: the call a user of this repo makes.

``TextToWav.synthesize`` takes one collated HOST batch (the dict DictTTSDataset.collater builds,
tasks/tts/dataset_utils.py:264-302), copies it to the device, runs the acoustic model and the vocoder back to
back on the GPU (the mel never leaves HBM -- the reference round-trips it through numpy,
tasks/tts/dict_tts.py:231,255 and vocoders/hifigan.py:57-61) and returns the waveforms on the host.
"""
import os
from typing import Dict, Optional

import torch

from .config import AcousticConfig, VocoderConfig
from .engine import DictTTSEngine, HifiGanEngine
from .profiling import Timer

_INPUT_KEYS = ("word_tokens", "pron_modified", "keys", "values", "key_map", "pinyin", "pinyin_map", "mel2word", "z_p",
               "dict_ids")


class TextToWav:
    def __init__(self, acoustic_sd, vocoder_sd, acfg: Optional[AcousticConfig] = None,
                 vcfg: Optional[VocoderConfig] = None, device="cuda:0", arenas=None, vocoder_precision: int = 6,
                 acoustic_precision: int = 1, s2pa_route: int = 0, trim_padding: bool = True):
        """trim_padding: vocode only up to each utterance's valid length (the waveform past it is 0 instead of the
        vocoded padding frames; valid samples are bit-identical either way)."""
        a_arena = a_table = v_arena = v_table = None
        if arenas is not None:
            (a_arena, a_table), (v_arena, v_table) = arenas
        self.device = torch.device(device)
        self.acoustic = DictTTSEngine(acoustic_sd, acfg, device, a_arena, a_table, precision=acoustic_precision,
                                      s2pa_route=s2pa_route)
        self.vocoder = HifiGanEngine(vocoder_sd, vcfg, device, v_arena, v_table, precision=vocoder_precision)
        self.events = None
        self.trim_padding = trim_padding

    @property
    def launches(self) -> int:
        return self.acoustic.launches + self.vocoder.launches

    def to_device(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for k in _INPUT_KEYS:
            v = batch.get(k)
            if k == "values" and v is not None and v is batch.get("keys"):
                continue                          # one tensor for both (reference data: value IS key): copied once
            if v is not None:
                out[k] = v.to(self.device, non_blocking=True)
        if out.get("values") is None and "keys" in out:
            out["values"] = out["keys"]
        if "keys" not in out:                     # bank batch: the (Lk, Lp) the collater would pad to, from host offsets
            local = batch.get("dict_bank")        # ragged batch: its own un-padded bank travels with it
            if local is not None:
                out["_dims"] = local.batch_dims(batch["dict_ids"])
                out["_bank"] = local.to(self.device, non_blocking=True)
            else:
                out["_dims"] = self.acoustic.bank.batch_dims(batch["dict_ids"])
        return out

    def run_device(self, dev: Dict[str, torch.Tensor], record=None):
        """Device-resident inputs -> (ret dict, wav [B, T*hop]) on the device.  ``record(name)`` marks stage ends."""
        t, lens = self.run_acoustic(dev, record)
        with torch.cuda.device(self.device):
            # valid frames per utterance: the vocoder skips what only the padded tail depends on (dtts_vocode_lens);
            # HifiGanEngine.forward opens the 'hifigan' range itself
            wav = self.vocoder(t["mel_out"], lens)
            if record:
                record("vocode")
        return t, wav

    def run_acoustic(self, dev: Dict[str, torch.Tensor], record=None):
        """Text -> mel on the current stream: (ret dict incl. ``mel_out``, valid frames per utterance or None)."""
        eng = self.acoustic
        prof = eng.profile_infer
        with torch.cuda.device(self.device):
            with Timer("encoder", enable=prof):       # the reference's stage names (profiling.py)
                with Timer("dict_encoder", enable=prof):
                    if "keys" in dev:
                        t = eng.text_encode(dev["word_tokens"], dev.get("pron_modified"), dev["keys"], dev["values"],
                                            dev["key_map"], dev["pinyin"], dev["pinyin_map"])
                    else:                             # GPU-resident dictionary bank: only ids cross the bus
                        if "_bank" in dev:
                            eng.set_dict_bank(dev["_bank"])
                        t = eng.text_encode_bank(dev["word_tokens"], dev.get("pron_modified"), dev["dict_ids"],
                                                 *dev["_dims"])
                if record:
                    record("text_encode")
                m2w = dev.get("mel2word")
                if m2w is None:
                    m2w = eng.length_regulate(t["dur_int"], t["ilens"])
                elif m2w.shape[1] % eng.cfg.frames_multiple:
                    pad = eng.cfg.frames_multiple - m2w.shape[1] % eng.cfg.frames_multiple
                    m2w = torch.cat([m2w] + [m2w[:, -1:]] * pad, -1).contiguous()
                dec_in, g_bct, x_mask = eng.expand(t["word_encoder_out"], m2w)
                if record:
                    record("length_regulate")
            z = dev.get("z_p")
            if z is None:
                z = torch.distributions.Normal(0, 1).sample([g_bct.shape[0], eng.cfg.latent,
                                                             g_bct.shape[2] // eng.cfg.frames_multiple])
            with Timer("fvae", enable=prof):
                mel, z_p = eng.decode_mel(g_bct, z)
            if record:
                record("decode_mel")
            lens = (m2w > 0).sum(-1) if self.trim_padding else None
        t.update(mel2word=m2w, decoder_inp=dec_in, x_mask=x_mask, mel_out=mel, z_p=z_p)
        return t, lens

    def synthesize(self, batch: Dict[str, torch.Tensor], wav_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host batch -> host waveforms [B, T*hop] (float32).  Synchronises once, at the end."""
        dev = self.to_device(batch)
        _, wav = self.run_device(dev)
        if wav_out is None:
            wav_out = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
        wav_out.copy_(wav, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return wav_out

    def synthesize_stream(self, batches, wav_bufs=None, overlap_acoustic: Optional[bool] = None):
        """Throughput API: iterate over collated HOST batches, yield host waveforms [B, T*hop] in order.

        The host->device copy of batch i+1 runs on a second CUDA stream while batch i is computed, and the waveform of
        batch i is read back while batch i+1 starts (double-buffered device inputs and pinned output buffers), so a
        step costs max(copy, compute) instead of their sum.  ``wav_bufs``: optional list of two pinned tensors to
        write into (they are reused alternately; consume a result before requesting the one after next).

        ``overlap_acoustic`` (default: the ``DTTS_OVERLAP_ACOUSTIC`` environment switch, OFF): the acoustic model of batch
        i+1 -- a latency-bound chain of ~100 short launches on <= 60 SMs -- runs on its own high-priority stream while
        the vocoder of batch i (throughput-bound, every SM) runs on the compute stream, instead of behind it.  Measured
        on a B200 (cfg 2, two runs each): 26.7 / 26.8 ms per step with the overlap against 26.0 / 26.1 without -- the
        vocoder's persistent kernels own every SM (227 KB of shared memory, all of TMEM), so the acoustic CTAs only run
        in the slots they take away from the next vocoder kernel and delay its statically scheduled tiles.  Kept as an
        opt-in for smaller vocoder batches."""
        if overlap_acoustic is None:
            overlap_acoustic = os.environ.get("DTTS_OVERLAP_ACOUSTIC", "0") not in ("", "0")
        if overlap_acoustic:
            yield from self._synthesize_stream_overlapped(batches, wav_bufs)
            return
        dev = self.device
        compute = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
        readback = getattr(self, "_readback_stream", None)
        if readback is None:
            readback = self._readback_stream = torch.cuda.Stream(dev)
        it = iter(batches)
        slots = [None, None]                 # (device batch, copied event)
        free = [None, None]                  # event: compute has finished reading slot i

        def upload(i, batch):
            with torch.cuda.stream(copy):
                if free[i] is not None:
                    copy.wait_event(free[i])
                d = self.to_device(batch)
                for t in list(d.values()) + (d["_bank"].tensors() if "_bank" in d else []):
                    if torch.is_tensor(t):
                        t.record_stream(compute)      # allocated on the copy stream, consumed on the compute stream
                ev = torch.cuda.Event()
                ev.record(copy)
            slots[i] = (d, ev)

        try:
            first = next(it)
        except StopIteration:
            return
        upload(0, first)
        pending = None                       # (wav_host, done event) of the previous batch
        k = 0
        while slots[k % 2] is not None:
            cur = k % 2
            d, ev = slots[cur]
            slots[cur] = None
            nxt = next(it, None)
            if nxt is not None:
                upload(1 - cur, nxt)         # overlaps with the compute below
            compute.wait_event(ev)
            _, wav = self.run_device(d)
            done_reading = torch.cuda.Event()
            done_reading.record(compute)
            free[cur] = done_reading
            if wav_bufs is not None:
                out = wav_bufs[cur][:wav.shape[0], :wav.shape[1]]
            else:
                out = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
            # the read-back runs on its own stream, so the next batch's kernels do not queue behind it
            readback.wait_event(done_reading)
            with torch.cuda.stream(readback):
                out.copy_(wav, non_blocking=True)
                wav.record_stream(readback)
                done = torch.cuda.Event()
                done.record(readback)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (out, done)
            k += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]

    def _synthesize_stream_overlapped(self, batches, wav_bufs=None):
        """synthesize_stream with the acoustic model of batch i+1 overlapping the vocoder of batch i.

        Streams: copy (H2D of batch i+1), acoustic (text -> mel, high priority so that its short CTAs take the SMs a
        vocoder kernel frees in its tail), compute = the caller's stream (vocoder), readback (D2H of the waveform).
        The acoustic workspace is only touched on the acoustic stream, the vocoder workspace only on the compute
        stream; tensors that cross streams are handed over with events + record_stream."""
        dev = self.device
        compute = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(dev)
        readback = getattr(self, "_readback_stream", None)
        if readback is None:
            readback = self._readback_stream = torch.cuda.Stream(dev)
        ac = getattr(self, "_acoustic_stream", None)
        if ac is None:
            ac = self._acoustic_stream = torch.cuda.Stream(dev, priority=-1)
        it = iter(batches)
        free = [None, None]                  # event: the acoustic stream has finished reading input slot i

        def upload(i, batch):
            with torch.cuda.stream(copy):
                if free[i] is not None:
                    copy.wait_event(free[i])
                d = self.to_device(batch)
                for t in list(d.values()) + (d["_bank"].tensors() if "_bank" in d else []):
                    if torch.is_tensor(t):
                        t.record_stream(ac)
                ev = torch.cuda.Event()
                ev.record(copy)
            return d, ev

        def acoustic(i, d, ev):
            """text -> mel of one batch on the acoustic stream: (mel, lens, event 'mel is ready')."""
            with torch.cuda.stream(ac):
                ac.wait_event(ev)
                t, lens = self.run_acoustic(d)
                mel = t["mel_out"]
                mel.record_stream(compute)
                if lens is not None:
                    lens.record_stream(compute)
                done = torch.cuda.Event()
                done.record(ac)
            free[i] = done
            return mel, lens, done

        try:
            first = next(it)
        except StopIteration:
            return
        ac.wait_stream(compute)              # whatever the caller enqueued before (weights, bank) is visible
        stage = acoustic(0, *upload(0, first))
        pending = None                       # (wav_host, done event) of the previous batch
        k = 0
        while stage is not None:
            mel, lens, mel_ready = stage
            nxt = next(it, None)
            # enqueue the NEXT batch's acoustic model before this batch's vocoder: both are then in flight together
            stage = acoustic((k + 1) % 2, *upload((k + 1) % 2, nxt)) if nxt is not None else None
            compute.wait_event(mel_ready)
            with torch.cuda.device(dev):
                wav = self.vocoder(mel, lens)
            voc_done = torch.cuda.Event()
            voc_done.record(compute)
            if wav_bufs is not None:
                out = wav_bufs[k % 2][:wav.shape[0], :wav.shape[1]]
            else:
                out = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
            readback.wait_event(voc_done)
            with torch.cuda.stream(readback):
                out.copy_(wav, non_blocking=True)
                wav.record_stream(readback)
                done = torch.cuda.Event()
                done.record(readback)
            if pending is not None:
                pending[1].synchronize()
                yield pending[0]
            pending = (out, done)
            k += 1
        if pending is not None:
            pending[1].synchronize()
            yield pending[0]
        compute.wait_stream(ac)

    def close(self):
        self.acoustic.close()
        self.vocoder.close()
