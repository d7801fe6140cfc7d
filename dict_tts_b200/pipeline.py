"""Text -> mel -> waveform pipeline: the call a user of this repo makes.

``TextToWav.synthesize`` takes one collated HOST batch (the dict DictTTSDataset.collater builds,
tasks/tts/dataset_utils.py:264-302), copies it to the device, runs the acoustic model and the vocoder back to
back on the GPU (the mel never leaves HBM -- the reference round-trips it through numpy,
tasks/tts/dict_tts.py:231,255 and vocoders/hifigan.py:57-61) and returns the waveforms on the host.
"""
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
            # valid frames per utterance: the vocoder skips what only the padded tail depends on (dtts_vocode_lens);
            # HifiGanEngine.forward opens the 'hifigan' range itself
            wav = self.vocoder(mel, (m2w > 0).sum(-1) if self.trim_padding else None)
            if record:
                record("vocode")
        t.update(mel2word=m2w, decoder_inp=dec_in, x_mask=x_mask, mel_out=mel, z_p=z_p)
        return t, wav

    def synthesize(self, batch: Dict[str, torch.Tensor], wav_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host batch -> host waveforms [B, T*hop] (float32).  Synchronises once, at the end."""
        dev = self.to_device(batch)
        _, wav = self.run_device(dev)
        if wav_out is None:
            wav_out = torch.empty(wav.shape, dtype=wav.dtype, pin_memory=True)
        wav_out.copy_(wav, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return wav_out

    def synthesize_stream(self, batches, wav_bufs=None):
        """Throughput API: iterate over collated HOST batches, yield host waveforms [B, T*hop] in order.

        The host->device copy of batch i+1 runs on a second CUDA stream while batch i is computed, and the waveform of
        batch i is read back while batch i+1 starts (double-buffered device inputs and pinned output buffers), so a
        step costs max(copy, compute) instead of their sum.  ``wav_bufs``: optional list of two pinned tensors to
        write into (they are reused alternately; consume a result before requesting the one after next)."""
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

    def close(self):
        self.acoustic.close()
        self.vocoder.close()
