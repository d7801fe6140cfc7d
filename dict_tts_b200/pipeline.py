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

_INPUT_KEYS = ("word_tokens", "pron_modified", "keys", "values", "key_map", "pinyin", "pinyin_map", "mel2word", "z_p")


class TextToWav:
    def __init__(self, acoustic_sd, vocoder_sd, acfg: Optional[AcousticConfig] = None,
                 vcfg: Optional[VocoderConfig] = None, device="cuda:0", arenas=None, vocoder_precision: int = 3,
                 acoustic_precision: int = 1):
        a_arena = a_table = v_arena = v_table = None
        if arenas is not None:
            (a_arena, a_table), (v_arena, v_table) = arenas
        self.device = torch.device(device)
        self.acoustic = DictTTSEngine(acoustic_sd, acfg, device, a_arena, a_table, precision=acoustic_precision)
        self.vocoder = HifiGanEngine(vocoder_sd, vcfg, device, v_arena, v_table, precision=vocoder_precision)
        self.events = None

    @property
    def launches(self) -> int:
        return self.acoustic.launches + self.vocoder.launches

    def to_device(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for k in _INPUT_KEYS:
            v = batch.get(k)
            if v is not None:
                out[k] = v.to(self.device, non_blocking=True)
        if out.get("values") is None:
            out["values"] = out["keys"]
        return out

    def run_device(self, dev: Dict[str, torch.Tensor], record=None):
        """Device-resident inputs -> (ret dict, wav [B, T*hop]) on the device.  ``record(name)`` marks stage ends."""
        eng = self.acoustic
        with torch.cuda.device(self.device):
            t = eng.text_encode(dev["word_tokens"], dev.get("pron_modified"), dev["keys"], dev["values"],
                                dev["key_map"], dev["pinyin"], dev["pinyin_map"])
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
            mel, z_p = eng.decode_mel(g_bct, z)
            if record:
                record("decode_mel")
            wav = self.vocoder(mel)
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

    def close(self):
        self.acoustic.close()
        self.vocoder.close()
