"""Host-side mirror of the reference's model interface on top of libdtts (tensor ownership + glue only).

``DictTTSEngine.forward`` keeps the call signature and the returned dict of
``PortaSpeech_dict.forward`` (modules/dict_tts/model.py:36-62) so ``DictTTSTask.test_step``
(tasks/tts/dict_tts.py:179-196) can call it unchanged; ``HifiGanEngine`` is the generator behind
``BaseVocoder.spec2wav`` (vocoders/hifigan.py:54-62).  All FLOPs run in the CUDA library; PyTorch only
allocates buffers and hands out pointers.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import binding
from .config import AcousticConfig, PortaSpeechConfig, VocoderConfig
from .profiling import Timer
from .weights import drop_dead, fold_weight_norm, pack_arena


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, device):
    return t.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()


def _dev_i64(t, device):
    return t.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


class _Workspace:
    """Grow-only device scratch buffer owned by PyTorch."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            if self.buf is not None:
                # the old block goes back to the allocator of the stream it was created on while kernels of ANOTHER stream
                # (synthesize_stream runs the acoustic model on its own stream) may still use it: growth is rare, wait
                torch.cuda.synchronize(self.device)
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.buf


class DictTTSEngine:
    """Acoustic model (text -> mel).  ``state_dict`` uses the reference checkpoint keys (weight-norm pairs allowed)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Optional[AcousticConfig] = None, device="cuda:0",
                 arena: Optional[torch.Tensor] = None, table=None, precision: int = 1, s2pa_route: int = 0):
        """precision 0: every convolution on the fp32 FMA pipe (exact); 1 (default): dense convolutions on tcgen05 with
        bf16 hi/lo split operands (3 MMAs per product, fp32-class accuracy).
        s2pa_route 0 (default): folded streaming S2PA; 1: K/V projection of every gloss token as one tcgen05 GEMM, as
        the reference computes it (layers/dict_encoder.py:40-58) -- see dtts_acoustic_desc.s2pa_route."""
        if not torch.cuda.is_available():
            raise RuntimeError("DictTTSEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = binding.load()
        self.cfg = cfg or AcousticConfig()
        self.device = torch.device(device)
        if arena is None:
            host, table = pack_arena(drop_dead(fold_weight_norm(state_dict)))
            arena = host.to(self.device)
        self.arena, self.table = arena, table
        c = self.cfg
        desc = binding.AcousticDesc(c.hidden, c.n_heads, c.enc_layers, c.ffn_kernel, c.ffn_filter, c.dict_dim,
                                    c.word_size, c.pinyin_size, c.dur_layers, c.dur_kernel, c.dur_chans,
                                    c.frames_multiple, c.latent, c.dec_layers, c.dec_kernel, c.flow_hidden,
                                    c.flow_kernel, c.flow_blocks, c.flow_layers, c.n_mel, int(c.language_zh), int(precision),
                                    int(s2pa_route), binding.MODEL_DICT, 0, 0, 0)
        tab, self._keep = binding.make_table(table)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            binding.check(self.lib.dtts_acoustic_create(C.byref(desc), _ptr(self.arena), self.arena.numel(), tab,
                                                        len(table), _stream(), C.byref(self.handle)), "acoustic_create")
        self.ws = _Workspace(self.device)
        self._t_raw = C.c_int32(0)
        self.profile_infer = False     # hparams['profile_infer']: the stage Timers synchronise and print like upstream

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dtts_acoustic_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.dtts_acoustic_launch_count(self.handle))

    # -- stages (each maps to one C-ABI call) --------------------------------------------------------------
    def text_encode(self, word_tokens, pron_modified, keys, values, key_map, pinyin, pinyin_map):
        dev = self.device
        wt = _dev_i64(word_tokens, dev)
        B, Tw = wt.shape
        same = values is None or values is keys
        keys = _dev_f32(keys, dev)
        values = keys if same else _dev_f32(values, dev)
        key_map = _dev_f32(key_map, dev)
        pinyin = _dev_i64(pinyin, dev)
        pinyin_map = _dev_i64(pinyin_map, dev)
        pm = None if pron_modified is None else _dev_i64(pron_modified, dev)
        Lk, Lp = key_map.shape[2], pinyin.shape[2]
        if tuple(keys.shape) != (B, Tw, Lk, self.cfg.dict_dim) or tuple(pinyin_map.shape) != (B, Tw, Lp):
            raise ValueError("dict_msg tensors have inconsistent shapes")
        H = self.cfg.hidden
        out = dict(word_encoder_out=torch.empty(B, Tw, H, device=dev),
                   dict_attn=torch.empty(B, 1, Lk, Tw, device=dev),
                   pron_attn=torch.empty(B, Tw, Lp, device=dev),
                   dur=torch.empty(B, Tw, device=dev),
                   dur_int=torch.empty(B, Tw, dtype=torch.int64, device=dev),
                   ilens=torch.empty(B, dtype=torch.int64, device=dev))
        tin = binding.TextIn(_ptr(wt), _ptr(pm), _ptr(keys), _ptr(values), _ptr(key_map), _ptr(pinyin),
                             _ptr(pinyin_map), B, Tw, Lk, Lp)
        tout = binding.TextOut(*[_ptr(out[k]) for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur",
                                                          "dur_int", "ilens")])
        nbytes = self.lib.dtts_text_workspace_bytes(self.handle, B, Tw, Lk, Lp)
        ws = self.ws.get(nbytes)
        binding.check(self.lib.dtts_text_encode(self.handle, C.byref(tin), C.byref(tout), _ptr(ws), ws.numel(),
                                                _stream()), "text_encode")
        return out

    # -- GPU-resident dictionary bank ------------------------------------------------------
    def set_dict_bank(self, bank):
        """Uploads (if needed) and registers a dict_tts_b200.bank.DictBank; afterwards forward(dict_ids=...) and
        text_encode_bank() name characters by id instead of shipping keys/values for every batch."""
        self.bank = bank if bank.keys.is_cuda else bank.to(self.device)
        return self.bank

    def text_encode_bank(self, word_tokens, pron_modified, dict_ids, Lk: int = 0, Lp: int = 0):
        bank = getattr(self, "bank", None)
        if bank is None:
            raise RuntimeError("no dictionary bank registered: call set_dict_bank() first")
        dev = self.device
        if Lk <= 0 or Lp <= 0:                                # host-side validation of the ids + the collater's widths
            lk, lp = bank.batch_dims(dict_ids)                #   (pass Lk, Lp to skip the device->host read of the ids)
            Lk, Lp = max(int(Lk), lk), max(int(Lp), lp)
        wt = _dev_i64(word_tokens, dev)
        ids = _dev_i64(dict_ids, dev)
        B, Tw = wt.shape
        if tuple(ids.shape) != (B, Tw):
            raise ValueError("dict_ids must be [B,Tw] like word_tokens")
        pm = None if pron_modified is None else _dev_i64(pron_modified, dev)
        H = self.cfg.hidden
        out = dict(word_encoder_out=torch.empty(B, Tw, H, device=dev),
                   dict_attn=torch.empty(B, 1, Lk, Tw, device=dev),
                   pron_attn=torch.empty(B, Tw, Lp, device=dev),
                   dur=torch.empty(B, Tw, device=dev),
                   dur_int=torch.empty(B, Tw, dtype=torch.int64, device=dev),
                   ilens=torch.empty(B, dtype=torch.int64, device=dev))
        tin = binding.TextInBank(_ptr(wt), _ptr(pm), _ptr(ids), B, Tw, Lk, Lp)
        tout = binding.TextOut(*[_ptr(out[k]) for k in ("word_encoder_out", "dict_attn", "pron_attn", "dur",
                                                          "dur_int", "ilens")])
        ws = self.ws.get(self.lib.dtts_text_bank_workspace_bytes(self.handle, B, Tw, Lk, Lp))
        st = bank.c_struct()
        binding.check(self.lib.dtts_text_encode_bank(self.handle, C.byref(st), C.byref(tin), C.byref(tout), _ptr(ws),
                                                     ws.numel(), _stream()), "text_encode_bank")
        return out

    # -- after_infer on the device ----------------------------------------------------------
    def pcm16(self, wav: torch.Tensor) -> torch.Tensor:
        """float waveform [.., n] on the device -> int16 PCM, (wav * 32767).astype(int16) of utils/audio.py:15-16."""
        wav = _dev_f32(wav, self.device)
        n = wav.numel()
        if n % 4:
            raise ValueError("sample count must be a multiple of 4 (it is a multiple of hop_size)")
        pcm = torch.empty(wav.shape, dtype=torch.int16, device=self.device)
        with torch.cuda.device(self.device):
            binding.check(self.lib.dtts_wav_to_pcm16(_ptr(wav), n, _ptr(pcm), _stream()), "wav_to_pcm16")
        return pcm

    def pron_tokens(self, pron_attn: torch.Tensor, pinyin=None, dict_ids=None) -> torch.Tensor:
        """[B,Tw,2] pinyin ids picked by argmax(pron_attn) (tasks/tts/dict_tts.py:295-304); -1 past the row end."""
        pa = _dev_f32(pron_attn, self.device)
        B, Tw, Lp = pa.shape
        pairs = torch.empty(B, Tw, 2, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            if pinyin is not None:
                py = _dev_i64(pinyin, self.device)
                if tuple(py.shape) != (B, Tw, Lp):
                    raise ValueError("pinyin must be [B,Tw,Lp] like pron_attn")
                rc = self.lib.dtts_pron_tokens(_ptr(pa), _ptr(py), None, None, B, Tw, Lp, _ptr(pairs), _stream())
            else:
                bank = getattr(self, "bank", None)
                if bank is None or dict_ids is None:
                    raise ValueError("pron_tokens needs the pinyin tensor, or dict_ids with a registered bank")
                ids = _dev_i64(dict_ids, self.device)
                st = bank.c_struct()
                rc = self.lib.dtts_pron_tokens(_ptr(pa), None, C.byref(st), _ptr(ids), B, Tw, Lp, _ptr(pairs), _stream())
            binding.check(rc, "pron_tokens")
        return pairs

    def length_regulate(self, dur_int, ilens):
        """LengthRegulator + pad to frames_multiple.  Syncs once to learn T (data-dependent shape)."""
        dev = self.device
        B, Tw = dur_int.shape
        cum = torch.empty(B, Tw, dtype=torch.int32, device=dev)
        totals = torch.empty(B + 1, dtype=torch.int32, device=dev)
        binding.check(self.lib.dtts_length_regulate_scan(self.handle, _ptr(dur_int), _ptr(ilens), B, Tw, _ptr(cum),
                                                         _ptr(totals), C.byref(self._t_raw), _stream()), "lr_scan")
        t_raw = int(self._t_raw.value)
        if t_raw <= 0:
            raise RuntimeError("length regulator produced an empty batch")
        fm = self.cfg.frames_multiple
        T = (t_raw + fm - 1) // fm * fm
        mel2word = torch.empty(B, T, dtype=torch.int64, device=dev)
        binding.check(self.lib.dtts_length_regulate_fill(self.handle, _ptr(cum), _ptr(ilens), B, Tw, t_raw, T,
                                                         _ptr(mel2word), _stream()), "lr_fill")
        return mel2word

    def expand(self, word_encoder_out, mel2word):
        dev = self.device
        B, Tw, H = word_encoder_out.shape
        T = mel2word.shape[1]
        decoder_inp = torch.empty(B, T, H, device=dev)
        g_bct = torch.empty(B, H, T, device=dev)
        x_mask = torch.empty(B, T, device=dev)
        binding.check(self.lib.dtts_expand(self.handle, _ptr(word_encoder_out), _ptr(mel2word), B, Tw, T,
                                           _ptr(decoder_inp), _ptr(g_bct), _ptr(x_mask), _stream()), "expand")
        return decoder_inp, g_bct, x_mask

    def decode_mel(self, g_bct, z_in):
        dev = self.device
        B, H, T = g_bct.shape
        z_in = _dev_f32(z_in, dev)
        if tuple(z_in.shape) != (B, self.cfg.latent, T // self.cfg.frames_multiple):
            raise ValueError(f"z_p must be [B,{self.cfg.latent},T/{self.cfg.frames_multiple}]")
        mel = torch.empty(B, T, self.cfg.n_mel, device=dev)
        z_p = torch.empty_like(z_in)
        nbytes = self.lib.dtts_decode_workspace_bytes(self.handle, B, T)
        ws = self.ws.get(nbytes)
        binding.check(self.lib.dtts_decode_mel(self.handle, _ptr(g_bct), _ptr(z_in), B, T, _ptr(mel), _ptr(z_p),
                                               _ptr(ws), ws.numel(), _stream()), "decode_mel")
        return mel, z_p

    # -- reference-shaped forward ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, txt_tokens, pron_modified, key_value_map=None, ph2word=None, word_len=None, dict_msg=None,
                mel2word=None, mel2ph=None, spk_embed=None, infer=True, tgt_mels=None, forward_post_glow=True,
                two_stage=True, z_p=None, dict_ids=None):
        """Same arguments as PortaSpeech_dict.forward (model.py:36-37).  Inference only.  ``z_p`` (extension):
        the N(0,1) prior sample; when None it is drawn exactly like the reference does -- on the host with the
        global CPU generator (fvae_semantics.py:110-111)."""
        if not infer:
            raise NotImplementedError("the B200 engine implements the inference path only (infer=True)")
        if spk_embed is not None:
            raise NotImplementedError("multi-speaker conditioning is off in dict_tts.yaml (use_spk_embed: false)")
        word_tokens = txt_tokens[0] if isinstance(txt_tokens, (tuple, list)) else txt_tokens
        prof = self.profile_infer
        with torch.cuda.device(self.device):
            ret = {}
            with Timer("encoder", enable=prof):                   # model.py:50 (run_text_encoder)
                with Timer("dict_encoder", enable=prof):          # model.py:86
                    if dict_ids is not None and dict_msg is None:     # extension: characters named by dictionary-bank id
                        t = self.text_encode_bank(word_tokens, pron_modified, dict_ids)
                    else:
                        keys, values, key_map, pinyin, pinyin_map = dict_msg
                        t = self.text_encode(word_tokens, pron_modified, keys, values, key_map, pinyin, pinyin_map)
                ret["dict_attn"], ret["rel"], ret["dp_attn"] = t["dict_attn"], None, None
                ret["pron_attn"], ret["dur"], ret["word_encoder_out"] = t["pron_attn"], t["dur"], t["word_encoder_out"]
                fm = self.cfg.frames_multiple
                if mel2word is None:
                    mel2word = self.length_regulate(t["dur_int"], t["ilens"])
                else:
                    mel2word = _dev_i64(mel2word, self.device)
                    if mel2word.shape[1] % fm:                    # model.py:98-100
                        pad = fm - mel2word.shape[1] % fm
                        mel2word = torch.cat([mel2word] + [mel2word[:, -1:]] * pad, -1).contiguous()
                ret["mel2word"] = mel2word
                decoder_inp, g_bct, x_mask = self.expand(t["word_encoder_out"], mel2word)
            ret["x_mask"], ret["decoder_inp"] = x_mask.unsqueeze(-1), decoder_inp
            ret["synta"] = torch.zeros_like(g_bct)
            B, _, T = g_bct.shape
            if z_p is None:
                z_p = torch.distributions.Normal(0, 1).sample([B, self.cfg.latent, T // fm])
            with Timer("fvae", enable=prof):                      # model.py:57
                mel, z_out = self.decode_mel(g_bct, z_p)
            ret["mel_out_fvae"] = ret["mel_out"] = mel
            ret["z_p"] = z_out
        return ret

    __call__ = forward


class PortaSpeechEngine(DictTTSEngine):
    """The PortaSpeech (non-dict) sibling, SURVEY.md §8f-3: ``forward`` keeps the call signature and the returned dict of
    ``PortaSpeech.forward`` (modules/portaspeech/model.py:202-237) at ``dur_level: word`` / ``use_post_glow: False`` (the
    only inference path the reference checkout can construct: ``modules/glow`` is absent from it).  It shares the length
    regulator and ``decode_mel`` with the dict model; the text side runs through ``dtts_ps_text_encode`` /
    ``dtts_ps_attend``."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Optional[PortaSpeechConfig] = None, device="cuda:0",
                 arena: Optional[torch.Tensor] = None, table=None, precision: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("PortaSpeechEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = binding.load()
        self.cfg = cfg or PortaSpeechConfig()
        self.device = torch.device(device)
        if arena is None:
            sd = {k: v for k, v in fold_weight_norm(state_dict).items()
                  if not k.startswith(("fvae.encoder.", "mel_disc.", "post_flow.", "g_proj.")) and v.is_floating_point()}
            host, table = pack_arena(sd)
            arena = host.to(self.device)
        self.arena, self.table = arena, table
        c = self.cfg
        desc = binding.AcousticDesc(c.hidden, c.n_heads, c.enc_layers, c.ffn_kernel, c.ffn_filter, c.dict_dim,
                                    c.word_size, c.pinyin_size, c.dur_layers, c.dur_kernel, c.dur_chans,
                                    c.frames_multiple, c.latent, c.dec_layers, c.dec_kernel, c.flow_hidden,
                                    c.flow_kernel, c.flow_blocks, c.flow_layers, c.n_mel, 0, int(precision), 0,
                                    binding.MODEL_PORTASPEECH, c.ph_size, c.word_enc_layers, c.rel_window)
        tab, self._keep = binding.make_table(table)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            binding.check(self.lib.dtts_acoustic_create(C.byref(desc), _ptr(self.arena), self.arena.numel(), tab,
                                                        len(table), _stream(), C.byref(self.handle)), "acoustic_create")
        self.ws = _Workspace(self.device)
        self._t_raw = C.c_int32(0)
        self.profile_infer = False

    def ps_text_encode(self, txt_tokens, ph2word, word_len: int):
        dev = self.device
        txt = _dev_i64(txt_tokens, dev)
        p2w = _dev_i64(ph2word, dev)
        B, Tp = txt.shape
        Tw = int(word_len)
        if tuple(p2w.shape) != (B, Tp) or Tw <= 0:
            raise ValueError("ph2word must be [B,Tp] like txt_tokens and word_len positive")
        H = self.cfg.hidden
        out = dict(ph_encoder_out=torch.empty(B, Tp, H, device=dev), word_encoder_out=torch.empty(B, Tw, H, device=dev),
                   dur=torch.empty(B, Tw, device=dev), dur_int=torch.empty(B, Tw, dtype=torch.int64, device=dev),
                   ilens=torch.empty(B, dtype=torch.int64, device=dev))
        tin = binding.PsTextIn(_ptr(txt), _ptr(p2w), B, Tp, Tw)
        tout = binding.PsTextOut(*[_ptr(out[k]) for k in ("ph_encoder_out", "word_encoder_out", "dur", "dur_int", "ilens")])
        ws = self.ws.get(self.lib.dtts_ps_text_workspace_bytes(self.handle, B, Tp, Tw))
        binding.check(self.lib.dtts_ps_text_encode(self.handle, C.byref(tin), C.byref(tout), _ptr(ws), ws.numel(),
                                                   _stream()), "ps_text_encode")
        out["ph2word"] = p2w
        return out

    def ps_attend(self, t, mel2word):
        dev = self.device
        B, Tp, H = t["ph_encoder_out"].shape
        Tw = t["word_encoder_out"].shape[1]
        T = mel2word.shape[1]
        attn = torch.empty(B, T, Tp, device=dev)
        decoder_inp = torch.empty(B, T, H, device=dev)
        g_bct = torch.empty(B, H, T, device=dev)
        x_mask = torch.empty(B, T, device=dev)
        ws = self.ws.get(self.lib.dtts_ps_attend_workspace_bytes(self.handle, B, Tp, Tw, T))
        binding.check(self.lib.dtts_ps_attend(self.handle, _ptr(t["ph_encoder_out"]), _ptr(t["word_encoder_out"]),
                                              _ptr(t["ph2word"]), _ptr(mel2word), B, Tp, Tw, T, _ptr(attn),
                                              _ptr(decoder_inp), _ptr(g_bct), _ptr(x_mask), _ptr(ws), ws.numel(),
                                              _stream()), "ps_attend")
        return attn, decoder_inp, g_bct, x_mask

    @torch.no_grad()
    def forward(self, txt_tokens, ph2word, word_len, mel2word=None, mel2ph=None, spk_embed=None, infer=True,
                tgt_mels=None, forward_post_glow=False, two_stage=True, z_p=None):
        """Same arguments as PortaSpeech.forward (portaspeech/model.py:202-203); ``z_p`` (extension) as in the dict engine."""
        if not infer:
            raise NotImplementedError("the B200 engine implements the inference path only (infer=True)")
        if spk_embed is not None:
            raise NotImplementedError("multi-speaker conditioning is off in ps_flow.yaml (use_spk_embed: false)")
        prof = self.profile_infer
        fm = self.cfg.frames_multiple
        with torch.cuda.device(self.device):
            ret = {}
            with Timer("encoder", enable=prof):
                t = self.ps_text_encode(txt_tokens, ph2word, int(word_len))
                ret.update(ph_encoder_out=t["ph_encoder_out"], word_encoder_out=t["word_encoder_out"], dur=t["dur"])
                if mel2word is None:
                    mel2word = self.length_regulate(t["dur_int"], t["ilens"])
                else:
                    mel2word = _dev_i64(mel2word, self.device)
                    if mel2word.shape[1] % fm:                    # model.py:250-252
                        pad = fm - mel2word.shape[1] % fm
                        mel2word = torch.cat([mel2word] + [mel2word[:, -1:]] * pad, -1).contiguous()
                ret["mel2word"] = mel2word
                attn, decoder_inp, g_bct, x_mask = self.ps_attend(t, mel2word)
            ret["attn"], ret["x_mask"], ret["decoder_inp"] = attn, x_mask.unsqueeze(-1), decoder_inp
            B, _, T = g_bct.shape
            if z_p is None:
                z_p = torch.distributions.Normal(0, 1).sample([B, self.cfg.latent, T // fm])
            with Timer("fvae", enable=prof):
                mel, z_out = self.decode_mel(g_bct, z_p)
            ret["mel_out_fvae"] = ret["mel_out"] = mel
            ret["z_p"] = z_out
        return ret

    __call__ = forward


class HifiGanEngine:
    """HiFi-GAN V1 generator: mel [B,T,80] -> wav [B,T*hop] on the device."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: Optional[VocoderConfig] = None, device="cuda:0",
                 arena: Optional[torch.Tensor] = None, table=None, precision: int = 6):
        if not torch.cuda.is_available():
            raise RuntimeError("HifiGanEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = binding.load()
        self.cfg = cfg or VocoderConfig()
        self.device = torch.device(device)
        if arena is None:
            host, table = pack_arena(fold_weight_norm(state_dict))
            arena = host.to(self.device)
        self.arena, self.table = arena, table
        for dl in self.cfg.rb_dilations:
            if len(dl) != 3:
                raise ValueError("ResBlock1 needs 3 dilations per block")
        self.precision = int(precision)
        self.handle, self._keep = self._create(self.precision)
        self.ws = _Workspace(self.device)
        self.profile_infer = False

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dtts_vocoder_destroy(self.handle)
            self.handle = None

    def _create(self, precision: int):
        c = self.cfg
        desc = binding.VocoderDesc()
        desc.n_mel, desc.init_ch, desc.n_ups, desc.n_rb = c.n_mel, c.init_ch, len(c.up_rates), len(c.rb_kernels)
        for i, (u, k) in enumerate(zip(c.up_rates, c.up_kernels)):
            desc.up_rates[i], desc.up_kernels[i] = u, k
        for j, (k, dl) in enumerate(zip(c.rb_kernels, c.rb_dilations)):
            desc.rb_kernels[j] = k
            for m in range(3):
                desc.rb_dilations[j][m] = dl[m]
        desc.precision = precision
        tab, keep = binding.make_table(self.table)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            binding.check(self.lib.dtts_vocoder_create(C.byref(desc), _ptr(self.arena), self.arena.numel(), tab,
                                                       len(self.table), _stream(), C.byref(handle)), "vocoder_create")
        return handle, keep

    @torch.no_grad()
    def self_check(self, tol: float = 1e-4, margin: float = 0.85, frames: int = 48, seed: int = 0,
                   mel: Optional[torch.Tensor] = None) -> Dict:
        """Guard for the reduced-precision tensor-core modes on THIS checkpoint: the waveform
        tolerance is absolute (1e-4 RMS) and the fixtures behind the default mode are random-weight, so a trained
        generator with a wider dynamic range could leave it -- fp16 operands saturate at 65504.  Runs a probe mel
        (U(mel_vmin, mel_vmax) unless given) through the current mode and through the fp32-class mode (precision 1: bf16
        hi/lo x hi/lo, 3 MMAs, bf16 range), both on the GPU, and if their RMS difference exceeds ``margin * tol`` switches
        this engine to precision 1 for good.  Returns {'precision', 'rms', 'switched'}."""
        from . import synth
        if self.precision in (0, 1):
            return dict(precision=self.precision, rms=0.0, switched=False)
        probe = synth.make_mel(seed, 2, frames) if mel is None else mel
        got = self.forward(probe)
        ref_handle, ref_keep = self._create(1)
        cur = (self.handle, self._keep, self.precision)
        self.handle, self._keep, self.precision = ref_handle, ref_keep, 1
        want = self.forward(probe)
        rms = float((got - want).double().pow(2).mean().sqrt())
        ok = bool(torch.isfinite(got).all()) and rms <= margin * tol
        if ok:                                            # keep the fast mode, drop the probe handle
            self.lib.dtts_vocoder_destroy(ref_handle)
            self.handle, self._keep, self.precision = cur
        else:                                             # stay on the fp32-class handle
            self.lib.dtts_vocoder_destroy(cur[0])
        return dict(precision=self.precision, rms=rms, switched=not ok)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.dtts_vocoder_launch_count(self.handle))

    @torch.no_grad()
    def forward(self, mel: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """mel [B,T,n_mel] -> wav [B,T*hop].  ``lengths`` (optional, [B] valid mel frames per item): samples before
        lengths[b]*hop are bit for bit those of the full-length call, samples after it are 0 and the work only they
        depend on is skipped (dtts_vocode_lens)."""
        mel = _dev_f32(mel, self.device)
        if mel.dim() != 3 or mel.shape[2] != self.cfg.n_mel:
            raise ValueError("mel must be [B,T,n_mel]")
        B, T, _ = mel.shape
        wav = torch.empty(B, T * self.cfg.hop, device=self.device)
        with torch.cuda.device(self.device), Timer("hifigan", enable=self.profile_infer):   # vocoders/hifigan.py:59
            ws = self.ws.get(self.lib.dtts_vocode_workspace_bytes(self.handle, B, T))
            if lengths is None:
                binding.check(self.lib.dtts_vocode(self.handle, _ptr(mel), B, T, _ptr(wav), _ptr(ws), ws.numel(),
                                                   _stream()), "vocode")
            else:
                lens = torch.as_tensor(lengths).to(self.device, torch.int32).contiguous()
                if tuple(lens.shape) != (B,):
                    raise ValueError("lengths must be [B]")
                lens = lens.clamp(0, T)
                binding.check(self.lib.dtts_vocode_lens(self.handle, _ptr(mel), _ptr(lens), B, T, _ptr(wav), _ptr(ws),
                                                        ws.numel(), _stream()), "vocode_lens")
        return wav

    __call__ = forward
