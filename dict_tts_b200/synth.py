"""Seeded synthetic checkpoints and Biaobei-shaped batches (no network: no real checkpoints/datasets).

The state dicts use the reference's checkpoint key names and shapes (SURVEY.md Appendix B;
``state_dict['model']`` of utils/trainer.py:436-449 and ``model_gen`` of vocoders/hifigan.py:19-24),
including weight-norm ``weight_g``/``weight_v`` pairs and a few dead-at-inference tensors, so that the
same dict loads into the reference modules (golden generation) and into this engine.

Batches follow the collater contract of tasks/tts/dataset_utils.py:264-302 and the dictionary schema of
data_gen/tts/binarizer_zh.py:236-314.
"""
import math
from typing import Dict, Optional

import torch

from .config import AcousticConfig, VocoderConfig


class _Init:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)
        self.sd: Dict[str, torch.Tensor] = {}

    def uniform(self, name, shape, bound):
        self.sd[name] = (torch.rand(shape, generator=self.g) * 2 - 1) * bound
        return self.sd[name]

    def normal(self, name, shape, std, mean=0.0):
        self.sd[name] = torch.randn(shape, generator=self.g) * std + mean
        return self.sd[name]

    def conv(self, prefix, cout, cin, k, bias=True, gain=1.0, wn=False, transposed=False):
        fan_in = cin * k
        bound = gain / math.sqrt(fan_in)
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        if wn:
            v = self.uniform(prefix + ".weight_v", shape, bound)
            norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
            self.sd[prefix + ".weight_g"] = norm * (0.8 + 0.4 * torch.rand(norm.shape, generator=self.g))
        else:
            self.uniform(prefix + ".weight", shape, bound)
        if bias:
            self.uniform(prefix + ".bias", (cout,), 0.5 / math.sqrt(fan_in))

    def linear(self, prefix, cout, cin, bias=True, gain=1.0):
        self.uniform(prefix + ".weight", (cout, cin), gain / math.sqrt(cin))
        if bias:
            self.uniform(prefix + ".bias", (cout,), 0.5 / math.sqrt(cin))

    def ln(self, prefix, c, names=("gamma", "beta")):
        self.normal(prefix + "." + names[0], (c,), 0.1, 1.0)
        self.normal(prefix + "." + names[1], (c,), 0.1)


def make_acoustic_state_dict(seed: int = 1234, cfg: Optional[AcousticConfig] = None,
                             n_vocab: int = 6, with_dead: bool = True) -> Dict[str, torch.Tensor]:
    """Checkpoint-format ``state_dict['model']`` of PortaSpeech_dict (modules/dict_tts/model.py:14-33)."""
    cfg = cfg or AcousticConfig()
    H = cfg.hidden
    it = _Init(seed)
    if with_dead:  # dead at inference, must be tolerated by the loader
        it.linear("enc_pos_proj", H, 2 * H)
        it.linear("dec_query_proj", H, 2 * H)
        it.linear("dec_res_proj", H, 2 * H)
        it.uniform("attn.in_proj_weight", (3 * H, H), 1 / math.sqrt(H))
        it.uniform("attn.out_proj.weight", (H, H), 1 / math.sqrt(H))
    # duration predictor (portaspeech/model.py:38-66)
    for i in range(cfg.dur_layers):
        it.conv(f"dur_predictor.conv.{i}.1", cfg.dur_chans, H if i == 0 else cfg.dur_chans, cfg.dur_kernel)
        it.ln(f"dur_predictor.conv.{i}.3", cfg.dur_chans, names=("weight", "bias"))
    it.linear("dur_predictor.linear.0", 1, cfg.dur_chans, gain=0.5)
    it.sd["dur_predictor.linear.0.bias"] = torch.tensor([2.3])   # log-durations around exp(2.3)-1 ~ 9 frames
    # FVAE (fvae_semantics.py:62-83)
    it.conv("fvae.g_pre_net.0", H, H, 8)
    for f in range(cfg.flow_blocks):
        p = f"fvae.prior_flow.flows.{2 * f}"
        it.conv(p + ".pre", cfg.flow_hidden, cfg.latent // 2, 1)
        for j in range(cfg.flow_layers):
            it.conv(p + f".enc.in_layers.{j}", 2 * cfg.flow_hidden, cfg.flow_hidden, cfg.flow_kernel, wn=True)
        for j in range(cfg.flow_layers):
            rs = 2 * cfg.flow_hidden if j < cfg.flow_layers - 1 else cfg.flow_hidden
            it.conv(p + f".enc.res_skip_layers.{j}", rs, cfg.flow_hidden, 1, wn=True)
        it.conv(p + ".enc.cond_layer", 2 * cfg.flow_hidden * cfg.flow_layers, H, 1, wn=True)
        it.conv(p + ".post", cfg.latent // 2, cfg.flow_hidden, 1, gain=0.5)   # zero-init upstream; non-trivial here
    it.conv("fvae.decoder.pre_net.0", H, cfg.latent, 4, transposed=True)
    for j in range(cfg.dec_layers):
        it.conv(f"fvae.decoder.wn.in_layers.{j}", 2 * H, H, cfg.dec_kernel, wn=True)
    for j in range(cfg.dec_layers):
        rs = 2 * H if j < cfg.dec_layers - 1 else H
        it.conv(f"fvae.decoder.wn.res_skip_layers.{j}", rs, H, 1, wn=True)
    it.conv("fvae.decoder.wn.cond_layer", 2 * H * cfg.dec_layers, H, 1, wn=True)
    it.conv("fvae.decoder.out_proj", cfg.n_mel, H, 1)
    # dict encoder (dict_encoder.py:68-163)
    p = "dict_encoder.S2PA_module"
    if with_dead:
        it.normal(p + ".emb.weight", (n_vocab, H), H ** -0.5)
    w = it.normal(p + ".word_emb.weight", (cfg.word_size, H), H ** -0.5)
    w[0].zero_()                                                   # padding_idx=0
    for enc in ("semantic_encoder", "linguistic_encoder"):
        q = f"{p}.{enc}"
        for i in range(cfg.enc_layers):
            for c in "qkvo":
                it.conv(f"{q}.attn_layers.{i}.conv_{c}", H, H, 1)
        for i in range(cfg.enc_layers):
            it.ln(f"{q}.norm_layers_1.{i}", H)
        for i in range(cfg.enc_layers):
            it.conv(f"{q}.ffn_layers.{i}.conv_1", cfg.ffn_filter, H, cfg.ffn_kernel)
            it.conv(f"{q}.ffn_layers.{i}.conv_2", H, cfg.ffn_filter, 1)
        for i in range(cfg.enc_layers):
            it.ln(f"{q}.norm_layers_2.{i}", H)
        it.ln(f"{q}.last_ln", H)
    a = p + ".s2pa_attention"
    it.linear(a + ".q_transform", H, H, bias=False, gain=2.0)
    it.linear(a + ".k_transform", H, cfg.dict_dim, bias=False, gain=4.0)
    it.linear(a + ".v_transform", H, cfg.dict_dim, bias=False)
    it.linear(a + ".output_transform", H, H, bias=False)
    e = it.normal(a + ".pinyin_embedding.weight", (cfg.pinyin_size, H), 0.3)
    e[0].zero_()                                                   # padding_idx=0
    return it.sd


def make_vocoder_state_dict(seed: int = 4321, cfg: Optional[VocoderConfig] = None,
                            hot: float = 1.0) -> Dict[str, torch.Tensor]:
    """Checkpoint-format ``model_gen`` of HifiGanGenerator (modules/hifigan/hifigan.py:101-124).

    hot: "trained-like" dynamic range -- conv_pre is scaled by ``hot`` and conv_post by ``1 / hot``, so every activation
    between them is ``hot`` times larger (the default init keeps them at O(1..10)) while the waveform stays in range.
    hot ~ 1e2 puts stage-1 activations at 1e2..1e3; hot ~ 2e4 pushes them past the fp16 range (65504)."""
    cfg = cfg or VocoderConfig()
    it = _Init(seed)
    it.conv("conv_pre", cfg.init_ch, cfg.n_mel, 7, wn=True)
    ch = cfg.init_ch
    for i, (u, k) in enumerate(zip(cfg.up_rates, cfg.up_kernels)):
        # a stride-u transposed conv sums cin*k/u taps per output sample
        bound = 1.7 / math.sqrt(ch * k / u)
        v = it.uniform(f"ups.{i}.weight_v", (ch, ch // 2, k), bound)
        norm = v.flatten(1).norm(dim=1).view(-1, 1, 1)
        it.sd[f"ups.{i}.weight_g"] = norm * (0.8 + 0.4 * torch.rand(norm.shape, generator=it.g))
        it.uniform(f"ups.{i}.bias", (ch // 2,), 0.05)
        ch //= 2
        for j, k_r in enumerate(cfg.rb_kernels):
            r = f"resblocks.{i * len(cfg.rb_kernels) + j}"
            for m in range(len(cfg.rb_dilations[j])):
                it.conv(f"{r}.convs1.{m}", ch, ch, k_r, wn=True, gain=1.2)
                it.conv(f"{r}.convs2.{m}", ch, ch, k_r, wn=True, gain=1.2)
    it.conv("conv_post", 1, ch, 7, wn=True, gain=0.6)
    if hot != 1.0:
        it.sd["conv_pre.weight_g"] = it.sd["conv_pre.weight_g"] * hot
        it.sd["conv_pre.bias"] = it.sd["conv_pre.bias"] * hot
        it.sd["conv_post.weight_g"] = it.sd["conv_post.weight_g"] / hot
    return it.sd


# ----------------------------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------------------------

_NPRON_P = [0.86, 0.09, 0.03, 0.012, 0.005, 0.003]     # zh-dict.json statistics


def make_batch(seed: int = 1234, B: int = 60, min_chars: int = 12, max_chars: int = 20,
               max_frames: int = 400, Lk_cap: int = 96, dict_dim: int = 768, word_size: int = 8000,
               pinyin_size: int = 185, pron_modified_p: float = 0.02, frames_multiple: int = 4,
               with_mel2word: bool = True, alias_values: bool = False) -> Dict[str, torch.Tensor]:
    """One collated batch in the layout DictTTSDataset.collater produces (dataset_utils.py:264-302).

    alias_values: ``values`` IS the ``keys`` tensor, as the binarized data has it (one feature tensor stored as both
    'key' and 'value', data_gen/tts/binarizer_zh.py:231-233) and as data.DictTTSTestSet keeps it; default: a separate
    equal tensor, as the reference collater builds it.

    Returns word_tokens[B,Tw] i64, pron_modified[B,Tw] i64, keys/values[B,Tw,Lk,768] f32, key_map[B,Tw,Lk] f32,
    pinyin/pinyin_map[B,Tw,Lp] i64, word_lengths[B], mel2word[B,T] i64 (supplied durations: the reference's own
    profiling mode, tasks/tts/dict_tts.py:194), z_p[B,16,T/4] drawn from the CPU generator
    (fvae_semantics.py:110-111), mel_lengths[B].
    """
    g = torch.Generator().manual_seed(seed)
    n_chars = torch.randint(min_chars, max_chars + 1, (B,), generator=g)
    n_chars[0] = max_chars                                   # sorted-by-length bucket: longest first
    n_chars, _ = torch.sort(n_chars, descending=True)
    Tw = int(n_chars.max()) + 2
    per_utt = []
    Lk = 3
    Lp = 2
    for b in range(B):
        n = int(n_chars[b])
        npron = torch.multinomial(torch.tensor(_NPRON_P), n, replacement=True, generator=g) + 1
        chars = []
        for c in range(n):
            segs = []
            total = 0
            for i in range(int(npron[c])):
                ln = int(torch.randint(6, 31, (1,), generator=g)) + 2   # min(len(gloss),30)+2 tokens
                if total + ln > Lk_cap:
                    ln = Lk_cap - total
                    if ln < 3:
                        break
                segs.append(ln)
                total += ln
            chars.append(segs)
            Lk = max(Lk, total)
            Lp = max(Lp, 2 * len(segs))
        per_utt.append(chars)
    word_tokens = torch.zeros(B, Tw, dtype=torch.long)
    key_map = torch.zeros(B, Tw, Lk)
    keys = torch.zeros(B, Tw, Lk, dict_dim)
    pinyin = torch.zeros(B, Tw, Lp, dtype=torch.long)
    pinyin_map = torch.zeros(B, Tw, Lp, dtype=torch.long)
    pron_modified = torch.zeros(B, Tw, dtype=torch.long)
    for b in range(B):
        chars = per_utt[b]
        n = len(chars)
        word_tokens[b, 0] = 1                                # BOS/EOS rows: keys 0, key_map 1, pinyin 0, pinyin_map 1
        word_tokens[b, n + 1] = 1
        word_tokens[b, 1:n + 1] = torch.randint(3, word_size, (n,), generator=g)
        key_map[b, 0] = 1
        key_map[b, n + 1] = 1
        pinyin_map[b, 0] = 1
        pinyin_map[b, n + 1] = 1
        for c, segs in enumerate(chars):
            t = c + 1
            o = 0
            for i, ln in enumerate(segs):
                key_map[b, t, o + 1:o + ln - 1] = i + 1
                keys[b, t, o:o + ln] = torch.randn(ln, dict_dim, generator=g) * 0.5
                o += ln
                pinyin[b, t, 2 * i:2 * i + 2] = torch.randint(1, pinyin_size, (2,), generator=g)
                pinyin_map[b, t, 2 * i:2 * i + 2] = i + 1
            if len(segs) > 1 and float(torch.rand(1, generator=g)) < pron_modified_p * 8:
                pron_modified[b, t] = int(torch.randint(1, len(segs) + 1, (1,), generator=g))
    values = keys if alias_values else keys.clone()          # binarizer stores the same LM features as key and value
    out = dict(word_tokens=word_tokens, pron_modified=pron_modified, keys=keys, values=values, key_map=key_map,
               pinyin=pinyin, pinyin_map=pinyin_map, word_lengths=n_chars + 2)
    if with_mel2word:
        mel2word = torch.zeros(B, max_frames, dtype=torch.long)
        mel_lengths = torch.zeros(B, dtype=torch.long)
        for b in range(B):
            n = int(n_chars[b]) + 2
            d = torch.randint(8, 29, (n,), generator=g).float()
            target = max_frames if b == 0 else int(max_frames * (0.75 + 0.25 * float(torch.rand(1, generator=g))))
            d = torch.clamp((d * target / d.sum()).floor(), min=1).long()
            d[-1] += target - int(d.sum())
            if d[-1] < 1:
                d[-1] = 1
            m = torch.repeat_interleave(torch.arange(1, n + 1), d)[:max_frames]
            mel2word[b, :m.numel()] = m
            mel_lengths[b] = m.numel()
        assert max_frames % frames_multiple == 0
        out["mel2word"] = mel2word
        out["mel_lengths"] = mel_lengths
        out["z_p"] = draw_z(B, 16, max_frames // frames_multiple, seed + 7)
    return out


def draw_z(B: int, latent: int, T_sqz: int, seed: int) -> torch.Tensor:
    """Prior sample exactly as the reference draws it: torch.distributions.Normal(0,1).sample on the CPU
    generator (fvae_semantics.py:110-111).  Callers pass the same tensor to the engine and to the oracle."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    z = torch.distributions.Normal(0, 1).sample([B, latent, T_sqz])
    torch.random.set_rng_state(state)
    return z


def make_mel(seed: int, B: int, T: int, n_mel: int = 80) -> torch.Tensor:
    """cfg-4 vocoder input: mel ~ U(mel_vmin, mel_vmax) = U(-6, 1.5) (tts/base.yaml:59-60), layout [B,T,80]."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, T, n_mel, generator=g) * 7.5 - 6.0


# ---------------------------------------------------------------------------------------------------------------
# PortaSpeech (non-dict) sibling, SURVEY.md §8f-3
# ---------------------------------------------------------------------------------------------------------------
def make_ps_state_dict(seed: int = 2468, cfg=None) -> Dict[str, torch.Tensor]:
    """Checkpoint-format ``state_dict['model']`` of PortaSpeech (modules/portaspeech/model.py:132-200) with
    ``use_post_glow: False``; key names and shapes as tests/golden/ps_small.npz holds them (written from the real model)."""
    from .config import PortaSpeechConfig
    cfg = cfg or PortaSpeechConfig()
    H, F, dk = cfg.hidden, cfg.ffn_filter, cfg.hidden // cfg.n_heads
    it = _Init(seed)
    e = it.normal("ph_encoder.emb.weight", (cfg.ph_size, H), H ** -0.5)
    e[0].zero_()                                                   # padding_idx=0
    for i in range(3):                                             # ConvReluNorm(kernel 5, 3 layers)
        it.conv(f"ph_encoder.pre.conv_layers.{i}", H, H, 5)
        it.ln(f"ph_encoder.pre.norm_layers.{i}", H)
    it.conv("ph_encoder.pre.proj", H, H, 1, gain=0.5)              # zero-init upstream; non-trivial here
    q = "ph_encoder.encoder"
    for i in range(cfg.enc_layers):
        it.normal(f"{q}.attn_layers.{i}.emb_rel_k", (1, 2 * cfg.rel_window + 1, dk), dk ** -0.5)
        it.normal(f"{q}.attn_layers.{i}.emb_rel_v", (1, 2 * cfg.rel_window + 1, dk), dk ** -0.5)
        for c in "qkvo":
            it.conv(f"{q}.attn_layers.{i}.conv_{c}", H, H, 1)
        it.ln(f"{q}.norm_layers_1.{i}", H)
        it.conv(f"{q}.ffn_layers.{i}.conv_1", F, H, cfg.ffn_kernel)
        it.conv(f"{q}.ffn_layers.{i}.conv_2", H, F, 1)
        it.ln(f"{q}.norm_layers_2.{i}", H)
    it.sd["word_encoder.pos_embed_alpha"] = torch.tensor([1.0])
    it.sd["word_encoder.embed_positions._float_tensor"] = torch.zeros(1)
    for i in range(cfg.word_enc_layers):
        w = f"word_encoder.layers.{i}.op"
        it.ln(w + ".layer_norm1", H, names=("weight", "bias"))
        it.uniform(w + ".self_attn.in_proj_weight", (3 * H, H), 1 / math.sqrt(H))
        it.uniform(w + ".self_attn.out_proj.weight", (H, H), 1 / math.sqrt(H))
        it.ln(w + ".layer_norm2", H, names=("weight", "bias"))
        it.conv(w + ".ffn.ffn_1", F, H, 1)
        it.linear(w + ".ffn.ffn_2", H, F)
    it.ln("word_encoder.layer_norm", H, names=("weight", "bias"))
    it.linear("enc_pos_proj", H, 2 * H)
    it.linear("dec_query_proj", H, 2 * H)
    it.linear("dec_res_proj", H, 2 * H)
    it.uniform("attn.in_proj_weight", (3 * H, H), 2 / math.sqrt(H))
    it.uniform("attn.out_proj.weight", (H, H), 1 / math.sqrt(H))
    for i in range(cfg.dur_layers):
        it.conv(f"dur_predictor.conv.{i}.1", cfg.dur_chans, H if i == 0 else cfg.dur_chans, cfg.dur_kernel)
        it.ln(f"dur_predictor.conv.{i}.3", cfg.dur_chans, names=("weight", "bias"))
    it.linear("dur_predictor.linear.0", 1, cfg.dur_chans, gain=0.5)
    it.sd["dur_predictor.linear.0.bias"] = torch.tensor([0.1])     # phoneme-level log-durations are SUMMED per word
    it.conv("fvae.g_pre_net.0", H, H, 8)
    for f in range(cfg.flow_blocks):
        p = f"fvae.prior_flow.flows.{2 * f}"
        it.conv(p + ".pre", cfg.flow_hidden, cfg.latent // 2, 1)
        for j in range(cfg.flow_layers):
            it.conv(p + f".enc.in_layers.{j}", 2 * cfg.flow_hidden, cfg.flow_hidden, cfg.flow_kernel, wn=True)
        for j in range(cfg.flow_layers):
            rs = 2 * cfg.flow_hidden if j < cfg.flow_layers - 1 else cfg.flow_hidden
            it.conv(p + f".enc.res_skip_layers.{j}", rs, cfg.flow_hidden, 1, wn=True)
        it.conv(p + ".enc.cond_layer", 2 * cfg.flow_hidden * cfg.flow_layers, H, 1, wn=True)
        it.conv(p + ".post", cfg.latent // 2, cfg.flow_hidden, 1, gain=0.5)
    it.conv("fvae.decoder.pre_net.0", H, cfg.latent, 4, transposed=True)
    for j in range(cfg.dec_layers):
        it.conv(f"fvae.decoder.wn.in_layers.{j}", 2 * H, H, cfg.dec_kernel, wn=True)
    for j in range(cfg.dec_layers):
        rs = 2 * H if j < cfg.dec_layers - 1 else H
        it.conv(f"fvae.decoder.wn.res_skip_layers.{j}", rs, H, 1, wn=True)
    it.conv("fvae.decoder.wn.cond_layer", 2 * H * cfg.dec_layers, H, 1, wn=True)
    it.conv("fvae.decoder.out_proj", cfg.n_mel, H, 1)
    return it.sd


def make_ps_batch(seed: int = 31, B: int = 4, min_words: int = 3, max_words: int = 9, max_ph_per_word: int = 4,
                  max_frames: int = 64, ph_size: int = 80, frames_multiple: int = 4) -> Dict[str, torch.Tensor]:
    """One collated word-level PortaSpeech batch (FastSpeechWordDataset.collater, tasks/tts/dataset_utils.py): phoneme ids
    ``txt_tokens [B,Tp]``, ``ph2word [B,Tp]`` (1-based, 0 = padding), ``word_lengths [B]``, supplied ``mel2word [B,T]``,
    ``mel_lengths [B]`` and the prior sample ``z_p [B,16,T/4]``."""
    g = torch.Generator().manual_seed(seed)
    n_words = torch.randint(min_words, max_words + 1, (B,), generator=g)
    n_words[0] = max_words
    n_words, _ = torch.sort(n_words, descending=True)
    per_word = [torch.randint(1, max_ph_per_word + 1, (int(n),), generator=g) for n in n_words]
    Tp = max(int(p.sum()) for p in per_word)
    txt = torch.zeros(B, Tp, dtype=torch.long)
    ph2word = torch.zeros(B, Tp, dtype=torch.long)
    mel2word = torch.zeros(B, max_frames, dtype=torch.long)
    mel_lengths = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        n = int(n_words[b])
        seg = torch.repeat_interleave(torch.arange(1, n + 1), per_word[b])
        ph2word[b, :seg.numel()] = seg
        txt[b, :seg.numel()] = torch.randint(3, ph_size, (seg.numel(),), generator=g)
        d = torch.randint(4, 15, (n,), generator=g).float()
        target = max_frames if b == 0 else int(max_frames * (0.7 + 0.3 * float(torch.rand(1, generator=g))))
        d = torch.clamp((d * target / d.sum()).floor(), min=1).long()
        d[-1] += target - int(d.sum())
        if d[-1] < 1:
            d[-1] = 1
        m = torch.repeat_interleave(torch.arange(1, n + 1), d)[:max_frames]
        mel2word[b, :m.numel()] = m
        mel_lengths[b] = m.numel()
    assert max_frames % frames_multiple == 0
    return dict(txt_tokens=txt, ph2word=ph2word, word_lengths=n_words, mel2word=mel2word, mel_lengths=mel_lengths,
                z_p=draw_z(B, 16, max_frames // frames_multiple, seed + 7))
