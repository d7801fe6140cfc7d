"""TEST INFRASTRUCTURE ONLY -- CPU/PyTorch fp32 restatement of the Dict-TTS text->mel->wav forward.

This file is the parity oracle for the CUDA engine.  It is NOT part of the product path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned
by running the reference's own modules in the build container (oracle/make_golden.py asserts agreement
<= 2e-5 max-abs stage by stage and writes tests/golden/*.npz from the *reference* outputs).

All functions operate on a dict ``W`` of weight-norm-folded fp32 tensors keyed by the reference's
checkpoint names (see dict_tts_b200/weights.py:fold_weight_norm).  Layout conventions follow the
reference: activations [B, C, T] inside encoders/decoders.
"""
import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# text encoder  (modules/commons/rel_transformer_encoder.py)
# ---------------------------------------------------------------------------------------------

def channel_layer_norm(x, gamma, beta, eps=1e-4):
    """LayerNorm over the channel axis of [B,C,T] (rel_transformer_encoder.py:261-279)."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def self_attention(W, p, x, attn_mask, n_heads):
    """MultiHeadAttention.forward/attention with window_size=None (rel_transformer_encoder.py:117-158)."""
    B, C, T = x.shape
    dk = C // n_heads
    q = F.conv1d(x, W[p + ".conv_q.weight"], W[p + ".conv_q.bias"])
    k = F.conv1d(x, W[p + ".conv_k.weight"], W[p + ".conv_k.bias"])
    v = F.conv1d(x, W[p + ".conv_v.weight"], W[p + ".conv_v.bias"])
    q = q.view(B, n_heads, dk, T).transpose(2, 3)
    k = k.view(B, n_heads, dk, T).transpose(2, 3)
    v = v.view(B, n_heads, dk, T).transpose(2, 3)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p_attn = F.softmax(scores, dim=-1)
    o = torch.matmul(p_attn, v).transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(o, W[p + ".conv_o.weight"], W[p + ".conv_o.bias"])


def encoder(W, p, x, x_mask, n_layers, n_heads, kernel):
    """Pre-LN Encoder.forward (rel_transformer_encoder.py:55-79); FFN is conv-k/ReLU/1x1 (:250-258)."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    for i in range(n_layers):
        x = x * x_mask
        h = channel_layer_norm(x, W[f"{p}.norm_layers_1.{i}.gamma"], W[f"{p}.norm_layers_1.{i}.beta"])
        x = x + self_attention(W, f"{p}.attn_layers.{i}", h, attn_mask, n_heads)
        h = channel_layer_norm(x, W[f"{p}.norm_layers_2.{i}.gamma"], W[f"{p}.norm_layers_2.{i}.beta"])
        f = F.conv1d(h * x_mask, W[f"{p}.ffn_layers.{i}.conv_1.weight"], W[f"{p}.ffn_layers.{i}.conv_1.bias"],
                     padding=kernel // 2)
        f = torch.relu(f)
        f = F.conv1d(f * x_mask, W[f"{p}.ffn_layers.{i}.conv_2.weight"], W[f"{p}.ffn_layers.{i}.conv_2.bias"])
        x = x + f * x_mask
    x = channel_layer_norm(x, W[f"{p}.last_ln.gamma"], W[f"{p}.last_ln.beta"])
    return x * x_mask


# ---------------------------------------------------------------------------------------------
# S2PA dictionary attention  (modules/dict_tts/layers/dict_encoder.py:32-66, layers/utils.py)
# ---------------------------------------------------------------------------------------------

def s2pa_attention(W, p, x, keys, values, key_map, pinyin, pinyin_map, pron_modified, language_zh=True):
    """x [B,H,Tw]; keys/values [B,Tw,Lk,D]; key_map [B,Tw,Lk]; pinyin/pinyin_map [B,Tw,Lp].
    Returns context [B,H,Tw], align [B,1,Lk,Tw], pron [B,H,Tw], pron_weights [B,Tw,Lp]."""
    B, H, Tw = x.shape
    D = keys.shape[-1]
    q = F.linear(x.transpose(1, 2), W[p + ".q_transform.weight"]) * (D ** -0.5)       # [B,Tw,H]; scale is key_size^-1/2
    k = F.linear(keys, W[p + ".k_transform.weight"])                                  # [B,Tw,Lk,H]
    v = F.linear(values, W[p + ".v_transform.weight"])
    logits = torch.einsum("btlh,bth->btl", k, q)
    logits = torch.where(key_map != 0, logits, torch.full_like(logits, -1e9))         # mask_logits, utils.py:40-47
    weights = F.softmax(logits, dim=-1)                                               # [B,Tw,Lk]
    align = weights.permute(0, 2, 1).unsqueeze(1)                                     # [B,1,Lk,Tw]
    context = torch.einsum("btl,btlh->bth", weights, v)
    context = F.linear(context, W[p + ".output_transform.weight"]).transpose(1, 2)    # [B,H,Tw]
    # pronunciation weights: segment-sum of attention mass per pronunciation id (utils.py:49-58)
    kmax = int(key_map.max().item())
    pm = pinyin_map.unsqueeze(-1)                                                     # [B,Tw,Lp,1]
    same = (key_map.unsqueeze(2) == pm.to(key_map.dtype)) & (pm >= 1) & (pm <= kmax)  # [B,Tw,Lp,Lk]
    pron_w = (same.to(weights.dtype) * weights.unsqueeze(2)).sum(-1)                  # [B,Tw,Lp]
    if language_zh and pron_modified is not None:                                     # add_pron_rule, utils.py:109-115
        pmax = int(pinyin_map.max().item())
        sel = (pron_modified >= 1) & (pron_modified <= pmax)
        onehot = (pinyin_map == pron_modified.unsqueeze(-1)).to(weights.dtype)
        forced = torch.where(sel.unsqueeze(-1), onehot, pron_w)
        pron_w = forced - pron_w + pron_w                                             # same op order as the reference
    emb = F.embedding(pinyin, W[p + ".pinyin_embedding.weight"])                      # [B,Tw,Lp,H]
    pron = torch.einsum("btp,btph->bth", pron_w, emb).transpose(1, 2)
    return context, align, pron, pron_w


def text_encode(W, cfg, word_tokens, pron_modified, keys, values, key_map, pinyin, pinyin_map):
    """DictEncoder.forward -> S2PATextEncoder.forward (dict_encoder.py:130-144,165-172).
    Returns word_encoder_out [B,Tw,H], dict_attn [B,1,Lk,Tw], pron_attn [B,Tw,Lp], context [B,Tw,H]."""
    p = "dict_encoder.S2PA_module"
    H = cfg.hidden
    x_lengths = (word_tokens > 0).long().sum(-1)
    x = F.embedding(word_tokens, W[p + ".word_emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, 2)
    Tw = x.shape[2]
    x_mask = (torch.arange(Tw, device=x.device).unsqueeze(0) < x_lengths.unsqueeze(1)).unsqueeze(1).to(x.dtype)
    x = encoder(W, p + ".semantic_encoder", x, x_mask, cfg.enc_layers, cfg.n_heads, cfg.ffn_kernel)
    context, dict_attn, pron, pron_attn = s2pa_attention(
        W, p + ".s2pa_attention", x, keys, values, key_map, pinyin, pinyin_map, pron_modified, cfg.language_zh)
    context = context * x_mask
    x = encoder(W, p + ".linguistic_encoder", context + pron, x_mask, cfg.enc_layers, cfg.n_heads, cfg.ffn_kernel)
    x = x.transpose(1, 2) * (word_tokens > 0).float().unsqueeze(-1)
    return x, dict_attn, pron_attn, context.transpose(1, 2)


# ---------------------------------------------------------------------------------------------
# duration predictor + length regulator
# ---------------------------------------------------------------------------------------------

def duration_predictor(W, cfg, dur_input):
    """DurationPredictor.forward (portaspeech/model.py:58-66); dur_input [B,Tw,H] already masked.
    src_padding is recomputed from the input exactly like add_dur does (dict_tts/model.py:73)."""
    src_padding = dur_input.abs().sum(-1) == 0
    keep = (1 - src_padding.float()).unsqueeze(1)
    xs = dur_input.transpose(1, 2)
    pad = (cfg.dur_kernel - 1) // 2
    for i in range(cfg.dur_layers):
        xs = F.conv1d(F.pad(xs, (pad, pad)), W[f"dur_predictor.conv.{i}.1.weight"], W[f"dur_predictor.conv.{i}.1.bias"])
        xs = torch.relu(xs)
        xs = F.layer_norm(xs.transpose(1, 2), (xs.shape[1],), W[f"dur_predictor.conv.{i}.3.weight"],
                          W[f"dur_predictor.conv.{i}.3.bias"], eps=1e-5).transpose(1, 2)
        xs = xs * keep
    d = F.softplus(F.linear(xs.transpose(1, 2), W["dur_predictor.linear.0.weight"], W["dur_predictor.linear.0.bias"]))
    return d[:, :, 0] * (1 - src_padding.float()), src_padding


def durations_to_int(dur):
    """log-scale durations -> integer frames (dict_tts/model.py:77-80): round half-to-even, clamp >= 0."""
    return torch.clamp(torch.round(dur.exp() - 1), min=0).long()


def length_regulate(dur_int, ilens, frames_multiple=4):
    """LengthRegulator.forward (tts_modules.py:215-251) + pad-to-multiple by repeating the last column
    (dict_tts/model.py:98-100).  Pure integer arithmetic, written as loops on purpose."""
    B = dur_int.shape[0]
    rows = []
    for b in range(B):
        d = [int(v) for v in dur_int[b, :int(ilens[b])]]
        if sum(d) == 0:
            d = [1] * len(d)
        row = []
        for w, n in enumerate(d):
            row.extend([w + 1] * n)
        rows.append(row)
    T = max(len(r) for r in rows)
    m = torch.zeros(B, T, dtype=torch.long)
    for b, r in enumerate(rows):
        if r:
            m[b, :len(r)] = torch.tensor(r, dtype=torch.long)
    if T % frames_multiple:
        extra = frames_multiple - T % frames_multiple
        m = torch.cat([m] + [m[:, -1:]] * extra, dim=1)
    return m


def expand_by_mel2word(word_encoder_out, mel2word):
    """F.pad one zero row + torch.gather (dict_tts/model.py:105-107), then x * tgt_nonpadding (:53)."""
    B, Tw, H = word_encoder_out.shape
    padded = torch.cat([torch.zeros(B, 1, H, dtype=word_encoder_out.dtype, device=word_encoder_out.device),
                        word_encoder_out], dim=1)
    x = torch.gather(padded, 1, mel2word.unsqueeze(-1).expand(-1, -1, H))
    nonpad = (mel2word > 0).float().unsqueeze(-1)
    return x * nonpad, nonpad


# ---------------------------------------------------------------------------------------------
# FVAE decoder + prior flow
# ---------------------------------------------------------------------------------------------

def wavenet(W, p, x, g, hidden, n_layers, kernel):
    """WN.forward with x_mask = 1 (modules/commons/wavenet.py:54-78)."""
    out = torch.zeros_like(x)
    cond = F.conv1d(g, W[p + ".cond_layer.weight"], W[p + ".cond_layer.bias"])
    for i in range(n_layers):
        a = F.conv1d(x, W[f"{p}.in_layers.{i}.weight"], W[f"{p}.in_layers.{i}.bias"], padding=kernel // 2)
        a = a + cond[:, 2 * hidden * i:2 * hidden * (i + 1)]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = F.conv1d(acts, W[f"{p}.res_skip_layers.{i}.weight"], W[f"{p}.res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = x + rs[:, :hidden]
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out


def prior_flow_reverse(W, cfg, z, g_sqz):
    """ResidualCouplingBlock.forward(reverse=True), mean_only, masks = 1 (glow_modules.py:108-128,157-163)."""
    half = cfg.latent // 2
    for f in reversed(range(cfg.flow_blocks)):
        z = torch.flip(z, [1])                                   # Flip comes first when iterating reversed(flows)
        p = f"fvae.prior_flow.flows.{2 * f}"
        x0, x1 = z[:, :half], z[:, half:]
        h = F.conv1d(x0, W[p + ".pre.weight"], W[p + ".pre.bias"])
        h = wavenet(W, p + ".enc", h, g_sqz, cfg.flow_hidden, cfg.flow_layers, cfg.flow_kernel)
        m = F.conv1d(h, W[p + ".post.weight"], W[p + ".post.bias"])
        z = torch.cat([x0, x1 - m], 1)
    return z


def decode_mel(W, cfg, decoder_inp, z):
    """FVAE_semantics.forward(infer=True) (fvae_semantics.py:84-115) with semantics = 0.
    decoder_inp [B,T,H] (already multiplied by tgt_nonpadding), z [B,latent,T/4] -> mel [B,T,80], z_p."""
    g = decoder_inp.transpose(1, 2)
    g_sqz = F.conv1d(g, W["fvae.g_pre_net.0.weight"], W["fvae.g_pre_net.0.bias"], stride=4, padding=2)
    z_p = prior_flow_reverse(W, cfg, z, g_sqz)
    x = F.conv_transpose1d(z_p, W["fvae.decoder.pre_net.0.weight"], W["fvae.decoder.pre_net.0.bias"], stride=4)
    x = wavenet(W, "fvae.decoder.wn", x, g, cfg.hidden, cfg.dec_layers, cfg.dec_kernel)
    mel = F.conv1d(x, W["fvae.decoder.out_proj.weight"], W["fvae.decoder.out_proj.bias"])
    return mel.transpose(1, 2), z_p


def acoustic_forward(W, cfg, batch, mel2word=None, z=None):
    """PortaSpeech_dict.forward(infer=True) (modules/dict_tts/model.py:36-62) as one function."""
    ret = {}
    wt = batch["word_tokens"]
    enc, dict_attn, pron_attn, _ = text_encode(W, cfg, wt, batch.get("pron_modified"), batch["keys"],
                                               batch["values"], batch["key_map"], batch["pinyin"], batch["pinyin_map"])
    ret.update(word_encoder_out=enc, dict_attn=dict_attn, pron_attn=pron_attn)
    dur_input = enc * (wt != 0).float().unsqueeze(-1)
    dur, src_padding = duration_predictor(W, cfg, dur_input)
    ret["dur"] = dur
    if mel2word is None:
        mel2word = length_regulate(durations_to_int(dur), (1 - src_padding.long()).sum(-1), 1)
    if mel2word.shape[1] % cfg.frames_multiple:
        extra = cfg.frames_multiple - mel2word.shape[1] % cfg.frames_multiple
        mel2word = torch.cat([mel2word] + [mel2word[:, -1:]] * extra, dim=1)
    ret["mel2word"] = mel2word
    x, nonpad = expand_by_mel2word(enc, mel2word)
    ret["decoder_inp"], ret["x_mask"] = x, nonpad
    if z is None:
        z = torch.distributions.Normal(0, 1).sample([x.shape[0], cfg.latent, x.shape[1] // cfg.frames_multiple])
    ret["mel_out"], ret["z_p"] = decode_mel(W, cfg, x, z)
    ret["mel_out_fvae"] = ret["mel_out"]
    return ret


# ---------------------------------------------------------------------------------------------
# HiFi-GAN V1 generator (modules/hifigan/hifigan.py:27-58,101-142)
# ---------------------------------------------------------------------------------------------

def hifigan_forward(W, vcfg, mel):
    """mel [B,T,80] -> wav [B, T*hop].  spec2wav transposes to [B,80,T] first (vocoders/hifigan.py:57-58)."""
    x = F.conv1d(mel.transpose(1, 2), W["conv_pre.weight"], W["conv_pre.bias"], padding=3)
    nk = len(vcfg.rb_kernels)
    for i, (u, k) in enumerate(zip(vcfg.up_rates, vcfg.up_kernels)):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, W[f"ups.{i}.weight"], W[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (kr, dils) in enumerate(zip(vcfg.rb_kernels, vcfg.rb_dilations)):
            r = f"resblocks.{i * nk + j}"
            y = x
            for m, d in enumerate(dils):
                t = F.conv1d(F.leaky_relu(y, 0.1), W[f"{r}.convs1.{m}.weight"], W[f"{r}.convs1.{m}.bias"],
                             dilation=d, padding=(kr * d - d) // 2)
                t = F.conv1d(F.leaky_relu(t, 0.1), W[f"{r}.convs2.{m}.weight"], W[f"{r}.convs2.{m}.bias"],
                             padding=(kr - 1) // 2)
                y = t + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                        # default slope 0.01 (hifigan.py:138)
    x = torch.tanh(F.conv1d(x, W["conv_post.weight"], W["conv_post.bias"], padding=3))
    return x.squeeze(1)
