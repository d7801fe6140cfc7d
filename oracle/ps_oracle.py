"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch, fp32) of the PortaSpeech (non-dict) sibling's inference
forward, SURVEY.md §8f-3: the oracle for the NEXT row of the scope table.  No product code runs this path yet.

Follows (relative to /root/reference):
  modules/portaspeech/model.py:69-129    TextEncoder (phoneme encoder: embedding, ConvReluNorm pre-net, post-LN encoder)
  modules/portaspeech/glow_modules.py:40-72   ConvReluNorm
  modules/commons/rel_transformer_encoder.py:26-247   Encoder / MultiHeadAttention with relative positions (window 4) / FFN
  modules/portaspeech/utils.py:3-16      group_hidden_by_segs
  modules/fastspeech/tts_modules.py:458-566, modules/commons/common_layers.py:93-148,624-673   FFT-block word encoder
  modules/portaspeech/model.py:239-366   PortaSpeech.forward / run_text_encoder / attention / add_dur / position embeddings
  modules/portaspeech/fvae.py:62-112     FVAE (infer): same g_pre_net / prior flow / decoder as the dict model's

Scope notes pinned by running the reference here (oracle/make_golden_ps.py):
  * `modules/glow` is NOT part of the reference checkout, so `use_post_glow: true` (the shipped ps_flow.yaml) cannot even
    construct the model; the runnable inference path is `use_post_glow=False`, i.e. mel_out = mel_out_fvae.
  * the encoder is post-LN (pre_ln=False), its FFN activation is ReLU, attention mask value -1e4, LayerNorm eps 1e-4;
    the word encoder's FFN has kernel size 1 (third positional argument of FastspeechDecoder) and exact GELU.
"""
import math

import torch
import torch.nn.functional as F

from oracle import dtts_oracle as O


# ---------------------------------------------------------------------------------------------
# phoneme encoder
# ---------------------------------------------------------------------------------------------
def conv_relu_norm(W, p, x, x_mask, n_layers=3, kernel=5):
    """ConvReluNorm.forward (glow_modules.py:65-72): n x {conv k5 on x*mask, channel LN, ReLU}, 1x1 proj, residual."""
    x_org = x
    for i in range(n_layers):
        x = F.conv1d(x * x_mask, W[f"{p}.conv_layers.{i}.weight"], W[f"{p}.conv_layers.{i}.bias"], padding=kernel // 2)
        x = O.channel_layer_norm(x, W[f"{p}.norm_layers.{i}.gamma"], W[f"{p}.norm_layers.{i}.beta"])
        x = torch.relu(x)
    x = x_org + F.conv1d(x, W[f"{p}.proj.weight"], W[f"{p}.proj.bias"])
    return x * x_mask


def rel_self_attention(W, p, x, attn_mask, n_heads, window):
    """MultiHeadAttention.forward/attention with window_size (rel_transformer_encoder.py:117-158).  The reference's
    pad-and-reshape trick adds q . emb_rel_k[s - t + w] to the score of (t, s) when |s - t| <= w (nothing otherwise) and
    sum_s p[t, s] * emb_rel_v[s - t + w] to the output; written here as an explicit gather."""
    q = F.conv1d(x, W[p + ".conv_q.weight"], W[p + ".conv_q.bias"])
    k = F.conv1d(x, W[p + ".conv_k.weight"], W[p + ".conv_k.bias"])
    v = F.conv1d(x, W[p + ".conv_v.weight"], W[p + ".conv_v.bias"])
    b, d, t = q.shape
    dk = d // n_heads
    q = q.view(b, n_heads, dk, t).transpose(2, 3)
    k = k.view(b, n_heads, dk, t).transpose(2, 3)
    v = v.view(b, n_heads, dk, t).transpose(2, 3)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
    pos = torch.arange(t, device=x.device)
    rel = pos[None, :] - pos[:, None]                                   # s - t
    inside = rel.abs() <= window
    idx = (rel + window).clamp(0, 2 * window)
    ek = W[p + ".emb_rel_k"][0][idx] * inside[..., None]                # [t, s, dk] (heads share the embeddings)
    ev = W[p + ".emb_rel_v"][0][idx] * inside[..., None]
    scores = scores + torch.einsum("bhtd,tsd->bhts", q, ek) / math.sqrt(dk)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    pr = F.softmax(scores, dim=-1)
    out = torch.matmul(pr, v) + torch.einsum("bhts,tsd->bhtd", pr, ev)
    out = out.transpose(2, 3).contiguous().view(b, d, t)
    return F.conv1d(out, W[p + ".conv_o.weight"], W[p + ".conv_o.bias"])


def rel_encoder(W, p, x, x_mask, n_layers, n_heads, kernel, window):
    """Encoder.forward with pre_ln=False (rel_transformer_encoder.py:55-79): post-LN residual blocks."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    for i in range(n_layers):
        x = x * x_mask
        y = rel_self_attention(W, f"{p}.attn_layers.{i}", x, attn_mask, n_heads, window)
        x = O.channel_layer_norm(x + y, W[f"{p}.norm_layers_1.{i}.gamma"], W[f"{p}.norm_layers_1.{i}.beta"])
        f = F.conv1d(x * x_mask, W[f"{p}.ffn_layers.{i}.conv_1.weight"], W[f"{p}.ffn_layers.{i}.conv_1.bias"],
                     padding=kernel // 2)
        f = torch.relu(f)                                               # FFN activation None -> ReLU (:252-255)
        f = F.conv1d(f * x_mask, W[f"{p}.ffn_layers.{i}.conv_2.weight"], W[f"{p}.ffn_layers.{i}.conv_2.bias"]) * x_mask
        x = O.channel_layer_norm(x + f, W[f"{p}.norm_layers_2.{i}.gamma"], W[f"{p}.norm_layers_2.{i}.beta"])
    return x * x_mask


def ph_encode(W, cfg, txt_tokens, window=4):
    """TextEncoder.forward (portaspeech/model.py:119-129) -> [B, T_ph, H] (before the caller's * src_nonpadding)."""
    H = cfg.hidden
    lengths = (txt_tokens > 0).long().sum(-1)
    x = F.embedding(txt_tokens, W["ph_encoder.emb.weight"]) * math.sqrt(H)
    x = x.transpose(1, 2)
    T = x.shape[2]
    x_mask = (torch.arange(T, device=x.device)[None] < lengths[:, None]).unsqueeze(1).to(x.dtype)
    x = conv_relu_norm(W, "ph_encoder.pre", x, x_mask)
    x = rel_encoder(W, "ph_encoder.encoder", x, x_mask, cfg.enc_layers, cfg.n_heads, cfg.ffn_kernel, window)
    return x.transpose(1, 2)


# ---------------------------------------------------------------------------------------------
# word level
# ---------------------------------------------------------------------------------------------
def group_hidden_by_segs(h, seg_ids, max_len):
    """Mean of the phoneme vectors of every word (portaspeech/utils.py:3-16); segment 0 (padding) is dropped."""
    B, T, H = h.shape
    s = h.new_zeros(B, max_len + 1, H).scatter_add_(1, seg_ids[:, :, None].expand(B, T, H), h)
    n = h.new_zeros(B, max_len + 1).scatter_add_(1, seg_ids, h.new_ones(B, T))
    return s[:, 1:] / n[:, 1:, None].clamp(min=1)


def sinusoidal_table(n, dim):
    """SinusoidalPositionalEmbedding.get_embedding (common_layers.py:110-127), padding_idx = 0."""
    half = dim // 2
    e = torch.exp(torch.arange(half, dtype=torch.float) * -(math.log(10000) / (half - 1)))
    e = torch.arange(n, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    e[0] = 0
    return e


def fft_blocks(W, p, x, n_layers, n_heads, kernel):
    """FFTBlocks.forward (tts_modules.py:493-518) with EncSALayer (common_layers.py:653-673): pre-LN blocks of
    nn.LayerNorm -> multi-head self attention (no biases, key padding mask) -> residual -> nn.LayerNorm -> conv FFN
    (k, * k^-1/2, GELU, linear) -> residual, every step re-masked; final nn.LayerNorm."""
    B, T, C = x.shape
    pad = x.abs().sum(-1).eq(0)
    keep = (1 - pad.float())[:, :, None]
    first = x[..., 0]                                                   # positions: make_positions on the first channel
    m = first.ne(0).int()
    positions = (torch.cumsum(m, dim=1) * m).long()
    x = x + W[p + ".pos_embed_alpha"] * sinusoidal_table(T + 1, C).to(x.device)[positions]
    x = x * keep
    for i in range(n_layers):
        q = f"{p}.layers.{i}.op"
        r = x
        h = F.layer_norm(x, (C,), W[q + ".layer_norm1.weight"], W[q + ".layer_norm1.bias"])
        h, _ = F.multi_head_attention_forward(
            h.transpose(0, 1), h.transpose(0, 1), h.transpose(0, 1), C, n_heads, W[q + ".self_attn.in_proj_weight"], None,
            None, None, False, 0.0, W[q + ".self_attn.out_proj.weight"], None, training=False, key_padding_mask=pad,
            need_weights=False)
        x = (r + h.transpose(0, 1)) * keep
        r = x
        h = F.layer_norm(x, (C,), W[q + ".layer_norm2.weight"], W[q + ".layer_norm2.bias"])
        h = F.conv1d(h.transpose(1, 2), W[q + ".ffn.ffn_1.weight"], W[q + ".ffn.ffn_1.bias"], padding=kernel // 2)
        h = F.gelu(h.transpose(1, 2) * kernel ** -0.5)
        h = F.linear(h, W[q + ".ffn.ffn_2.weight"], W[q + ".ffn.ffn_2.bias"])
        x = (r + h) * keep
    return F.layer_norm(x, (C,), W[p + ".layer_norm.weight"], W[p + ".layer_norm.bias"]) * keep


def sin_pos_emb(x, dim):
    """SinusoidalPosEmb.forward (portaspeech/model.py:22-33): continuous positions x [B, T] -> [B, T, dim]."""
    half = dim // 2
    e = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
    e = x[:, :, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def build_word_mask(x2word, y2word):
    return (x2word[:, :, None] == y2word[:, None, :]).long()


def build_pos_embed(word2word, x2word, dim):
    """Relative position of every phoneme / frame inside its word in (0, 1] -> sinusoids (model.py:359-363)."""
    m = build_word_mask(word2word, x2word).float()                      # [B, T_word, T_x]
    pos = (m.cumsum(-1) / m.sum(-1).clamp(min=1)[..., None] * m).sum(1)
    return sin_pos_emb(pos, dim)


def word_to_phoneme_attention(W, H, ph_encoder_out, enc_pos, word_encoder_out, dec_pos, mel2word, dec_word_mask):
    """PortaSpeech.attention (model.py:304-315): one head, no biases, frames only see the phonemes of their own word."""
    ph_kv = F.linear(torch.cat([ph_encoder_out, enc_pos], -1), W["enc_pos_proj.weight"], W["enc_pos_proj.bias"])
    expanded = F.pad(word_encoder_out, [0, 0, 1, 0]).gather(1, mel2word[:, :, None].expand(-1, -1, H))
    qin = torch.cat([expanded, dec_pos], -1)
    dec_q = F.linear(qin, W["dec_query_proj.weight"], W["dec_query_proj.bias"])
    x_res = F.linear(qin, W["dec_res_proj.weight"], W["dec_res_proj.bias"])
    Win = W["attn.in_proj_weight"]
    q = F.linear(dec_q, Win[:H]) * H ** -0.5                            # head_dim = H for one head
    k = F.linear(ph_kv, Win[H:2 * H])
    v = F.linear(ph_kv, Win[2 * H:])
    scores = torch.bmm(q, k.transpose(1, 2)) + (1 - dec_word_mask) * -1e9
    weight = F.softmax(scores, dim=-1)
    x = F.linear(torch.bmm(weight, v), W["attn.out_proj.weight"])
    return x + x_res, weight


# ---------------------------------------------------------------------------------------------
# whole forward
# ---------------------------------------------------------------------------------------------
def ps_forward(W, cfg, txt_tokens, ph2word, word_len, mel2word=None, z=None):
    """PortaSpeech.forward(infer=True, forward_post_glow=False) with dur_level = word (model.py:239-302)."""
    H = cfg.hidden
    ret = {}
    nonpad = (txt_tokens > 0).float()[:, :, None]
    ph = ph_encode(W, cfg, txt_tokens) * nonpad
    ret["ph_encoder_out"] = ph
    Tw = int(word_len)
    word = fft_blocks(W, "word_encoder", group_hidden_by_segs(ph, ph2word, Tw), getattr(cfg, "word_enc_layers", 4), cfg.n_heads, 1)
    ret["word_encoder_out"] = word
    # add_dur (model.py:317-340): phoneme-level prediction summed per word
    dur_ph, src_padding = O.duration_predictor(W, cfg, ph)
    dur = torch.zeros(ph.shape[0], Tw + 1).scatter_add(1, ph2word, dur_ph)[:, 1:]
    ret["dur"] = dur
    if mel2word is None:
        mel2word = O.length_regulate(O.durations_to_int(dur), (1 - src_padding.long()).sum(-1), 1)
    if mel2word.shape[1] % cfg.frames_multiple:
        extra = cfg.frames_multiple - mel2word.shape[1] % cfg.frames_multiple
        mel2word = torch.cat([mel2word] + [mel2word[:, -1:]] * extra, dim=1)
    ret["mel2word"] = mel2word
    tgt_nonpadding = (mel2word > 0).float()[:, :, None]
    word2word = torch.arange(Tw)[None, :] + 1
    enc_pos = build_pos_embed(word2word, ph2word, H)
    dec_pos = build_pos_embed(word2word, mel2word, H)
    x, weight = word_to_phoneme_attention(W, H, ph, enc_pos, word, dec_pos, mel2word,
                                          build_word_mask(mel2word, ph2word).float())
    ret["attn"] = weight
    x = x * tgt_nonpadding
    ret["x_mask"], ret["decoder_inp"] = tgt_nonpadding, x
    if z is None:
        z = torch.distributions.Normal(0, 1).sample([x.shape[0], cfg.latent, x.shape[1] // cfg.frames_multiple])
    ret["mel_out"], ret["z_p"] = O.decode_mel(W, cfg, x, z)
    ret["mel_out_fvae"] = ret["mel_out"]
    return ret
