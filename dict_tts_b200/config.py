"""Resolved hyper-parameters of the Biaobei ``dict_tts.yaml`` hot path.

Values mirror the effective config of the reference (SURVEY.md header table; chain
egs/egs_bases/config_base.yaml -> tts/base.yaml -> tts/fs2.yaml -> tts/ps_flow.yaml -> tts/dict_tts.yaml,
egs/datasets/audio/biaobei/{base_text2mel,dict_tts}.yaml).  ``from_hparams`` rebuilds it from a loaded
hparams dict so a differently sized checkpoint/config still works.
"""
from dataclasses import dataclass, field
from typing import List


@dataclass
class AcousticConfig:
    hidden: int = 192            # hidden_size            (ps_flow.yaml:10)
    n_heads: int = 2             # num_heads              (tts/base.yaml:70)
    enc_layers: int = 4          # hard-coded 4+4         (dict_encoder.py:104-128)
    ffn_kernel: int = 5          # enc_ffn_kernel_size    (ps_flow.yaml:11)
    ffn_filter: int = 768        # 4*hidden               (dict_encoder.py:159-162)
    dict_dim: int = 768          # S2PA key/value size    (dict_encoder.py:18)
    word_size: int = 8000        # word vocab             (base_zh.yaml:5)
    pinyin_size: int = 185       # value_embedding_size   (biaobei/dict_tts.yaml:12)
    dur_layers: int = 3          # dur_predictor_layers   (ps_flow.yaml:17-22)
    dur_kernel: int = 5
    dur_chans: int = 128         # portaspeech/model.py:164-169
    frames_multiple: int = 4     # ps_flow.yaml:61
    latent: int = 16             # latent_size
    dec_layers: int = 4          # fvae_dec_n_layers
    dec_kernel: int = 5          # fvae_kernel_size
    flow_hidden: int = 64        # prior_glow_hidden
    flow_kernel: int = 3         # glow_kernel_size
    flow_blocks: int = 4         # prior_glow_n_blocks (number of coupling layers)
    flow_layers: int = 4         # WN layers per coupling layer (fvae_semantics.py:77-79)
    n_mel: int = 80              # audio_num_mel_bins
    language_zh: bool = True     # language == 'zh' -> add_pron_rule

    @staticmethod
    def from_hparams(hp) -> "AcousticConfig":
        return AcousticConfig(
            hidden=hp["hidden_size"], n_heads=hp["num_heads"], ffn_kernel=hp["enc_ffn_kernel_size"],
            ffn_filter=4 * hp["hidden_size"], word_size=hp["word_size"],
            pinyin_size=hp["value_embedding_size"], dur_layers=hp["dur_predictor_layers"],
            dur_kernel=hp["dur_predictor_kernel"], frames_multiple=hp["frames_multiple"],
            latent=hp["latent_size"], dec_layers=hp["fvae_dec_n_layers"], dec_kernel=hp["fvae_kernel_size"],
            flow_hidden=hp["prior_glow_hidden"], flow_kernel=hp["glow_kernel_size"],
            flow_blocks=hp["prior_glow_n_blocks"], n_mel=hp["audio_num_mel_bins"],
            language_zh=(hp.get("language", "zh") == "zh"))


@dataclass
class PortaSpeechConfig(AcousticConfig):
    """The non-dict sibling (egs/egs_bases/tts/ps_flow.yaml; modules/portaspeech/model.py:132-200): same predictor /
    FVAE sizes, plus the phoneme vocabulary, the word encoder depth and the relative-position window."""
    ph_size: int = 80            # len(phone dictionary): rows of ph_encoder.emb.weight
    word_enc_layers: int = 4     # word_enc_layers        (ps_flow.yaml)
    rel_window: int = 4          # TextEncoder(window_size=4) (portaspeech/model.py:79)

    @staticmethod
    def from_hparams(hp, ph_size: int = 80) -> "PortaSpeechConfig":
        return PortaSpeechConfig(
            hidden=hp["hidden_size"], n_heads=hp["num_heads"], enc_layers=hp["enc_layers"],
            ffn_kernel=hp["enc_ffn_kernel_size"], ffn_filter=4 * hp["hidden_size"],
            dur_layers=hp["dur_predictor_layers"], dur_kernel=hp["dur_predictor_kernel"],
            frames_multiple=hp["frames_multiple"], latent=hp["latent_size"], dec_layers=hp["fvae_dec_n_layers"],
            dec_kernel=hp["fvae_kernel_size"], flow_hidden=hp["prior_glow_hidden"], flow_kernel=hp["glow_kernel_size"],
            flow_blocks=hp["prior_glow_n_blocks"], n_mel=hp["audio_num_mel_bins"], ph_size=ph_size,
            word_enc_layers=hp["word_enc_layers"])


@dataclass
class VocoderConfig:
    """HiFi-GAN V1 generator (egs/egs_bases/tts/vocoder/hifigan.yaml:3-10)."""
    n_mel: int = 80
    init_ch: int = 512
    up_rates: List[int] = field(default_factory=lambda: [8, 8, 2, 2])
    up_kernels: List[int] = field(default_factory=lambda: [16, 16, 4, 4])
    rb_kernels: List[int] = field(default_factory=lambda: [3, 7, 11])
    rb_dilations: List[List[int]] = field(default_factory=lambda: [[1, 3, 5], [1, 3, 5], [1, 3, 5]])

    @property
    def hop(self) -> int:
        h = 1
        for u in self.up_rates:
            h *= u
        return h

    @staticmethod
    def from_dict(h) -> "VocoderConfig":
        assert str(h.get("resblock", "1")) == "1", "only ResBlock1 (HiFi-GAN V1) is on the hot path"
        return VocoderConfig(init_ch=h["upsample_initial_channel"], up_rates=list(h["upsample_rates"]),
                             up_kernels=list(h["upsample_kernel_sizes"]),
                             rb_kernels=list(h["resblock_kernel_sizes"]),
                             rb_dilations=[list(d) for d in h["resblock_dilation_sizes"]])


SAMPLE_RATE = 22050
HOP_SIZE = 256
