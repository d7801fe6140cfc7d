"""Fabricates a complete experiment tree in the reference's on-disk formats (there is no network: no real
checkpoints or binarized Biaobei data), so the ``--infer`` entry point and the loaders can be exercised end to end:

    <root>/checkpoints/<exp>/model_ckpt_steps_<N>.ckpt + config.yaml     utils/trainer.py:436-449
    <root>/checkpoints/<voc>/model_ckpt_steps_<M>.ckpt + config.yaml     vocoders/hifigan.py:16-52
    <root>/data/binary/fake/{test.data,test.idx,test_lengths.npy,word_set.json,pinyin_encoder.pkl,dict_embed.*}
"""
import json
import os
import pickle

import numpy as np
import torch
import yaml

from dict_tts_b200 import synth
from dict_tts_b200.data import IndexedDatasetBuilder

HPARAMS = dict(
    task_cls="tasks.tts.dict_tts.DictTTSTask", vocoder="HifiGAN", hidden_size=192, num_heads=2,
    enc_ffn_kernel_size=5, word_size=8000, value_embedding_size=185, dur_predictor_layers=3, dur_predictor_kernel=5,
    frames_multiple=4, latent_size=16, fvae_dec_n_layers=4, fvae_kernel_size=5, prior_glow_hidden=64,
    glow_kernel_size=3, prior_glow_n_blocks=4, audio_num_mel_bins=80, language="zh", hop_size=256,
    audio_sample_rate=22050, max_frames=1548, min_frames=0, num_test_samples=0, test_ids=[], use_dict=True,
    use_word_input=True, two_stage=True, profile_infer=False, out_wav_norm=False, gen_dir_name="",
    max_valid_sentences=1, test_set_name="test", seed=1234)

VOC_CONFIG = dict(resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
                  upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
                  resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], audio_num_mel_bins=80)


def write(root: str, exp: str = "fake_dict_tts", n_items: int = 5, n_vocab: int = 40, seed: int = 7,
          steps=(1000, 3000, 2000), original_hifigan_layout: bool = False, same_key_value: bool = False) -> dict:
    g = torch.Generator().manual_seed(seed)
    ck = os.path.join(root, "checkpoints", exp)
    vk = os.path.join(root, "checkpoints", "fake_hifigan")
    bd = os.path.join(root, "data", "binary", "fake")
    for d in (ck, vk, bd):
        os.makedirs(d, exist_ok=True)
    # ---- acoustic checkpoints: the NEWEST step holds the real weights, older ones hold junk ----
    sd = synth.make_acoustic_state_dict(1234)
    for st in steps:
        model = sd if st == max(steps) else {k: torch.zeros_like(v) for k, v in sd.items()}
        torch.save({"epoch": 1, "global_step": st, "checkpoint_callback_best": np.float64(0.5),
                    "optimizer_states": [], "state_dict": {"model": model, "mel_disc": {"dummy.weight": torch.ones(3)}}},
                   os.path.join(ck, f"model_ckpt_steps_{st}.ckpt"), _use_new_zipfile_serialization=False)
    hp = dict(HPARAMS, binary_data_dir=os.path.join(root, "data", "binary", "fake"), vocoder_ckpt=vk)
    with open(os.path.join(ck, "config.yaml"), "w") as f:
        yaml.safe_dump(hp, f)
    # ---- vocoder checkpoint ----
    vsd = synth.make_vocoder_state_dict(4321)
    if original_hifigan_layout:
        with open(os.path.join(vk, "config.json"), "w") as f:
            json.dump(VOC_CONFIG, f)
        torch.save({"generator": vsd}, os.path.join(vk, "generator_v1"))
    else:
        with open(os.path.join(vk, "config.yaml"), "w") as f:
            yaml.safe_dump(VOC_CONFIG, f)
        torch.save({"state_dict": {"model_gen": vsd, "model_disc": {}}}, os.path.join(vk, "model_ckpt_steps_500.ckpt"))
        torch.save({"state_dict": {"model_gen": vsd, "model_disc": {}}}, os.path.join(vk, "model_ckpt_steps_1200.ckpt"))
    # ---- dictionary + vocabulary ----
    vocab = [chr(0x4E00 + i) for i in range(n_vocab)]
    with open(os.path.join(bd, "word_set.json"), "w") as f:
        json.dump(["<BOS>", "<EOS>"] + vocab, f, ensure_ascii=False)
    pinyin_tokens = ["<pad>"] + [f"py{i}" for i in range(1, 185)]
    with open(os.path.join(bd, "pinyin_encoder.pkl"), "wb") as f:
        pickle.dump(pinyin_tokens, f)
    db = IndexedDatasetBuilder(os.path.join(bd, "dict_embed"))
    n_ids = 3 + 2 + n_vocab                                   # reserved + <BOS>,<EOS> + characters
    for wid in range(n_ids):
        npron = 1 + int(torch.randint(0, 3, (1,), generator=g))
        key_map, pin, pmap = [], [], []
        for i in range(npron):
            ln = int(torch.randint(4, 12, (1,), generator=g)) + 2
            key_map += [0] + [i + 1] * (ln - 2) + [0]
            a = int(torch.randint(1, 185, (1,), generator=g))
            b = int(torch.randint(1, 185, (1,), generator=g))
            pin += [pinyin_tokens[a], pinyin_tokens[b]]
            pmap += [i + 1, i + 1]
        feats = (torch.randn(len(key_map), 768, generator=g) * 0.5).numpy()
        # same_key_value: ONE array as key and value, as the reference binarizer writes it (binarizer_zh.py:231-233)
        db.add_item({"key": feats, "value": feats if same_key_value else feats.copy(), "key_map": key_map, "tokens_gloss": [], "pinyin": pin,
                     "pinyin_map": pmap})
    db.finalize()
    # ---- test items ----
    ib = IndexedDatasetBuilder(os.path.join(bd, "test"))
    lengths = []
    word_to_id = {w: i for i, w in enumerate(["<pad>", "<EOS>", "<UNK>", "<BOS>"] + vocab)}
    word_to_id["<EOS>"] = 1
    for it in range(n_items):
        n = 3 + int(torch.randint(0, 6, (1,), generator=g))
        chars = [vocab[int(torch.randint(0, n_vocab, (1,), generator=g))] for _ in range(n)]
        words = ["<BOS>"] + chars + ["<EOS>"]
        tokens = [word_to_id[w] for w in words]
        dur = torch.randint(3, 9, (len(words),), generator=g)
        T = int(dur.sum()) // 4 * 4
        mel2word = torch.repeat_interleave(torch.arange(1, len(words) + 1), dur)[:T]
        lengths.append(T)
        ib.add_item({"item_name": f"fake_{it:03d}", "txt": "".join(chars), "words": words, "ph_words": words,
                     "word_tokens": tokens, "phone": tokens, "mel": np.zeros((T, 80), np.float32),
                     "mel2word": mel2word.numpy(), "ph2word": list(range(1, len(words) + 1)),
                     "pron_modified": [0] * len(words)})
    ib.finalize()
    np.save(os.path.join(bd, "test_lengths.npy"), np.array(lengths))
    return dict(root=root, exp=exp, work_dir=ck, vocoder_dir=vk, binary_dir=bd, n_items=n_items, hparams=hp)


PS_HPARAMS = dict(
    task_cls="tasks.tts.ps_flow.PortaSpeechFlowTask", vocoder="HifiGAN", hidden_size=192, num_heads=2, enc_layers=4,
    word_enc_layers=4, enc_ffn_kernel_size=5, dur_predictor_layers=3, dur_predictor_kernel=5, frames_multiple=4,
    latent_size=16, fvae_dec_n_layers=4, fvae_kernel_size=5, prior_glow_hidden=64, glow_kernel_size=3,
    prior_glow_n_blocks=4, audio_num_mel_bins=80, hop_size=256, audio_sample_rate=22050, max_frames=1548, min_frames=0,
    num_test_samples=0, test_ids=[], two_stage=True, profile_infer=False, out_wav_norm=False, gen_dir_name="",
    max_valid_sentences=1, test_set_name="test", seed=1234, use_post_glow=False, dur_level="word", max_input_tokens=1550)


def write_ps(root: str, exp: str = "fake_ps", n_items: int = 5, ph_size: int = 80, seed: int = 9) -> dict:
    """A PortaSpeech (non-dict) experiment: checkpoints/<exp>/{model_ckpt_steps_*.ckpt, config.yaml}, the vocoder
    checkpoint of write(), data/binary/fake_ps/{test.data, test.idx, test_lengths.npy, phone_set.json}."""
    from dict_tts_b200.config import PortaSpeechConfig
    g = torch.Generator().manual_seed(seed)
    ck = os.path.join(root, "checkpoints", exp)
    vk = os.path.join(root, "checkpoints", "fake_hifigan")
    bd = os.path.join(root, "data", "binary", "fake_ps")
    for d in (ck, vk, bd):
        os.makedirs(d, exist_ok=True)
    sd = synth.make_ps_state_dict(2468, PortaSpeechConfig(ph_size=ph_size))
    for st in (500, 2500):
        model = sd if st == 2500 else {k: torch.zeros_like(v) for k, v in sd.items()}
        torch.save({"epoch": 1, "global_step": st, "optimizer_states": [], "state_dict": {"model": model}},
                   os.path.join(ck, f"model_ckpt_steps_{st}.ckpt"), _use_new_zipfile_serialization=False)
    hp = dict(PS_HPARAMS, binary_data_dir=bd, vocoder_ckpt=vk)
    with open(os.path.join(ck, "config.yaml"), "w") as f:
        yaml.safe_dump(hp, f)
    if not os.path.exists(os.path.join(vk, "config.yaml")):
        with open(os.path.join(vk, "config.yaml"), "w") as f:
            yaml.safe_dump(VOC_CONFIG, f)
        torch.save({"state_dict": {"model_gen": synth.make_vocoder_state_dict(4321), "model_disc": {}}},
                   os.path.join(vk, "model_ckpt_steps_1200.ckpt"))
    phones = [f"ph{i}" for i in range(ph_size - 3)]
    with open(os.path.join(bd, "phone_set.json"), "w") as f:
        json.dump(phones, f)
    ib = IndexedDatasetBuilder(os.path.join(bd, "test"))
    lengths = []
    for it in range(n_items):
        n_words = 3 + int(torch.randint(0, 5, (1,), generator=g))
        per = torch.randint(1, 4, (n_words,), generator=g)
        ph2word = torch.repeat_interleave(torch.arange(1, n_words + 1), per)
        phone = torch.randint(3, ph_size, (int(per.sum()),), generator=g)
        dur = torch.randint(3, 9, (n_words,), generator=g)
        T = int(dur.sum()) // 4 * 4
        mel2word = torch.repeat_interleave(torch.arange(1, n_words + 1), dur)[:T]
        lengths.append(T)
        words = [f"w{j}" for j in range(n_words)]
        ib.add_item({"item_name": f"ps_{it:03d}", "txt": " ".join(words), "words": words, "ph_words": words,
                     "word_tokens": list(range(3, 3 + n_words)), "phone": phone.tolist(), "ph2word": ph2word.tolist(),
                     "mel": np.zeros((T, 80), np.float32), "mel2word": mel2word.numpy()})
    ib.finalize()
    np.save(os.path.join(bd, "test_lengths.npy"), np.array(lengths))
    return dict(root=root, exp=exp, work_dir=ck, vocoder_dir=vk, binary_dir=bd, n_items=n_items, hparams=hp)
