"""Checkpoint -> device weight arena (host side; the reference does this in test_start / load_model).

* ``fold_weight_norm`` reproduces ``remove_weight_norm`` (tasks/tts/ps_flow.py:257-268,
  vocoders/hifigan.py:27, modules/hifigan/hifigan.py:144-151): w = g * v / ||v||, norm over all dims but 0.
* ``pack_arena`` lays the live fp32 tensors out in one contiguous buffer (64-byte aligned entries) plus a
  (name, offset, numel) table -- the arena is what rank 0 broadcasts over NCCL at load.
* ``load_acoustic_checkpoint`` / ``load_vocoder_checkpoint`` read the reference's on-disk formats
  (utils/trainer.py:436-449, utils/ckpt_utils.py:8-16, vocoders/hifigan.py:16-52).
"""
import glob
import json
import os
import re
from typing import Dict, List, Tuple

import torch

# prefixes that exist in reference checkpoints but are never touched by the inference path
DEAD_PREFIXES = ("enc_pos_proj.", "dec_query_proj.", "dec_res_proj.", "attn.", "sin_pos.", "fvae.encoder.",
                 "dict_encoder.S2PA_module.emb.", "mel_disc.", "ph_encoder.", "word_encoder.")


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for name, t in sd.items():
        if name.endswith(".weight_g"):
            base = name[:-len("_g")]
            v = sd[base + "_v"]
            out[base] = torch._weight_norm(v.float(), t.float(), 0).contiguous()
        elif name.endswith(".weight_v"):
            continue
        else:
            out[name] = t.float().contiguous() if t.is_floating_point() else t
    return out


def drop_dead(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in sd.items() if not k.startswith(DEAD_PREFIXES)}


def pack_arena(sd: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, List[Tuple[str, int, int]]]:
    """Returns (flat fp32 CPU tensor, [(name, offset_in_floats, numel)]).  Entries are 16-float aligned."""
    table = []
    off = 0
    for name in sorted(sd):
        t = sd[name]
        if not t.is_floating_point():
            continue
        table.append((name, off, t.numel()))
        off += (t.numel() + 15) // 16 * 16
    arena = torch.zeros(off, dtype=torch.float32)
    for name, o, n in table:
        arena[o:o + n] = sd[name].reshape(-1).float()
    return arena, table


def get_last_checkpoint(work_dir: str):
    """Newest ``model_ckpt_steps_<N>.ckpt`` (utils/ckpt_utils.py:8-25)."""
    paths = glob.glob(os.path.join(work_dir, "model_ckpt_steps_*.ckpt"))
    if not paths:
        return None, None
    paths.sort(key=lambda p: -int(re.findall(r".*steps_(\d+)\.ckpt", p)[0]))
    ckpt = torch.load(paths[0], map_location="cpu", weights_only=False)   # holds numpy scalars (trainer.py:436-449)
    return ckpt, paths[0]


def load_acoustic_checkpoint(work_dir: str, with_step: bool = False):
    """Newest checkpoint of an experiment -> live, weight-norm-folded ``model`` tensors (the ``mel_disc`` child and the
    train-only parameters are dropped).  ``with_step``: also return ``global_step`` (names the output directory)."""
    ckpt, path = get_last_checkpoint(work_dir)
    if ckpt is None:
        raise FileNotFoundError(f"no model_ckpt_steps_*.ckpt under {work_dir}")
    sd = drop_dead(fold_weight_norm(ckpt["state_dict"]["model"]))
    return (sd, int(ckpt.get("global_step", 0))) if with_step else sd


def load_vocoder_checkpoint(base_dir: str):
    """Both layouts HifiGAN.__init__ accepts (vocoders/hifigan.py:40-52). Returns (folded state, config dict)."""
    cfg_yaml = os.path.join(base_dir, "config.yaml")
    if os.path.exists(cfg_yaml):
        paths = sorted(glob.glob(os.path.join(base_dir, "model_ckpt_steps_*.ckpt")),
                       key=lambda p: int(re.findall(r"model_ckpt_steps_(\d+)\.ckpt", p)[0]))
        from .hparams import set_hparams
        config = set_hparams(cfg_yaml, global_hparams=False, print_hparams=False)
        state = torch.load(paths[-1], map_location="cpu", weights_only=False)["state_dict"]["model_gen"]
    else:
        with open(os.path.join(base_dir, "config.json")) as f:
            config = json.load(f)
        state = torch.load(os.path.join(base_dir, "generator_v1"), map_location="cpu", weights_only=False)["generator"]
    return fold_weight_norm(state), config
