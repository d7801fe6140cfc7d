"""YAML config chains + ``--hparams`` overrides, compatible with the reference's loader so its ``egs/**.yaml``
files and the ``config.yaml`` saved beside a checkpoint load unchanged.

Behaviour mirrored from utils/hparams.py:25-126 of the reference:
  * ``base_config`` (string or list) is resolved depth-first; entries starting with '.' are relative to the file
    that names them, the rest are relative to the working directory (the reference repo root);
  * nested dicts are merged key by key, everything else is replaced (``override_config``, :17-22);
  * ``checkpoints/<exp_name>/config.yaml`` -- if present -- overrides the chain unless ``--reset``;
  * ``--hparams "a=1,b.c=2,d=[1 1 1]"`` casts each value to the type of the existing entry (bool/list/dict
    values are evaluated as Python literals, lists accept blanks as separators);
  * the resolved dict is stored in the module-level ``hparams`` (the reference's global) and returned.
Only inference is supported here, so nothing is ever written to the work dir.
"""
import argparse
import ast
import os
from typing import Dict, Optional

import yaml

hparams: Dict = {}


def override_config(old: dict, new: dict) -> None:
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(old.get(k), dict):
            override_config(old[k], v)
        else:
            old[k] = v


def _load_chain(path: str, seen: set, chain: list, root: Optional[str]) -> dict:
    full = path if os.path.isabs(path) or root is None else os.path.join(root, path)
    if not os.path.exists(full):
        return {}
    with open(full) as f:
        cfg = yaml.safe_load(f) or {}
    seen.add(path)
    merged: dict = {}
    bases = cfg.get("base_config", [])
    if not isinstance(bases, list):
        bases = [bases]
    for b in bases:
        if b.startswith("."):
            b = os.path.normpath(os.path.join(os.path.dirname(path), b))
        if b not in seen:
            override_config(merged, _load_chain(b, seen, chain, root))
    override_config(merged, cfg)
    chain.append(path)
    return merged


def _cast(old, text: str):
    text = text.strip("'\" ")
    if text in ("True", "False") or isinstance(old, (bool, list, dict)):
        if isinstance(old, list):
            text = text.replace(" ", ",")
        return ast.literal_eval(text)
    if old is None:
        try:
            return ast.literal_eval(text)
        except (ValueError, SyntaxError):
            return text
    return type(old)(text)


def apply_overrides(cfg: dict, spec: str) -> None:
    """``a=1,b.c=2`` (commas inside [...] belong to the value)."""
    if not spec:
        return
    parts, depth, cur = [], 0, ""
    for ch in spec:
        depth += ch in "[({"
        depth -= ch in "])}"
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    for item in parts:
        if not item.strip():
            continue
        key, value = item.split("=", 1)
        node = cfg
        path = key.strip().split(".")
        for k in path[:-1]:
            node = node[k]
        node[path[-1]] = _cast(node.get(path[-1]), value)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="Dict-TTS B200 inference (tasks/run.py-compatible flags)")
    ap.add_argument("--config", type=str, default="")
    ap.add_argument("--exp_name", type=str, default="")
    ap.add_argument("--hparams", type=str, default="")
    ap.add_argument("--infer", action="store_true")
    ap.add_argument("--validate", action="store_true")
    ap.add_argument("--reset", action="store_true")
    ap.add_argument("--remove", action="store_true")
    ap.add_argument("--debug", action="store_true")
    args, _ = ap.parse_known_args(argv)
    return args


def set_hparams(config: str = "", exp_name: str = "", hparams_str: str = "", print_hparams: bool = False,
                global_hparams: bool = True, root: Optional[str] = None, reset: bool = False, infer: bool = True,
                argv=None) -> dict:
    """Same call shapes as the reference: no arguments -> parse sys.argv; otherwise explicit values.
    ``root``: directory the repo-relative YAML paths and ``checkpoints/`` resolve against (default: cwd)."""
    debug = validate = False
    if config == "" and exp_name == "":
        a = parse_args(argv)
        config, exp_name, hparams_str = a.config, a.exp_name, a.hparams
        reset, infer, debug, validate = a.reset, a.infer, a.debug, a.validate
    if config == "" and exp_name == "":
        raise ValueError("either --config or --exp_name is required")
    chain: list = []
    saved = {}
    work_dir = ""
    if exp_name:
        work_dir = f"checkpoints/{exp_name}"
        saved_path = os.path.join(root or "", work_dir, "config.yaml")
        if os.path.exists(saved_path):
            with open(saved_path) as f:
                saved = yaml.safe_load(f) or {}
    cfg: dict = {}
    if config:
        cfg.update(_load_chain(config, set(), chain, root))
    if not reset:
        cfg.update(saved)
    # the reference always runs from its repo root; with an explicit root the checkpoint directory follows it
    cfg["work_dir"] = os.path.join(root, work_dir) if (root and work_dir) else work_dir
    apply_overrides(cfg, hparams_str)
    cfg.update(infer=infer, debug=debug, validate=validate, exp_name=exp_name)
    if global_hparams:
        hparams.clear()
        hparams.update(cfg)
    if print_hparams:
        print("| Hparams chains: ", chain)
        for k, v in sorted(cfg.items()):
            print(f"{k}: {v}")
    return cfg
