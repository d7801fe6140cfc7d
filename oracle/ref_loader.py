"""TEST INFRASTRUCTURE ONLY -- loader shim for the upstream reference (read-only at /root/reference).

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to (1) validate oracle/dtts_oracle.py against the real reference modules
and (2) generate the committed fixtures under tests/golden/.

Recipe follows SURVEY.md Appendix A: two empty stub modules (chardet, librosa) are enough to
import modules.dict_tts.model.PortaSpeech_dict and modules.hifigan.hifigan.HifiGanGenerator.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """$DTTS_REFERENCE_ROOT, then /root/reference (the build container), then baseline/_ref (a driver-installed copy on
    the GPU pod) -- the first that holds the model code."""
    cands = [os.environ.get("DTTS_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "modules", "dict_tts")):
            return c
    return cands[1]


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "modules", "dict_tts"))


_loaded = {}


def load():
    """Returns dict(model_cls, hifigan_cls, hparams, TokenTextEncoder). cwd is restored afterwards."""
    if _loaded:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    for name in ("chardet", "librosa"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    try:
        from utils.hparams import set_hparams, hparams  # noqa
        set_hparams(config="egs/datasets/audio/biaobei/dict_tts.yaml", exp_name="",
                    hparams_str="use_word_input=True,word_size=8000,use_dict=True", print_hparams=False)
        from utils.text_encoder import TokenTextEncoder
        from modules.dict_tts.model import PortaSpeech_dict
        from modules.hifigan.hifigan import HifiGanGenerator
        import yaml
        with open("egs/egs_bases/tts/vocoder/hifigan.yaml") as f:
            voc_cfg = yaml.safe_load(f)
    finally:
        os.chdir(cwd)
    _loaded.update(model_cls=PortaSpeech_dict, hifigan_cls=HifiGanGenerator, hparams=hparams,
                   TokenTextEncoder=TokenTextEncoder, voc_cfg=voc_cfg)
    return _loaded


def build_models(acoustic_sd, vocoder_sd):
    """The UNMODIFIED reference modules holding the given checkpoints, in the state the reference's own inference
    leaves them in: eval mode, weight-norm removed (tasks/tts/ps_flow.py:257-268, vocoders/hifigan.py:16-32).
    Returns (PortaSpeech_dict, HifiGanGenerator)."""
    import torch
    R = load()
    enc = R["TokenTextEncoder"](None, vocab_list=["a", "b", "c"], replace_oov="<UNK>")
    model = R["model_cls"](enc).eval()
    missing, unexpected = model.load_state_dict(acoustic_sd, strict=False)
    assert not unexpected, unexpected
    assert all(m.startswith("fvae.encoder.") for m in missing), missing

    def _rm(m):
        try:
            torch.nn.utils.remove_weight_norm(m)
        except ValueError:
            pass
    model.apply(_rm)
    voc = R["hifigan_cls"](R["voc_cfg"]).eval()
    voc.load_state_dict(vocoder_sd, strict=True)
    voc.remove_weight_norm()
    return model, voc
