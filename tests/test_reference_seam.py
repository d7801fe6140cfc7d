"""The `task_cls` / `vocoder` seams against the reference's OWN classes, checked in the build container
where /root/reference exists (skipped elsewhere: the GPU box has no reference tree, and this container has no GPU, so the
reference Trainer cannot be run around the CUDA engine anywhere in this setup -- DESIGN.md §8).  The reference task
module is imported under import stubs for the packages this image lacks (matplotlib, librosa, tensorboard, ...); nothing
of the reference is executed beyond class construction and signature inspection."""
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

_MISSING = ("chardet", "librosa", "matplotlib", "parselmouth", "pypinyin", "jieba", "textgrid", "resemblyzer",
            "pytorch_memlab", "tensorboard", "skimage", "pyloudnorm", "webrtcvad", "g2p_en", "pycwt", "praatio",
            "torchaudio", "tensorboardX", "libtmux", "setproctitle", "pyworld")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Stub(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Stub("called")


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in _MISSING or name == "torch.utils.tensorboard":
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Stub(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


@pytest.fixture(scope="module")
def reference_task():
    import numpy as np
    finder = _Finder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, ref_loader.REF_ROOT)
    cwd = os.getcwd()
    os.chdir(ref_loader.REF_ROOT)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                     # utils/trainer.py:71 (numpy 1.x spelling)
    try:
        import tasks.tts.dict_tts as ref_task
        import vocoders.base_vocoder as ref_voc
        import modules.dict_tts.model as ref_model
        yield ref_task, ref_voc, ref_model
    finally:
        os.chdir(cwd)
        sys.meta_path.remove(finder)
        sys.path.remove(ref_loader.REF_ROOT)


def test_plugin_task_is_a_subclass_of_the_reference_task(reference_task):
    ref_task, _, _ = reference_task
    from dict_tts_b200 import plugin
    cls = plugin.B200DictTTSTask                      # module __getattr__: built around the importable reference task
    assert issubclass(cls, ref_task.DictTTSTask) and cls is not ref_task.DictTTSTask
    assert cls.__name__ == "B200DictTTSTask"
    # only test_start is overridden: test_step / after_infer / test_end are the reference's own functions
    assert "test_start" in cls.__dict__
    for name in ("test_step", "after_infer", "test_end", "build_model"):
        assert name not in cls.__dict__ and getattr(cls, name) is getattr(ref_task.DictTTSTask, name)
    # tasks/run.py resolves the class exactly like this (run.py:6-11)
    pkg, name = "dict_tts_b200.plugin.B200DictTTSTask".rsplit(".", 1)
    assert getattr(importlib.import_module(pkg), name).__mro__[1] is ref_task.DictTTSTask


def test_engine_forward_accepts_the_reference_call(reference_task):
    """DictTTSTask.test_step calls self.model(...) with these keywords (dict_tts.py:183-196): the engine's forward must
    bind them all, in the positions the reference uses for the positional ones."""
    _, _, ref_model = reference_task
    from dict_tts_b200.engine import DictTTSEngine
    ref_sig = inspect.signature(ref_model.PortaSpeech_dict.forward)
    eng_sig = inspect.signature(DictTTSEngine.forward)
    ref_names = [p for p in ref_sig.parameters if p != "self"]
    eng_names = [p for p in eng_sig.parameters if p != "self"]
    assert eng_names[:len(ref_names)] == ref_names, (ref_names, eng_names)
    for n in ref_names:                                # same defaults where the reference has one
        rd, ed = ref_sig.parameters[n].default, eng_sig.parameters[n].default
        if rd is not inspect.Parameter.empty and n != "infer":
            assert ed == rd, n
    eng_sig.bind(None, ("w", "p"), None, (None, None, None), ph2word=None, word_len=3, dict_msg=(1, 2, 3, 4, 5), infer=True,
                 forward_post_glow=False, spk_embed=None, two_stage=True, mel2word=None)


def test_vocoder_plugin_matches_the_base_vocoder_contract(reference_task):
    _, ref_voc, _ = reference_task
    from dict_tts_b200.plugin import B200HifiGAN
    for name in ("spec2wav", "wav2spec"):
        assert callable(getattr(B200HifiGAN, name)) and hasattr(ref_voc.BaseVocoder, name)
    ref_params = list(inspect.signature(ref_voc.BaseVocoder.spec2wav).parameters)
    mine = list(inspect.signature(B200HifiGAN.spec2wav).parameters)
    assert mine[:2] == ref_params[:2] == ["self", "mel"]
    # get_vocoder_cls resolves dotted paths by import (vocoders/base_vocoder.py:15-23)
    assert ref_voc.get_vocoder_cls({"vocoder": "dict_tts_b200.plugin.B200HifiGAN"}) is B200HifiGAN
