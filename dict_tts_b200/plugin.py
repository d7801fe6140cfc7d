"""Drop-in plugin classes for the reference project's two seams.

* ``B200HifiGAN`` -- selected with ``--hparams vocoder=dict_tts_b200.plugin.B200HifiGAN`` through
  ``get_vocoder_cls`` (vocoders/base_vocoder.py:15-23).  Same constructor contract as ``vocoders.hifigan.HifiGAN``
  (:40-52: checkpoint directory = ``hparams['vocoder_ckpt']``, both on-disk layouts) and the same
  ``spec2wav(mel [T,80], **kw) -> float32 ndarray [T*hop]`` (:54-62); adds ``spec2wav_batch`` so a batch of mels
  stays on the device.
* ``B200DictTTSTask`` -- selected with ``--hparams task_cls=dict_tts_b200.plugin.B200DictTTSTask`` through
  ``tasks/run.py:6-11``.  When the reference package is importable this is a subclass of its ``DictTTSTask`` whose
  ``test_start`` swaps ``self.model`` for the CUDA engine after the checkpoint has been restored, so
  ``test_step`` / ``after_infer`` / ``test_end`` (tasks/tts/dict_tts.py:179-311) run unchanged.  Without the reference
  on the path it resolves to the standalone task in ``dict_tts_b200.task``.
Neither class has a CPU or PyTorch fallback: constructing them without libdtts.so / a CUDA device raises.
"""
import numpy as np
import torch

from . import hparams as hp_mod
from .config import AcousticConfig, VocoderConfig
from .engine import DictTTSEngine, HifiGanEngine
from .weights import drop_dead, fold_weight_norm, load_vocoder_checkpoint


def _reference_hparams():
    """The reference's global hparams dict when we run inside the reference process, else our own."""
    try:
        from utils.hparams import hparams as ref_hp      # noqa: WPS433 (reference package, optional)
        if ref_hp:
            return ref_hp
    except Exception:                                    # noqa: BLE001
        pass
    return hp_mod.hparams


class B200HifiGAN:
    """HiFi-GAN V1 generator on the B200 engine behind the BaseVocoder API."""

    def __init__(self, state_dict=None, config=None, device="cuda:0", precision=None):
        hp = _reference_hparams()
        if state_dict is None:
            state_dict, config = load_vocoder_checkpoint(hp["vocoder_ckpt"])
            print("| load B200 HifiGAN: ", hp["vocoder_ckpt"])
        else:
            state_dict = fold_weight_norm(state_dict)
        self.config = config
        cfg = VocoderConfig.from_dict(config) if config is not None else VocoderConfig()
        if precision is None:
            precision = int(hp.get("b200_vocoder_precision", 6))
        self.device = torch.device(device)
        self.engine = HifiGanEngine(state_dict, cfg, device, precision=precision)
        # reduced-precision modes are checked against the fp32-class mode on this very checkpoint and fall back if the
        # waveform tolerance is at risk (HifiGanEngine.self_check); b200_vocoder_selfcheck=False skips the probe
        self.selfcheck = None
        if hp.get("b200_vocoder_selfcheck", True):
            self.selfcheck = self.engine.self_check()
            if self.selfcheck["switched"]:
                print("| B200 HifiGAN: precision %d left the 1e-4 RMS budget on this checkpoint (probe RMS %.2e); "
                      "running the fp32-class mode (precision 1)" % (precision, self.selfcheck["rms"]))

    def spec2wav(self, mel, **kwargs):
        """mel: ndarray or tensor [T, n_mel] -> float32 ndarray [T*hop] (vocoders/hifigan.py:54-62)."""
        m = torch.as_tensor(np.asarray(mel) if not torch.is_tensor(mel) else mel, dtype=torch.float32)
        if m.dim() != 2:
            raise ValueError("spec2wav expects one utterance [T, n_mel]")
        return self.engine(m.unsqueeze(0)).view(-1).cpu().numpy()

    def spec2wav_batch(self, mel: torch.Tensor, lengths=None) -> torch.Tensor:
        """mel [B,T,n_mel] (host or device) -> device tensor [B, T*hop]; the mel never leaves HBM.  ``lengths`` [B]
        (valid frames per utterance): the waveform is computed up to lengths[b]*hop and zero after it."""
        return self.engine(mel, lengths)

    @staticmethod
    def wav2spec(wav_fn, return_linear=False):
        raise NotImplementedError("feature extraction is data preparation, outside the inference hot path")


class EngineModel:
    """Stands where ``task.model`` (PortaSpeech_dict) stood: same call signature, same returned dict."""

    def __init__(self, engine: DictTTSEngine):
        self.engine = engine

    def __call__(self, *args, **kwargs):
        return self.engine.forward(*args, **kwargs)

    forward = __call__

    def eval(self):
        return self

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    def parameters(self):
        return iter(())


def engine_from_module(model: torch.nn.Module, hp=None, device="cuda:0") -> EngineModel:
    """Builds the CUDA engine from a restored reference ``PortaSpeech_dict`` (weight-norm pairs are folded here, so it
    does not matter whether ``remove_weight_norm`` already ran)."""
    hp = hp or _reference_hparams()
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    cfg = AcousticConfig.from_hparams(hp) if hp else AcousticConfig()
    return EngineModel(DictTTSEngine(drop_dead(fold_weight_norm(sd)), cfg, device))


def _make_reference_task():
    from tasks.tts.dict_tts import DictTTSTask          # reference package (needs its full dependency set)

    class B200DictTTSTaskRef(DictTTSTask):
        def test_start(self):
            super().test_start()                        # vocoder + weight-norm removal (ps_flow.py:257-268)
            self.model = engine_from_module(self.model)

    B200DictTTSTaskRef.__name__ = "B200DictTTSTask"
    return B200DictTTSTaskRef


def __getattr__(name):
    if name == "B200DictTTSTask":
        try:
            return _make_reference_task()
        except Exception:                               # noqa: BLE001 -- reference not importable here
            from .task import B200DictTTSTask
            return B200DictTTSTask
    raise AttributeError(name)
