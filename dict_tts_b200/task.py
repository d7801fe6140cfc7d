"""Standalone inference task: what ``tasks/run.py --infer`` does for ``DictTTSTask`` in the reference, without the
training stack (Trainer / DDP / TensorBoard / matplotlib).

Flow mirrored:
  start()       Trainer.test -> fit -> build_model -> restore newest checkpoint     utils/trainer.py:90-120
  test_start()  gen dir, vocoder = get_vocoder_cls(hparams)(), fold weight-norm     tts_base.py:247-254, ps_flow.py:257
  test_step()   model(...) with the reference's argument tuple                      dict_tts.py:179-196
  after_infer() spec2wav, int16 wav, argmax(pron_attn) -> pinyin tokens             dict_tts.py:227-311
  test_end()    meta.csv                                                            tts_base.py:371-376
Differences, all on purpose: utterances are batched (``max_sentences``), the mel stays on the device between the
acoustic model and the vocoder, and with several GPUs every rank owns ``batches[rank::world]`` and receives the
weights through one NCCL broadcast.
"""
import csv
import importlib
import os
import pickle
from typing import Dict, List, Optional

import numpy as np
import torch

from . import hparams as hp_mod
from .config import AcousticConfig, PortaSpeechConfig, VocoderConfig
from .data import DictTTSTestSet, PortaSpeechTestSet
from .engine import DictTTSEngine, HifiGanEngine, PortaSpeechEngine
from .weights import fold_weight_norm, get_last_checkpoint, load_acoustic_checkpoint, load_vocoder_checkpoint, pack_arena


def get_vocoder_cls(hp):
    """Name registry + dotted-path import, as vocoders/base_vocoder.py:15-23."""
    name = hp.get("vocoder", "HifiGAN")
    if name in ("HifiGAN", "hifigan", "vocoders.hifigan.HifiGAN", "B200HifiGAN"):
        from .plugin import B200HifiGAN
        return B200HifiGAN
    pkg, cls = name.rsplit(".", 1)
    return getattr(importlib.import_module(pkg), cls)


def save_wav(wav: np.ndarray, path: str, sr: int, norm: bool = False) -> None:
    """utils/audio.py:11-16: optional peak normalisation, x32767, int16 PCM."""
    from scipy.io import wavfile
    wav = np.asarray(wav, dtype=np.float32)
    if norm:
        wav = wav / np.abs(wav).max()
    wavfile.write(path, sr, (wav * 32767).astype(np.int16))


def broadcast_arena(host_arena: Optional[torch.Tensor], numel: int, device, rank: int, world: int) -> torch.Tensor:
    """Rank 0 holds the packed weights; everyone leaves with a device copy (one collective)."""
    import torch.distributed as dist
    if world == 1:
        return host_arena.to(device)
    buf = host_arena.to(device) if rank == 0 else torch.empty(numel, dtype=torch.float32, device=device)
    dist.broadcast(buf, 0)
    return buf


class B200DictTTSTask:
    def __init__(self, hp: Optional[Dict] = None, device: str = "cuda:0", rank: int = 0, world: int = 1):
        self.hp = hp if hp is not None else hp_mod.hparams
        self.device, self.rank, self.world = device, rank, world
        self.model: Optional[DictTTSEngine] = None
        self.vocoder = None
        self.global_step = 0
        self.results: List[Dict] = []

    # ---- reference entry point: task_cls.start() (tasks/base_task.py:318-352, --infer branch) -------------------
    @classmethod
    def start(cls):
        hp = hp_mod.hparams
        if not hp.get("infer", True):
            raise SystemExit("dict_tts_b200 implements the --infer path only")
        if not torch.cuda.is_available():
            raise RuntimeError("B200DictTTSTask needs a CUDA device (sm_100a); there is no CPU fallback")
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        task = cls(hp, f"cuda:{local}", rank, world)
        task.build_model()
        task.test_start()
        outputs = []
        ds = DictTTSTestSet(hp, hp.get("test_set_name", "test"))
        # dictionary features: "bank" (default) uploads dict_embed once, batches then carry ids only;
        # "ragged": every batch carries an un-padded bank of its own distinct characters (§8f-4, for dictionaries that
        # do not fit in HBM); "padded": the reference collater's [B,Tw,Lk,768] tensors.  b200_dict_bank=False = "padded".
        mode = str(hp.get("b200_dict_mode", "bank" if hp.get("b200_dict_bank", True) else "padded"))
        if mode not in ("bank", "ragged", "padded"):
            raise ValueError("b200_dict_mode must be bank, ragged or padded")
        if int(hp.get("b200_s2pa_route", 0)) == 1:
            mode = "padded"                                # the projection-GEMM route reads the collated tensors
        if mode == "bank":
            task.model.set_dict_bank(ds.build_bank())
        ds.ragged = mode == "ragged"
        bs = int(hp.get("b200_max_sentences", hp.get("max_valid_sentences", 1)) or 1)
        max_tokens = hp.get("b200_max_tokens")             # padded frames per batch and device, as max_tokens is upstream
        for i, batch in enumerate(ds.batches(bs, rank, world, max_tokens=max_tokens,
                                             deal=str(hp.get("b200_deal", "batches")))):
            if batch.get("dict_bank") is not None:
                task.model.set_dict_bank(batch["dict_bank"])
            outputs.extend(task.test_step(batch, i))
        task.test_end(outputs)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return outputs

    # ---- build_model + restore_weights ---------------------------------------------------------------------------
    def build_model(self):
        import torch.distributed as dist
        hp, rank, world = self.hp, self.rank, self.world
        acfg = AcousticConfig.from_hparams(hp)
        meta = [None]
        arena = None
        if rank == 0:
            sd, self.global_step = load_acoustic_checkpoint(hp["work_dir"], with_step=True)
            arena, table = pack_arena(sd)
            meta = [(table, arena.numel(), self.global_step)]
        if world > 1:
            dist.broadcast_object_list(meta, 0)
        table, numel, self.global_step = meta[0]
        dev_arena = broadcast_arena(arena, numel, self.device, rank, world)
        self.model = DictTTSEngine(None, acfg, self.device, arena=dev_arena, table=table,
                                   precision=int(hp.get("b200_acoustic_precision", 1)),
                                   s2pa_route=int(hp.get("b200_s2pa_route", 0)))
        self.model.profile_infer = bool(hp.get("profile_infer", False))     # stage Timers print like upstream
        return self.model

    def test_start(self):
        hp = self.hp
        self.gen_dir = os.path.join(hp["work_dir"], f'generated_{self.global_step}_{hp.get("gen_dir_name", "")}')
        os.makedirs(os.path.join(self.gen_dir, "wavs"), exist_ok=True)
        cls = get_vocoder_cls(hp)
        self.vocoder = cls(device=self.device) if cls.__name__ == "B200HifiGAN" else cls()
        if hasattr(self.vocoder, "engine"):
            self.vocoder.engine.profile_infer = bool(hp.get("profile_infer", False))
        path = os.path.join(hp["binary_data_dir"], "pinyin_encoder.pkl")
        self.pinyin_encoder = None
        if os.path.exists(path):
            with open(path, "rb") as f:
                self.pinyin_encoder = pickle.load(f)
        self.results_id = 0

    def run_model(self, sample: Dict) -> Dict:
        """The exact call of DictTTSTask.test_step (dict_tts.py:183-196)."""
        hp = self.hp
        bank = "dict_ids" in sample
        return self.model(
            (sample["word_tokens"], sample["txt_tokens"]), sample.get("pron_modified"), (None, None, None),
            ph2word=sample.get("ph2word"), word_len=sample["word_lengths"].max(),
            dict_msg=None if bank else (sample["keys"], sample["values"], sample["key_map"], sample["pinyin"],
                                        sample["pinyin_map"]),
            dict_ids=sample["dict_ids"] if bank else None,
            infer=True, forward_post_glow=False, spk_embed=None, two_stage=hp.get("two_stage", True),
            mel2word=sample["mel2word"] if hp.get("profile_infer", False) else None)

    @torch.no_grad()
    def test_step(self, sample: Dict, batch_idx: int) -> List[Dict]:
        out = self.run_model(sample)
        sample["outputs"] = out["mel_out"]
        sample["pron_attn"] = out["pron_attn"]
        sample["mel2word_pred"] = out["mel2word"]
        return self.after_infer(sample)

    def after_infer(self, sample: Dict) -> List[Dict]:
        hp = self.hp
        mel = sample["outputs"]                                   # [B,T,80] on the device
        B = mel.shape[0]
        hop = hp.get("hop_size", 256)
        pcm = None
        if hasattr(self.vocoder, "spec2wav_batch"):
            frames_dev = (sample["mel2word_pred"] > 0).sum(-1)
            # B = 1 keeps the padded tail like the reference; a batch is vocoded up to each utterance's valid length
            wav = self.vocoder.spec2wav_batch(mel, frames_dev if B > 1 else None)   # [B, T*hop], mel never leaves HBM
            if not hp.get("out_wav_norm", False):
                pcm = self.model.pcm16(wav).cpu().numpy()         # int16 on the device: half the D2H bytes
            else:
                wav = wav.cpu().numpy()
        else:
            wav = np.stack([self.vocoder.spec2wav(mel[b].cpu().numpy()) for b in range(B)])
        frames = (sample["mel2word_pred"] > 0).sum(-1).cpu().numpy()      # valid frames per utterance
        # dict_tts.py:295-304 on the device: argmax(pron_attn) -> the two pinyin ids of every character
        pairs = self.model.pron_tokens(sample["pron_attn"], pinyin=sample.get("pinyin"),
                                       dict_ids=sample.get("dict_ids")).cpu()
        results = []
        for b in range(B):
            name, text = sample["item_name"][b], sample["text"][b]
            # the reference numbers its files with results_id = position in the (un-shuffled, B = 1) test set; batches
            # here arrive longest-first and per rank, so the dataset index is that number
            uid = int(sample["id"][b]) if "id" in sample else self.results_id
            base_fn = f'[{uid:06d}][{str(name).replace("%", "_")}][%s]'
            if text is not None:
                base_fn += str(text).replace(":", "$3A")[:80]
            base_fn = base_fn.replace(" ", "_")
            total = pcm.shape[1] if pcm is not None else wav.shape[1]
            n = int(frames[b]) * hop if B > 1 else total          # B=1: keep the padded tail like the reference
            if not hp.get("profile_infer", False):
                path = os.path.join(self.gen_dir, "wavs", (base_fn % "P") + ".wav")
                if pcm is not None:
                    from scipy.io import wavfile
                    wavfile.write(path, hp.get("audio_sample_rate", 22050), pcm[b, :n])
                else:
                    save_wav(wav[b, :n], path, hp.get("audio_sample_rate", 22050), norm=True)
            # dict_tts.py:295-304: two pinyin tokens per character from argmax(pron_attn)
            tokens = []
            if self.pinyin_encoder is not None:
                n_words = int(sample["word_lengths"][b])
                for i in range(1, n_words - 1):
                    for t in pairs[b, i].tolist():
                        if t >= 0:
                            tokens.append(self.pinyin_encoder[int(t)])
            results.append(dict(id=uid, item_name=name,
                                text=None if text is None else str(text).replace(",", "，").replace(".", "。"),
                                pinyin_tokens=" ".join(tokens), wav_fn_pred=base_fn % "P", wav_fn_gt=base_fn % "G"))
            self.results_id += 1
        return results

    META_FIELDS = ["item_name", "text", "pinyin_tokens", "wav_fn_pred", "wav_fn_gt"]

    @classmethod
    def write_meta(cls, path: str, rows: List[Dict]) -> None:
        """meta.csv exactly as the reference writes it, ``pd.DataFrame(outputs).to_csv`` (tts_base.py:371-372): a
        leading unnamed index column, then the result fields, one row per utterance in DATASET order -- the order
        scripts/get_pron_error.py aligns against label_set0.csv, reading the pinyin tokens as ``line.split(',')[3]``."""
        rows = sorted(rows, key=lambda r: r.get("id", 0))
        with open(path, "w", newline="") as f:
            w = csv.writer(f, lineterminator="\n")
            w.writerow([""] + cls.META_FIELDS)
            for i, r in enumerate(rows):
                w.writerow([i] + ["" if r.get(k) is None else r[k] for k in cls.META_FIELDS])

    def test_end(self, outputs: List[Dict]):
        """One ordered meta.csv.  With several ranks every rank leaves its rows (with dataset ids) in a side file and
        rank 0 merges them after the barrier."""
        if self.world == 1:
            self.write_meta(os.path.join(self.gen_dir, "meta.csv"), outputs)
            return {}
        import torch.distributed as dist
        with open(os.path.join(self.gen_dir, f"meta.rank{self.rank}.pkl"), "wb") as f:
            pickle.dump(outputs, f)
        dist.barrier()
        if self.rank == 0:
            rows = []
            for r in range(self.world):
                part = os.path.join(self.gen_dir, f"meta.rank{r}.pkl")
                with open(part, "rb") as f:
                    rows.extend(pickle.load(f))
                os.remove(part)
            self.write_meta(os.path.join(self.gen_dir, "meta.csv"), rows)
        return {}


class B200PortaSpeechTask(B200DictTTSTask):
    """The PortaSpeech (non-dict) sibling behind the same ``task_cls`` seam: what ``tasks/run.py --infer``
    does for ``PortaSpeechFlowTask`` (tasks/tts/ps_flow.py:257-312, tasks/tts/tts_base.py:247-376) at ``dur_level: word`` /
    ``use_post_glow: False``.  Selected with ``--hparams task_cls=dict_tts_b200.task.B200PortaSpeechTask``."""
    META_FIELDS = ["item_name", "text", "ph_tokens", "wav_fn_pred", "wav_fn_gt"]

    @classmethod
    def start(cls):
        hp = hp_mod.hparams
        if not hp.get("infer", True):
            raise SystemExit("dict_tts_b200 implements the --infer path only")
        if hp.get("use_post_glow", False):
            raise SystemExit("use_post_glow: true needs modules/glow, which the reference checkout does not ship; "
                             "run with --hparams use_post_glow=False (mel_out = mel_out_fvae)")
        if not torch.cuda.is_available():
            raise RuntimeError("B200PortaSpeechTask needs a CUDA device (sm_100a); there is no CPU fallback")
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        if world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        task = cls(hp, f"cuda:{local}", rank, world)
        task.build_model()
        task.test_start()
        ds = PortaSpeechTestSet(hp, hp.get("test_set_name", "test"))
        task.phone_list = ds.phone_list
        outputs = []
        bs = int(hp.get("b200_max_sentences", hp.get("max_valid_sentences", 1)) or 1)
        for i, batch in enumerate(ds.batches(bs, rank, world, max_tokens=hp.get("b200_max_tokens"),
                                             deal=str(hp.get("b200_deal", "batches")))):
            outputs.extend(task.test_step(batch, i))
        task.test_end(outputs)
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return outputs

    def build_model(self):
        import torch.distributed as dist
        hp, rank, world = self.hp, self.rank, self.world
        meta, arena = [None], None
        if rank == 0:
            ckpt, _ = get_last_checkpoint(hp["work_dir"])
            if ckpt is None:
                raise FileNotFoundError(f'no model_ckpt_steps_*.ckpt under {hp["work_dir"]}')
            self.global_step = int(ckpt.get("global_step", 0))
            sd = {k: v for k, v in fold_weight_norm(ckpt["state_dict"]["model"]).items()
                  if not k.startswith(("fvae.encoder.", "post_flow.", "g_proj.")) and v.is_floating_point()}
            arena, table = pack_arena(sd)
            meta = [(table, arena.numel(), self.global_step, int(sd["ph_encoder.emb.weight"].shape[0]))]
        if world > 1:
            dist.broadcast_object_list(meta, 0)
        table, numel, self.global_step, ph_size = meta[0]
        cfg = PortaSpeechConfig.from_hparams(hp, ph_size)
        dev_arena = broadcast_arena(arena, numel, self.device, rank, world)
        self.model = PortaSpeechEngine(None, cfg, self.device, arena=dev_arena, table=table,
                                       precision=int(hp.get("b200_acoustic_precision", 1)))
        self.model.profile_infer = bool(hp.get("profile_infer", False))
        return self.model

    def run_model(self, sample: Dict) -> Dict:
        """The exact call of PortaSpeechFlowTask.test_step (ps_flow.py:274-284)."""
        hp = self.hp
        return self.model(sample["txt_tokens"], ph2word=sample["ph2word"], word_len=sample["word_lengths"].max(),
                          infer=True, forward_post_glow=False, spk_embed=None, two_stage=hp.get("two_stage", True),
                          mel2word=sample["mel2word"] if hp.get("profile_infer", False) else None)

    @torch.no_grad()
    def test_step(self, sample: Dict, batch_idx: int) -> List[Dict]:
        out = self.run_model(sample)
        sample["outputs"] = out["mel_out"]
        sample["mel2word_pred"] = out["mel2word"]
        return self.after_infer(sample)

    def after_infer(self, sample: Dict) -> List[Dict]:
        """TTSBaseTask.after_infer (tts_base.py:256-334) for a batch: vocode on the device, int16 wavs, one row per utterance."""
        hp = self.hp
        mel = sample["outputs"]
        B = mel.shape[0]
        hop = hp.get("hop_size", 256)
        frames_dev = (sample["mel2word_pred"] > 0).sum(-1)
        wav = self.vocoder.spec2wav_batch(mel, frames_dev if B > 1 else None)
        pcm = self.model.pcm16(wav).cpu().numpy()
        frames = frames_dev.cpu().numpy()
        results = []
        for b in range(B):
            name, text = sample["item_name"][b], sample["text"][b]
            uid = int(sample["id"][b])
            base_fn = f'[{uid:06d}][{str(name).replace("%", "_")}][%s]'
            if text is not None:
                base_fn += str(text).replace(":", "$3A")[:80]
            base_fn = base_fn.replace(" ", "_")
            n = int(frames[b]) * hop if B > 1 else pcm.shape[1]
            if not hp.get("profile_infer", False):
                from scipy.io import wavfile
                wavfile.write(os.path.join(self.gen_dir, "wavs", (base_fn % "P") + ".wav"),
                              hp.get("audio_sample_rate", 22050), pcm[b, :n])
            toks = [int(t) for t in sample["txt_tokens"][b].tolist() if t > 0]
            if getattr(self, "phone_list", None):                 # TokenTextEncoder.decode: ids -> phone strings
                reserved = ["<pad>", "<EOS>", "<UNK>"]
                vocab = reserved + [p for p in self.phone_list if p not in reserved]
                ph = " ".join(vocab[t] if t < len(vocab) else "<UNK>" for t in toks)
            else:
                ph = " ".join(str(t) for t in toks)
            results.append(dict(id=uid, item_name=name, text=text, ph_tokens=ph, wav_fn_pred=base_fn % "P",
                                wav_fn_gt=base_fn % "G"))
            self.results_id += 1
        return results
