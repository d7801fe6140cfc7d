"""CPU suite for the host side of the drop-in boundary: YAML/--hparams loader, checkpoint discovery in the
reference's on-disk formats, the binarized test-set reader + collate contract, rank dealing, and the world_size-2
weight broadcast over gloo (the N>1 path of SURVEY.md §8e without a GPU)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch
import yaml

from dict_tts_b200 import hparams as hp_mod, synth
from tests import fake_exp
from dict_tts_b200.config import AcousticConfig, VocoderConfig
from dict_tts_b200.data import DictTTSTestSet, IndexedDataset, IndexedDatasetBuilder
from dict_tts_b200.weights import (fold_weight_norm, get_last_checkpoint, load_acoustic_checkpoint,
                                   load_vocoder_checkpoint, pack_arena)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exp(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("exp"))
    return fake_exp.write(root, n_items=7)


# ---------------------------------------------------------------------------------------------------------------
# utils/hparams.py semantics
# ---------------------------------------------------------------------------------------------------------------
def test_hparams_chain_and_overrides(tmp_path):
    (tmp_path / "egs" / "base").mkdir(parents=True)
    (tmp_path / "egs" / "leaf").mkdir(parents=True)
    (tmp_path / "egs" / "base" / "a.yaml").write_text(yaml.safe_dump(dict(hidden_size=256, amp=False, lr=1.0,
                                                                            nested=dict(x=1, y=2), ks=[1, 2])))
    (tmp_path / "egs" / "base" / "b.yaml").write_text(yaml.safe_dump(dict(base_config="./a.yaml", hidden_size=192,
                                                                            nested=dict(y=3), vocoder="HifiGAN")))
    (tmp_path / "egs" / "leaf" / "c.yaml").write_text(yaml.safe_dump(dict(
        base_config=["egs/base/b.yaml", "egs/base/a.yaml"], word_size=8000)))
    cfg = hp_mod.set_hparams("egs/leaf/c.yaml", "", "amp=True,lr=0.5,nested.x=7,ks=[3 4 5],vocoder=pkg.Cls",
                             root=str(tmp_path), global_hparams=False)
    assert cfg["hidden_size"] == 192                  # the later file of the chain wins; a.yaml is not re-applied
    assert cfg["nested"] == dict(x=7, y=3)            # dicts merge key by key, overrides reach into them
    assert cfg["amp"] is True and cfg["lr"] == 0.5 and cfg["ks"] == [3, 4, 5]
    assert cfg["vocoder"] == "pkg.Cls" and cfg["word_size"] == 8000
    assert cfg["work_dir"] == "" and cfg["infer"] is True


def test_saved_config_overrides_chain_unless_reset(tmp_path):
    (tmp_path / "c.yaml").write_text(yaml.safe_dump(dict(hidden_size=192, hop_size=256)))
    ck = tmp_path / "checkpoints" / "e1"
    ck.mkdir(parents=True)
    (ck / "config.yaml").write_text(yaml.safe_dump(dict(hidden_size=128)))
    a = hp_mod.set_hparams("c.yaml", "e1", "", root=str(tmp_path), global_hparams=False)
    b = hp_mod.set_hparams("c.yaml", "e1", "", root=str(tmp_path), global_hparams=False, reset=True)
    assert a["hidden_size"] == 128 and b["hidden_size"] == 192 and a["hop_size"] == 256
    assert a["work_dir"] == os.path.join(str(tmp_path), "checkpoints/e1")       # follows root (cwd-relative without one)
    with pytest.raises(ValueError):
        hp_mod.set_hparams("", "", "", argv=[])


def test_cli_flags_match_reference():
    a = hp_mod.parse_args(["--config", "x.yaml", "--exp_name", "e", "--infer", "--hparams", "a=1", "--reset"])
    assert (a.config, a.exp_name, a.infer, a.hparams, a.reset, a.validate) == ("x.yaml", "e", True, "a=1", True, False)


def test_configs_from_hparams(exp):
    hp = exp["hparams"]
    a = AcousticConfig.from_hparams(hp)
    assert (a.hidden, a.n_heads, a.ffn_kernel, a.latent, a.frames_multiple, a.n_mel) == (192, 2, 5, 16, 4, 80)
    v = VocoderConfig.from_dict(fake_exp.VOC_CONFIG)
    assert list(v.up_rates) == [8, 8, 2, 2] and v.hop == 256 and v.init_ch == 512


# ---------------------------------------------------------------------------------------------------------------
# checkpoints (utils/ckpt_utils.py:8-25, vocoders/hifigan.py:16-52)
# ---------------------------------------------------------------------------------------------------------------
def test_newest_acoustic_checkpoint_is_loaded(exp):
    ckpt, path = get_last_checkpoint(exp["work_dir"])
    assert path.endswith("model_ckpt_steps_3000.ckpt") and ckpt["global_step"] == 3000
    sd, step = load_acoustic_checkpoint(exp["work_dir"], with_step=True)
    assert step == 3000
    want = fold_weight_norm(synth.make_acoustic_state_dict(1234))
    assert not any(k.startswith(("fvae.encoder.", "mel_disc.", "enc_pos_proj.")) for k in sd)
    k = "fvae.decoder.wn.in_layers.0.weight"
    assert torch.equal(sd[k], want[k]) and sd[k].abs().sum() > 0
    with pytest.raises(FileNotFoundError):
        load_acoustic_checkpoint(os.path.join(exp["root"], "nowhere"))


@pytest.mark.parametrize("original_layout", [False, True])
def test_vocoder_checkpoint_layouts(tmp_path, original_layout):
    e = fake_exp.write(str(tmp_path), n_items=1, original_hifigan_layout=original_layout)
    sd, cfg = load_vocoder_checkpoint(e["vocoder_dir"])
    assert cfg["upsample_rates"] == [8, 8, 2, 2]
    want = fold_weight_norm(synth.make_vocoder_state_dict(4321))
    assert set(sd) == set(want)
    assert torch.equal(sd["ups.1.weight"], want["ups.1.weight"])


# ---------------------------------------------------------------------------------------------------------------
# binarized data (utils/indexed_datasets.py, tasks/tts/dataset_utils.py:264-330)
# ---------------------------------------------------------------------------------------------------------------
def test_indexed_dataset_round_trip(tmp_path):
    b = IndexedDatasetBuilder(str(tmp_path / "ds"))
    items = [dict(i=i, a=np.arange(i + 1, dtype=np.float32)) for i in range(20)]
    for it in items:
        b.add_item(it)
    b.finalize()
    ds = IndexedDataset(str(tmp_path / "ds"))
    assert len(ds) == 20
    for i in (0, 7, 19):
        assert ds[i]["i"] == i and np.array_equal(ds[i]["a"], items[i]["a"])
    with pytest.raises(IndexError):
        ds[20]


def test_test_set_collate_contract(exp):
    ds = DictTTSTestSet(exp["hparams"])
    assert len(ds) == exp["n_items"]
    batch = next(ds.batches(max_sentences=4))
    B, Tw = batch["word_tokens"].shape
    assert B == 4 and batch["keys"].shape[:2] == (B, Tw) and batch["keys"].shape[3] == 768
    Lk, Lp = batch["key_map"].shape[2], batch["pinyin"].shape[2]
    assert batch["values"].shape == batch["keys"].shape and batch["pinyin_map"].shape == (B, Tw, Lp)
    assert batch["key_map"].dtype == torch.float32 and batch["pinyin"].dtype == torch.int64
    # BOS / EOS rows as the reference collater builds them (dataset_utils.py:286-296)
    assert (batch["keys"][:, 0] == 0).all() and (batch["key_map"][:, 0] == 1).all()
    assert (batch["pinyin"][:, 0] == 0).all() and (batch["pinyin_map"][:, 0] == 1).all()
    assert (batch["keys"][:, -1] == 0).all() and (batch["key_map"][:, -1] == 1).all()
    assert Lk >= 6 and Lp >= 2
    # sorted by length, longest first; padding of shorter utterances is zero
    ml = batch["mel_lengths"].tolist()
    assert ml == sorted(ml, reverse=True)
    for b in range(B):
        n = int(batch["word_lengths"][b])
        assert (batch["word_tokens"][b, n:] == 0).all() and (batch["word_tokens"][b, :n] > 0).all()
        assert (batch["mel2word"][b] > 0).sum() == batch["mel_lengths"][b]


def test_batches_are_dealt_round_robin(exp):
    ds = DictTTSTestSet(exp["hparams"])
    everything = [b["item_name"] for b in ds.batches(2)]
    r0 = [b["item_name"] for b in ds.batches(2, rank=0, world=2)]
    r1 = [b["item_name"] for b in ds.batches(2, rank=1, world=2)]
    assert r0 == everything[0::2] and r1 == everything[1::2]        # x[rank::world], tts_base.py:148-151
    names = sorted(n for b in everything for n in b)
    assert names == sorted(f"fake_{i:03d}" for i in range(exp["n_items"]))


def test_entry_point_needs_cuda(exp):
    """No CPU fallback: the --infer entry point must fail loudly when there is no CUDA device."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dict_tts_b200 import run
    cwd = os.getcwd()
    os.chdir(exp["root"])
    try:
        with pytest.raises(RuntimeError, match="CUDA"):
            run.main(["--exp_name", exp["exp"], "--infer"])
    finally:
        os.chdir(cwd)


# ---------------------------------------------------------------------------------------------------------------
# world_size 2 over gloo: one broadcast of the packed weight arena, then disjoint batches per rank
# ---------------------------------------------------------------------------------------------------------------
_WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r})
    import torch, torch.distributed as dist
    from dict_tts_b200 import synth
    from dict_tts_b200.data import DictTTSTestSet
    from dict_tts_b200.task import broadcast_arena
    from dict_tts_b200.weights import fold_weight_norm, pack_arena
    import yaml
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    meta, arena = [None], None
    if rank == 0:
        arena, table = pack_arena(fold_weight_norm(synth.make_vocoder_state_dict(4321)))
        meta = [(table, arena.numel())]
    dist.broadcast_object_list(meta, 0)
    table, numel = meta[0]
    got = broadcast_arena(arena, numel, "cpu", rank, world)
    want, _ = pack_arena(fold_weight_norm(synth.make_vocoder_state_dict(4321)))
    hp = yaml.safe_load(open({cfg!r}))
    names = [n for b in DictTTSTestSet(hp).batches(2, rank, world) for n in b["item_name"]]
    allnames = [None] * world
    dist.all_gather_object(allnames, names)
    if rank == 0:
        print(json.dumps(dict(equal=bool(torch.equal(got, want)), n=len(table), names=allnames)))
    ok = torch.tensor([int(torch.equal(got, want))])
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)
""")


def test_world_size_2_gloo_broadcast_and_sharding(exp, tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, cfg=os.path.join(exp["work_dir"], "config.yaml")))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["equal"] and res["n"] > 100
    r0, r1 = res["names"]
    assert not set(r0) & set(r1) and len(r0) + len(r1) == exp["n_items"]


# ---------------------------------------------------------------------------------------------------------------
# dictionary bank: host side
# ---------------------------------------------------------------------------------------------------------------
def test_dict_bank_round_trip_reproduces_the_collated_batch():
    from dict_tts_b200.bank import BOS_EOS, PAD, DictBank
    batch = synth.make_batch(seed=9, B=3, min_chars=2, max_chars=5, max_frames=32, Lk_cap=40)
    bank, ids = DictBank.from_batch(batch)
    n_chars = int((batch["word_tokens"] >= 3).sum())
    assert bank.n_entries == n_chars and int((ids >= 0).sum()) == n_chars
    assert int((ids == BOS_EOS).sum()) == 2 * 3 and int((ids == PAD).sum()) == int((batch["word_tokens"] == 0).sum())
    Lk, Lp = batch["key_map"].shape[2], batch["pinyin"].shape[2]
    assert bank.batch_dims(ids) <= (Lk, Lp) or bank.batch_dims(ids)[0] <= Lk
    again = bank.collate(ids, Lk, Lp)
    for k in ("keys", "values", "key_map", "pinyin", "pinyin_map"):
        assert torch.equal(again[k], batch[k]), k
    with pytest.raises(ValueError):
        bank.batch_dims(torch.tensor([[bank.n_entries]]))
    with pytest.raises(RuntimeError):
        bank.c_struct()                       # host bank: must be moved to the device first


def test_dict_ids_from_words():
    from dict_tts_b200.bank import ids_from_words
    w2i = {"<pad>": 0, "<EOS>": 1, "<UNK>": 2, "<BOS>": 3, "a": 4, "b": 5}
    ids = ids_from_words([["<BOS>", "a", "b", "<EOS>"], ["<BOS>", "zz", "<EOS>"]], w2i, 5)
    assert ids.tolist() == [[-1, 4, 5, -2, -1], [-1, 2, -2, -2, -1]]     # column 0 and Tw-1: the collater's added rows


def test_meta_csv_is_what_get_pron_error_reads(tmp_path):
    """meta.csv must parse the way scripts/get_pron_error.py reads it (reference: pd.DataFrame(outputs).to_csv,
    tts_base.py:371-372): a leading index column, the pinyin tokens in line.split(',')[3], one row per utterance in
    dataset order -- whatever order (longest-first batches, several ranks) the rows were produced in."""
    from dict_tts_b200.task import B200DictTTSTask
    rows = [dict(id=i, item_name=f"utt{i}", text=f"文本{i}", pinyin_tokens=f"a{i} 1 b{i} 4",
                 wav_fn_pred=f"[{i:06d}][utt{i}][P]", wav_fn_gt=f"[{i:06d}][utt{i}][G]") for i in (3, 0, 2, 1)]
    path = str(tmp_path / "meta.csv")
    B200DictTTSTask.write_meta(path, rows)
    with open(path) as f:
        lines = f.readlines()
    assert lines[0].strip() == ",item_name,text,pinyin_tokens,wav_fn_pred,wav_fn_gt"
    pred = []
    for line in lines[1:]:                                  # scripts/get_pron_error.py:31-44
        label = line.split(',')[3].replace('<UNK> ', '').replace('\n', '').split(' ')
        pron, out = '', []
        for i, item in enumerate(label):
            pron += item
            if i % 2 == 1:
                out.append(pron)
                pron = ''
        pred.append(" ".join(out))
    assert pred == [f"a{i}1 b{i}4" for i in range(4)]
    assert [line.split(',')[0] for line in lines[1:]] == ["0", "1", "2", "3"]
    import pandas as pd                                     # and it is byte for byte what pandas writes
    ref = str(tmp_path / "ref.csv")
    pd.DataFrame([{k: r[k] for k in B200DictTTSTask.META_FIELDS} for r in sorted(rows, key=lambda r: r["id"])]).to_csv(ref)
    assert open(ref).read() == open(path).read()
