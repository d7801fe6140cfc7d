"""Test-set reader for the reference's binarized data: produces the collated HOST batches the hot path consumes.

On-disk formats are the reference's, read unchanged:
  * ``<prefix>.data`` / ``<prefix>.idx``  -- concatenated pickles + offset table (utils/indexed_datasets.py:7-54);
  * ``<prefix>_lengths.npy``             -- mel length per item (tasks/tts/dataset_utils.py:25-26);
  * ``word_set.json``, ``pinyin_encoder.pkl``, ``dict_embed.{data,idx}`` (dataset_utils.py:236-330,
    schema written by data_gen/tts/binarizer_zh.py:301-307: key, value, key_map, tokens_gloss, pinyin, pinyin_map).
Batch layout mirrors ``DictTTSDataset.collater`` (dataset_utils.py:264-302) for the keys the forward pass reads:
BOS/EOS rows are added to the dictionary tensors (keys/values 0, key_map 1, pinyin 0, pinyin_map 1).
Unlike the reference (``max_valid_sentences: 1``) batches may hold many utterances, sorted by length.
"""
import json
import os
import pickle
from typing import Dict, Iterator, List, Optional

import numpy as np
import torch

RESERVED = ["<pad>", "<EOS>", "<UNK>"]      # utils/text_encoder.py: PAD=0, EOS=1, UNK=2; vocab ids start at 3


class IndexedDataset:
    """Random access over ``path.data`` by the offsets in ``path.idx`` (np.save of {'offsets': [...]})."""

    def __init__(self, path: str):
        self.offsets = np.load(f"{path}.idx", allow_pickle=True).item()["offsets"]
        self.file = open(f"{path}.data", "rb")

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i: int):
        if i < 0 or i >= len(self):
            raise IndexError("index out of range")
        self.file.seek(self.offsets[i])
        return pickle.loads(self.file.read(self.offsets[i + 1] - self.offsets[i]))

    def close(self):
        if self.file:
            self.file.close()
            self.file = None

    def __del__(self):
        self.close()


class IndexedDatasetBuilder:
    """Writer of the same format (used by the tests and tools to fabricate a binarized set)."""

    def __init__(self, path: str):
        self.path = path
        self.out = open(f"{path}.data", "wb")
        self.offsets = [0]

    def add_item(self, item):
        self.offsets.append(self.offsets[-1] + self.out.write(pickle.dumps(item)))

    def finalize(self):
        self.out.close()
        with open(f"{self.path}.idx", "wb") as f:
            np.save(f, {"offsets": self.offsets})


def pad_1d(seqs: List[torch.Tensor], pad=0) -> torch.Tensor:
    n = max(int(s.shape[0]) for s in seqs)
    out = seqs[0].new_full((len(seqs), n), pad)
    for i, s in enumerate(seqs):
        out[i, :s.shape[0]] = s
    return out


def pad_2d(seqs: List[torch.Tensor], pad=0) -> torch.Tensor:
    n = max(int(s.shape[0]) for s in seqs)
    out = seqs[0].new_full((len(seqs), n, seqs[0].shape[1]), pad)
    for i, s in enumerate(seqs):
        out[i, :s.shape[0]] = s
    return out


def pad_3d(seqs: List[torch.Tensor], pad=0) -> torch.Tensor:
    n1 = max(int(s.shape[0]) for s in seqs)
    n2 = max(int(s.shape[1]) for s in seqs)
    out = seqs[0].new_full((len(seqs), n1, n2, seqs[0].shape[2]), pad)
    for i, s in enumerate(seqs):
        out[i, :s.shape[0], :s.shape[1]] = s
    return out


def collate_dict_ids(char_ids: List[torch.Tensor]) -> torch.Tensor:
    """Per-utterance bank ids of the characters (no BOS/EOS) -> ``dict_ids [B, Tw]`` naming exactly the rows the
    reference collater builds (dataset_utils.py:286-296): it pads the per-character tensors with zeros to the longest
    utterance and only THEN adds one row in front and one behind, so column 0 and column Tw-1 of EVERY utterance are
    the (keys 0, key_map 1, pinyin 0, pinyin_map 1) row (-1) and the EOS position of a shorter utterance is an
    all-zero row (-2) like the rest of its padding."""
    n = max(int(c.numel()) for c in char_ids)
    ids = torch.full((len(char_ids), n + 2), -2, dtype=torch.long)
    ids[:, 0] = ids[:, n + 1] = -1
    for b, c in enumerate(char_ids):
        ids[b, 1:1 + c.numel()] = c
    return ids


class DictTTSTestSet:
    def __init__(self, hp: Dict, prefix: str = "test", data_dir: Optional[str] = None):
        self.hp = hp
        self.dir = data_dir or hp["binary_data_dir"]
        self.prefix = prefix
        sizes = np.load(os.path.join(self.dir, f"{prefix}_lengths.npy"))
        n_test = hp.get("num_test_samples", 0)
        if n_test and n_test > 0:                                    # dataset_utils.py:32-37
            idxs = [i for i in range(n_test) if i < len(sizes)]
            idxs = list(hp.get("test_ids", [])) + idxs
        else:
            idxs = list(range(len(sizes)))
        if hp.get("min_frames", 0) > 0:
            idxs = [i for i in idxs if sizes[i] >= hp["min_frames"]]
        self.idxs = idxs
        self.sizes = [int(sizes[i]) for i in idxs]
        with open(os.path.join(self.dir, "word_set.json")) as f:
            words = json.load(f)
        self.word_to_id = {w: i for i, w in enumerate(RESERVED + [w for w in words if w not in RESERVED])}
        with open(os.path.join(self.dir, "pinyin_encoder.pkl"), "rb") as f:
            self.pinyin_encoder = pickle.load(f)
        self._pinyin_index = {p: i for i, p in enumerate(self.pinyin_encoder)}
        self.items = None
        self.dict_ds = None
        self.ragged = False          # True: items carry word ids and collate_ragged() builds a batch-local DictBank

    def __len__(self):
        return len(self.idxs)

    def _dict_entry(self, word: str):
        if self.dict_ds is None:
            self.dict_ds = IndexedDataset(os.path.join(self.dir, "dict_embed"))
        return self.dict_ds[self.word_to_id.get(word, 2)]            # 2 = <UNK> (dataset_utils.py:312-315)

    def build_bank(self):
        """The whole ``dict_embed`` table as one DictBank (entry index = word id, as get_dict_embeddings looks it up,
        dataset_utils.py:305-330).  After this call the items carry ``dict_ids`` instead of keys/values tensors."""
        from .bank import DictBank
        if self.dict_ds is None:
            self.dict_ds = IndexedDataset(os.path.join(self.dir, "dict_embed"))
        entries = []
        for wid in range(len(self.dict_ds)):
            e = self.dict_ds[wid]
            entries.append(dict(key=e["key"], value=e["value"], key_map=e["key_map"],
                                pinyin=[self._pinyin_index[p] for p in e["pinyin"]], pinyin_map=e["pinyin_map"]))
        self.bank = DictBank.from_entries(entries)
        return self.bank

    def __getitem__(self, i: int) -> Dict:
        if self.items is None:
            self.items = IndexedDataset(os.path.join(self.dir, self.prefix))
        item = self.items[self.idxs[i]]
        hp = self.hp
        fm = hp.get("frames_multiple", 1)
        T = min(len(item["mel"]), hp.get("max_frames", 1 << 30)) // fm * fm
        s = dict(id=i, item_name=item["item_name"], text=item.get("txt"), words=item["words"],
                 word_tokens=torch.LongTensor(item["word_tokens"]), mel_length=T)
        if item.get("mel2word") is not None:
            s["mel2word"] = torch.LongTensor(item["mel2word"])[:T]
        if "pron_modified" in item:
            s["pron_modified"] = torch.LongTensor(item["pron_modified"])
        if getattr(self, "bank", None) is not None:                  # characters named by bank id
            s["dict_ids"] = torch.LongTensor([self.word_to_id.get(w, 2) for w in item["words"][1:-1]])
            return s
        if self.ragged:                                              # word ids only; collate_ragged reads the entries
            s["word_ids"] = [self.word_to_id.get(w, 2) for w in item["words"][1:-1]]
            return s
        keys, values, key_map, pinyin, pinyin_map = [], [], [], [], []
        # The binarizer stores ONE feature tensor as both key and value (binarizer_zh.py:231-233; pickle keeps the identity):
        # then `values` stays the same tensor as `keys` all the way to the device and only half the bytes cross PCIe.
        alias = True
        for word in item["words"][1:-1]:                             # BOS / EOS carry no dictionary entry
            e = self._dict_entry(word)
            alias = alias and (e["value"] is e["key"])
            keys.append(torch.as_tensor(e["key"], dtype=torch.float32))
            values.append(keys[-1] if e["value"] is e["key"] else torch.as_tensor(e["value"], dtype=torch.float32))
            key_map.append(torch.as_tensor(e["key_map"], dtype=torch.float32))
            pinyin.append(torch.LongTensor([self._pinyin_index[p] for p in e["pinyin"]]))
            pinyin_map.append(torch.LongTensor(e["pinyin_map"]))
        s.update(keys=pad_2d(keys), key_map=pad_1d(key_map), pinyin=pad_1d(pinyin), pinyin_map=pad_1d(pinyin_map))
        s["values"] = s["keys"] if alias and keys else pad_2d(values)
        return s

    @staticmethod
    def collate(samples: List[Dict]) -> Dict:
        F = torch.nn.functional
        b = dict(id=torch.LongTensor([s["id"] for s in samples]), item_name=[s["item_name"] for s in samples],
                 text=[s["text"] for s in samples], words=[s["words"] for s in samples], nsamples=len(samples),
                 word_tokens=pad_1d([s["word_tokens"] for s in samples]),
                 word_lengths=torch.LongTensor([len(s["word_tokens"]) for s in samples]),
                 mel_lengths=torch.LongTensor([s["mel_length"] for s in samples]))
        b["txt_tokens"] = b["word_tokens"]
        b["mel2word"] = pad_1d([s["mel2word"] for s in samples]) if "mel2word" in samples[0] else None
        b["pron_modified"] = (pad_1d([s["pron_modified"] for s in samples]) if "pron_modified" in samples[0]
                              else None)
        if "dict_ids" in samples[0]:
            b["dict_ids"] = collate_dict_ids([s["dict_ids"] for s in samples])
            return b
        b["keys"] = F.pad(pad_3d([s["keys"] for s in samples]), (0, 0, 0, 0, 1, 1))
        if all(s["values"] is s["keys"] for s in samples):           # same tensor in every item: keep it one tensor
            b["values"] = b["keys"]
        else:
            b["values"] = F.pad(pad_3d([s["values"] for s in samples]), (0, 0, 0, 0, 1, 1))
        b["key_map"] = F.pad(pad_3d([s["key_map"].unsqueeze(-1) for s in samples]).squeeze(-1), (0, 0, 1, 1), value=1)
        b["pinyin"] = F.pad(pad_3d([s["pinyin"].unsqueeze(-1) for s in samples]).squeeze(-1), (0, 0, 1, 1), value=0)
        b["pinyin_map"] = F.pad(pad_3d([s["pinyin_map"].unsqueeze(-1) for s in samples]).squeeze(-1), (0, 0, 1, 1),
                                value=1)
        return b

    def collate_ragged(self, samples: List[Dict]) -> Dict:
        """Ragged (CSR) ``dict_msg``: instead of ``collate_3d``'s ``[B,Tw,Lk,768]`` padding
        (utils/__init__.py:153-167) the batch carries a batch-local DictBank holding each DISTINCT character of the
        batch once, un-padded, plus ``dict_ids [B,Tw]`` into it (-1 BOS/EOS row, -2 padding).  The engine reads the
        gloss tokens through the offsets (``dtts_text_encode_bank``), with results bit-identical to the padded batch."""
        from .bank import DictBank
        if self.dict_ds is None:
            self.dict_ds = IndexedDataset(os.path.join(self.dir, "dict_embed"))
        local, entries, rows = {}, [], []
        for s in samples:
            row = [-1]
            for wid in s["word_ids"]:
                if wid not in local:
                    e = self.dict_ds[wid]
                    local[wid] = len(entries)
                    entries.append(dict(key=e["key"], value=e["value"], key_map=e["key_map"],
                                        pinyin=[self._pinyin_index[p] for p in e["pinyin"]],
                                        pinyin_map=e["pinyin_map"]))
                row.append(local[wid])
            rows.append(torch.LongTensor(row[1:]))
        stripped = [{k: v for k, v in s.items() if k != "word_ids"} for s in samples]
        for s, r in zip(stripped, rows):
            s["dict_ids"] = r
        b = self.collate(stripped)
        b["dict_bank"] = DictBank.from_entries(entries) if entries else None
        return b

    def batches(self, max_sentences: int = 1, rank: int = 0, world: int = 1, sort_by_len: bool = True,
                max_tokens: Optional[int] = None, deal: str = "batches") -> Iterator[Dict]:
        """Collated batches for one rank.

        ``deal="batches"`` (default): longest-first groups of at most ``max_sentences`` utterances, whole groups dealt
        round-robin to the ranks -- every rank gets full-size batches of similar lengths (least padding).
        ``deal="reference"``: the reference's sampler (tasks/tts/tts_base.py:113-155 through
        ``batching.build_batch_sampler``): natural order, global groups of ``max_sentences * world`` (and at most
        ``max_tokens * world`` padded frames when given), every group dealt ``x[rank::world]``; groups that do not
        divide by ``world`` are kept here (the reference drops them -- it would lose utterances at inference)."""
        collate = self.collate_ragged if self.ragged else self.collate
        if deal == "reference":
            from .batching import build_batch_sampler
            groups = build_batch_sampler(self.sizes, max_tokens=max_tokens, max_sentences=max_sentences,
                                         by_size=max_tokens is not None, world=world, rank=rank,
                                         max_frames=self.hp.get("max_frames"), drop_ragged=False)
        elif deal == "batches":
            order = list(range(len(self)))
            if sort_by_len:
                order.sort(key=lambda i: -self.sizes[i])
            groups = [order[i:i + max_sentences] for i in range(0, len(order), max_sentences)][rank::world]
        else:
            raise ValueError("deal must be 'batches' or 'reference'")
        for g in groups:
            yield collate([self[i] for i in g])


class PortaSpeechTestSet:
    """Word-level PortaSpeech items as FastSpeechWordDataset reads and collates them (tasks/tts/dataset_utils.py:112-223):
    ``txt_tokens`` = phoneme ids (``item['phone']``), ``ph2word``, ``word_lengths`` (BOS / EOS included), ``mel2word`` for
    the profiling mode.  Same sampler entry points as DictTTSTestSet."""

    def __init__(self, hp: Dict, prefix: str = "test", data_dir: Optional[str] = None):
        self.hp = hp
        self.dir = data_dir or hp["binary_data_dir"]
        self.prefix = prefix
        sizes = np.load(os.path.join(self.dir, f"{prefix}_lengths.npy"))
        n_test = hp.get("num_test_samples", 0)
        if n_test and n_test > 0:
            idxs = list(hp.get("test_ids", [])) + [i for i in range(n_test) if i < len(sizes)]
        else:
            idxs = list(range(len(sizes)))
        if hp.get("min_frames", 0) > 0:
            idxs = [i for i in idxs if sizes[i] >= hp["min_frames"]]
        self.idxs = idxs
        self.sizes = [int(sizes[i]) for i in idxs]
        self.items = None
        path = os.path.join(self.dir, "phone_set.json")
        self.phone_list = None
        if os.path.exists(path):
            with open(path) as f:
                self.phone_list = json.load(f)

    def __len__(self):
        return len(self.idxs)

    def __getitem__(self, i: int) -> Dict:
        if self.items is None:
            self.items = IndexedDataset(os.path.join(self.dir, self.prefix))
        item = self.items[self.idxs[i]]
        hp = self.hp
        fm = hp.get("frames_multiple", 1)
        T = min(len(item["mel"]), hp.get("max_frames", 1 << 30)) // fm * fm
        n_in = hp.get("max_input_tokens", 1 << 30)
        s = dict(id=i, item_name=item["item_name"], text=item.get("txt"), words=item.get("words"),
                 txt_tokens=torch.LongTensor(item["phone"][:n_in]), ph2word=torch.LongTensor(item["ph2word"][:n_in]),
                 n_words=len(item["word_tokens"]), mel_length=T)
        if item.get("mel2word") is not None:
            s["mel2word"] = torch.LongTensor(item["mel2word"])[:T]
        return s

    @staticmethod
    def collate(samples: List[Dict]) -> Dict:
        b = dict(id=torch.LongTensor([s["id"] for s in samples]), item_name=[s["item_name"] for s in samples],
                 text=[s["text"] for s in samples], words=[s["words"] for s in samples], nsamples=len(samples),
                 txt_tokens=pad_1d([s["txt_tokens"] for s in samples]), ph2word=pad_1d([s["ph2word"] for s in samples]),
                 txt_lengths=torch.LongTensor([s["txt_tokens"].numel() for s in samples]),
                 word_lengths=torch.LongTensor([s["n_words"] for s in samples]),
                 mel_lengths=torch.LongTensor([s["mel_length"] for s in samples]))
        b["mel2word"] = pad_1d([s["mel2word"] for s in samples]) if "mel2word" in samples[0] else None
        return b

    def batches(self, max_sentences: int = 1, rank: int = 0, world: int = 1, sort_by_len: bool = True,
                max_tokens: Optional[int] = None, deal: str = "batches") -> Iterator[Dict]:
        if deal == "reference":
            from .batching import build_batch_sampler
            groups = build_batch_sampler(self.sizes, max_tokens=max_tokens, max_sentences=max_sentences,
                                         by_size=max_tokens is not None, world=world, rank=rank,
                                         max_frames=self.hp.get("max_frames"), drop_ragged=False)
        else:
            order = list(range(len(self)))
            if sort_by_len:
                order.sort(key=lambda i: -self.sizes[i])
            groups = [order[i:i + max_sentences] for i in range(0, len(order), max_sentences)][rank::world]
        for g in groups:
            yield self.collate([self[i] for i in g])
