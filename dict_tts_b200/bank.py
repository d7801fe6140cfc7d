"""GPU-resident dictionary bank.

In the reference every batch carries ``keys``/``values`` ``[B,Tw,Lk,768]`` (778 MB at batch 60) that
``DictTTSDataset.get_dict_embeddings`` / ``collater`` (tasks/tts/dataset_utils.py:264-330) re-read from
``dict_embed.{data,idx}`` and re-pad for every utterance, although they are a pure function of the character id.
Here the per-character entries are concatenated once (ragged, CSR offsets), uploaded once, and a batch only names
its characters: ``dict_ids [B,Tw]`` (>= 0 entry, -1 the BOS/EOS row the collater builds, -2 padding).
``DictBank.collate`` rebuilds the explicit tensors exactly as the collater would -- the compat path and the parity
oracle of the bank path.
"""
import ctypes as C
from typing import Dict, List, Sequence, Tuple

import torch

from . import binding

BOS_EOS, PAD = -1, -2


class DictBank:
    def __init__(self, keys: torch.Tensor, values: torch.Tensor, key_map: torch.Tensor, tok_offsets: torch.Tensor,
                 pinyin: torch.Tensor, pinyin_map: torch.Tensor, pin_offsets: torch.Tensor):
        self.keys, self.values, self.key_map = keys.float().contiguous(), values.float().contiguous(), key_map.float()
        self.tok_offsets = tok_offsets.long().contiguous()
        self.pinyin, self.pinyin_map = pinyin.long().contiguous(), pinyin_map.long().contiguous()
        self.pin_offsets = pin_offsets.long().contiguous()
        n = self.tok_offsets.numel() - 1
        if n <= 0 or self.pin_offsets.numel() != n + 1:
            raise ValueError("bank needs at least one entry and matching offset tables")
        if int(self.tok_offsets[-1]) != self.keys.shape[0] or self.keys.shape != self.values.shape \
                or self.key_map.numel() != self.keys.shape[0] or int(self.pin_offsets[-1]) != self.pinyin.numel():
            raise ValueError("bank tensors do not match their offset tables")
        self.n_entries = n
        self._host_tok = self.tok_offsets.cpu()
        self._host_pin = self.pin_offsets.cpu()
        self._struct = None

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_entries(cls, entries: Sequence[Dict]) -> "DictBank":
        """entries[i]: key [L,768], value [L,768], key_map [L], pinyin [P] (ids), pinyin_map [P] -- the schema
        data_gen/tts/binarizer_zh.py:301-307 writes, with pinyin already mapped through pinyin_encoder."""
        D = int(torch.as_tensor(entries[0]["key"]).shape[-1])
        keys = torch.cat([torch.as_tensor(e["key"], dtype=torch.float32).reshape(-1, D) for e in entries])
        values = torch.cat([torch.as_tensor(e["value"], dtype=torch.float32).reshape(-1, D) for e in entries])
        key_map = torch.cat([torch.as_tensor(e["key_map"], dtype=torch.float32).reshape(-1) for e in entries])
        pinyin = torch.cat([torch.as_tensor(e["pinyin"], dtype=torch.long).reshape(-1) for e in entries])
        pinyin_map = torch.cat([torch.as_tensor(e["pinyin_map"], dtype=torch.long).reshape(-1) for e in entries])
        tl = torch.tensor([0] + [len(e["key_map"]) for e in entries]).cumsum(0)
        pl = torch.tensor([0] + [len(e["pinyin_map"]) for e in entries]).cumsum(0)
        return cls(keys, values, key_map, tl, pinyin, pinyin_map, pl)

    @classmethod
    def from_batch(cls, batch: Dict[str, torch.Tensor]) -> Tuple["DictBank", torch.Tensor]:
        """Turns an explicit collated batch into (bank, dict_ids): every real character position becomes one entry.
        Used by the tests and by bench.py to drive the bank path with the same data as the explicit path."""
        wt, keys, values = batch["word_tokens"], batch["keys"], batch["values"]
        key_map, pinyin, pinyin_map = batch["key_map"], batch["pinyin"], batch["pinyin_map"]
        B, Tw = wt.shape
        ids = torch.full((B, Tw), PAD, dtype=torch.long)
        entries = []
        for b in range(B):
            for t in range(Tw):
                tok = int(wt[b, t])
                if tok <= 1:                                   # padding or BOS / EOS (utils/text_encoder.py: EOS = 1):
                    if bool((key_map[b, t] == 1).all()):       # the row the collater adds (keys 0, key_map 1, pinyin_map 1)
                        ids[b, t] = BOS_EOS                    # ... wherever THIS batch has it (dataset_utils.py:286-296)
                    continue
                nz = (keys[b, t].abs().sum(-1) != 0).nonzero()
                L = int(nz.max()) + 1 if nz.numel() else 1
                P = max(int((pinyin_map[b, t] != 0).sum()), 1)
                ids[b, t] = len(entries)
                entries.append(dict(key=keys[b, t, :L], value=values[b, t, :L], key_map=key_map[b, t, :L],
                                    pinyin=pinyin[b, t, :P], pinyin_map=pinyin_map[b, t, :P]))
        return cls.from_entries(entries), ids

    # ---- host-side helpers ----------------------------------------------------------------------------------------
    def batch_dims(self, dict_ids: torch.Tensor) -> Tuple[int, int]:
        """(Lk, Lp) the collater would pad this batch to."""
        ids = dict_ids.cpu().reshape(-1)
        ids = ids[ids >= 0]
        if ids.numel() and int(ids.max()) >= self.n_entries:
            raise ValueError("dict_ids refers to an entry outside the bank")
        if not ids.numel():
            return 1, 1
        lk = int((self._host_tok[ids + 1] - self._host_tok[ids]).max())
        lp = int((self._host_pin[ids + 1] - self._host_pin[ids]).max())
        return max(lk, 1), max(lp, 1)

    def collate(self, dict_ids: torch.Tensor, Lk: int = 0, Lp: int = 0) -> Dict[str, torch.Tensor]:
        """Explicit dict_msg tensors for dict_ids, padded like DictTTSDataset.collater (dataset_utils.py:286-296)."""
        lk, lp = self.batch_dims(dict_ids)
        Lk, Lp = max(Lk, lk), max(Lp, lp)
        B, Tw = dict_ids.shape
        D = self.keys.shape[1]
        keys, values = torch.zeros(B, Tw, Lk, D), torch.zeros(B, Tw, Lk, D)
        key_map = torch.zeros(B, Tw, Lk)
        pinyin = torch.zeros(B, Tw, Lp, dtype=torch.long)
        pinyin_map = torch.zeros(B, Tw, Lp, dtype=torch.long)
        hk, hv, hm = self.keys.cpu(), self.values.cpu(), self.key_map.cpu()
        hp, hpm = self.pinyin.cpu(), self.pinyin_map.cpu()
        for b in range(B):
            for t in range(Tw):
                i = int(dict_ids[b, t])
                if i == BOS_EOS:
                    key_map[b, t] = 1
                    pinyin_map[b, t] = 1
                elif i >= 0:
                    a, e = int(self._host_tok[i]), int(self._host_tok[i + 1])
                    keys[b, t, :e - a], values[b, t, :e - a], key_map[b, t, :e - a] = hk[a:e], hv[a:e], hm[a:e]
                    a, e = int(self._host_pin[i]), int(self._host_pin[i + 1])
                    pinyin[b, t, :e - a], pinyin_map[b, t, :e - a] = hp[a:e], hpm[a:e]
        return dict(keys=keys, values=values, key_map=key_map, pinyin=pinyin, pinyin_map=pinyin_map)

    # ---- device side ------------------------------------------------------------------------------------------------
    def to(self, device, non_blocking: bool = False) -> "DictBank":
        alias = self.values.data_ptr() == self.keys.data_ptr()
        mv = lambda t: t.to(device, non_blocking=non_blocking)      # noqa: E731
        keys = mv(self.keys)
        values = keys if alias or torch.equal(self.keys, self.values) else mv(self.values)
        out = DictBank.__new__(DictBank)                            # offsets were validated when self was built
        out.keys, out.values, out.key_map, out.tok_offsets = keys, values, mv(self.key_map), mv(self.tok_offsets)
        out.pinyin, out.pinyin_map, out.pin_offsets = mv(self.pinyin), mv(self.pinyin_map), mv(self.pin_offsets)
        out.n_entries, out._host_tok, out._host_pin, out._struct = self.n_entries, self._host_tok, self._host_pin, None
        return out

    def pin_memory(self) -> "DictBank":
        """Host bank in pinned memory, so that to(device, non_blocking=True) overlaps with compute."""
        alias = self.values.data_ptr() == self.keys.data_ptr() or torch.equal(self.keys, self.values)
        keys = self.keys.pin_memory()
        return DictBank(keys, keys if alias else self.values.pin_memory(), self.key_map.pin_memory(),
                        self.tok_offsets.pin_memory(), self.pinyin.pin_memory(), self.pinyin_map.pin_memory(),
                        self.pin_offsets.pin_memory())

    def tensors(self) -> List[torch.Tensor]:
        return [self.keys, self.values, self.key_map, self.tok_offsets, self.pinyin, self.pinyin_map, self.pin_offsets]

    @property
    def nbytes(self) -> int:
        n = self.keys.numel() * 4
        return n + (0 if self.values.data_ptr() == self.keys.data_ptr() else n) + self.key_map.numel() * 4

    def c_struct(self) -> "binding.DictBankStruct":
        if not self.keys.is_cuda:
            raise RuntimeError("the bank must be on the device first: bank.to('cuda:0')")
        if self._struct is None:
            p = lambda t: C.c_void_p(t.data_ptr())      # noqa: E731
            self._struct = binding.DictBankStruct(p(self.keys), p(self.values), p(self.key_map), p(self.tok_offsets),
                                                  p(self.pinyin), p(self.pinyin_map), p(self.pin_offsets),
                                                  self.n_entries)
        return self._struct


def ids_from_words(words: List[List[str]], word_to_id: Dict[str, int], Tw: int) -> torch.Tensor:
    """dict_ids for a batch of word lists as the test-set reader holds them (['<BOS>', c1, ..., '<EOS>']): entry
    index = word id in word_set.json order, unknown characters map to <UNK> = 2 (dataset_utils.py:312-315).  Rows are
    the ones the reference collater builds: BOS_EOS in column 0 and column Tw-1 of every utterance, PAD (all-zero row)
    everywhere else outside the characters -- see data.collate_dict_ids."""
    ids = torch.full((len(words), Tw), PAD, dtype=torch.long)
    ids[:, 0] = ids[:, Tw - 1] = BOS_EOS
    for b, ws in enumerate(words):
        for t, w in enumerate(ws[1:-1], start=1):
            ids[b, t] = word_to_id.get(w, 2)
    return ids
