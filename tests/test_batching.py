"""CPU suite for the input pipeline: the sampler mirror against golden outputs of the reference's own
``utils.batch_by_size`` (tests/golden/batching.json, written by oracle/make_golden_batching.py), and the ragged collate
(batch-local dictionary bank) against the padded collate."""
import json
import os

import numpy as np
import pytest
import torch

from tests import fake_exp
from dict_tts_b200.batching import batch_by_size, build_batch_sampler, ordered_indices
from dict_tts_b200.data import DictTTSTestSet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "batching.json")) as f:
    GOLDEN = json.load(f)


def lengths(seed, n, max_len):          # same generator as oracle/make_golden_batching.py
    return np.random.RandomState(seed).randint(max(1, max_len // 8), max_len + 1, size=n).tolist()


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: "seed%d" % c["seed"])
def test_batch_by_size_matches_reference(case):
    sizes = lengths(case["seed"], case["n"], case["max_len"])
    got = batch_by_size(np.arange(case["n"]), lambda i: sizes[i], case["max_tokens"], case["max_sentences"],
                        case["multiple"])
    assert [[int(i) for i in b] for b in got] == case["batches"]
    # invariants the reference guarantees: order kept, nothing lost, limits respected
    assert [i for b in got for i in b] == list(range(case["n"]))
    for b in got:
        if case["max_sentences"]:
            assert len(b) <= case["max_sentences"]
        if case["max_tokens"]:
            assert len(b) * max(sizes[i] for i in b) <= case["max_tokens"]


@pytest.mark.parametrize("case", GOLDEN, ids=lambda c: "seed%d" % c["seed"])
def test_rank_dealing_matches_reference(case):
    sizes = lengths(case["seed"], case["n"], case["max_len"])
    for rank in range(2):
        # the sampler scales per-device limits by the world size: hand it the per-device halves of the golden limits
        # only when they divide; otherwise check the dealing on the golden global batches directly
        dealt = [b[rank::2] for b in case["batches"] if len(b) % 2 == 0]
        assert dealt == case["dealt_w2"][rank]
    mt, ms = case["max_tokens"], case["max_sentences"]
    if (mt is None or mt % 2 == 0) and (ms is None or ms % 2 == 0):
        for rank in range(2):
            got = build_batch_sampler(sizes, max_tokens=None if mt is None else mt // 2,
                                      max_sentences=None if ms is None else ms // 2, world=2, rank=rank,
                                      required_batch_size_multiple=case["multiple"])
            assert got == [b for b in case["dealt_w2"][rank] if b]


def test_oversized_item_is_an_error():
    with pytest.raises(AssertionError):
        batch_by_size(range(3), lambda i: [10, 500, 10][i], max_tokens=100)


def test_ordered_indices():
    sizes = [5, 3, 9, 3, 7]
    assert ordered_indices(sizes).tolist() == [0, 1, 2, 3, 4]                   # test set: natural order
    order = ordered_indices(sizes, shuffle=True, rng=np.random.RandomState(0))
    assert sorted(order.tolist()) == [0, 1, 2, 3, 4]
    assert [sizes[i] for i in order] == sorted(sizes)                             # shuffled, then stable sort by length
    # same generator state -> same order as the reference's two numpy calls
    rs = np.random.RandomState(0)
    perm = rs.permutation(5)
    assert order.tolist() == perm[np.argsort(np.array(sizes)[perm], kind="mergesort")].tolist()


def test_fixed_groups_and_kept_remainder():
    sizes = list(range(10, 17))                                                  # 7 utterances
    ref_like = [build_batch_sampler(sizes, max_sentences=2, by_size=False, world=2, rank=r) for r in range(2)]
    assert ref_like == [[[0, 2]], [[1, 3]]]                                       # the 3-item tail group is dropped
    kept = [build_batch_sampler(sizes, max_sentences=2, by_size=False, world=2, rank=r, drop_ragged=False)
            for r in range(2)]
    assert kept == [[[0, 2], [4, 6]], [[1, 3], [5]]]
    assert sorted(i for r in kept for b in r for i in b) == list(range(7))


@pytest.fixture(scope="module")
def exp(tmp_path_factory):
    return fake_exp.write(str(tmp_path_factory.mktemp("exp_ragged")), n_items=6)


def test_ragged_collate_equals_padded_collate(exp):
    padded = next(DictTTSTestSet(exp["hparams"]).batches(max_sentences=4))
    ds = DictTTSTestSet(exp["hparams"])
    ds.ragged = True
    ragged = next(ds.batches(max_sentences=4))
    assert "keys" not in ragged and ragged["item_name"] == padded["item_name"]
    bank, ids = ragged["dict_bank"], ragged["dict_ids"]
    assert torch.equal(ragged["word_tokens"], padded["word_tokens"]) and ids.shape == padded["word_tokens"].shape
    # id rows name what the reference collater builds: its added row (-1) in column 0 and column Tw-1 of EVERY utterance,
    # all-zero rows (-2) at the EOS position and the padding of a shorter one
    assert (ids[:, 0] == -1).all() and (ids[:, -1] == -1).all()
    assert ((ids[:, 1:-1] >= 0) == (padded["word_tokens"][:, 1:-1] > 1)).all()
    # distinct characters are stored once
    assert bank.n_entries == len(set(ids[ids >= 0].tolist())) <= int((ids >= 0).sum())
    Lk, Lp = padded["key_map"].shape[2], padded["pinyin"].shape[2]
    assert bank.batch_dims(ids) == (Lk, Lp)
    again = bank.collate(ids)
    for k in ("keys", "values", "key_map", "pinyin", "pinyin_map"):
        assert torch.equal(again[k], padded[k]), k
    real = sum(t.numel() * t.element_size() for t in bank.tensors()[:1]) + ids.numel() * 8
    assert real < padded["keys"].numel() * 4                                     # fewer bytes than ONE padded tensor


def test_reference_dealing_of_the_test_set(exp):
    ds = DictTTSTestSet(exp["hparams"])
    r0 = [b["item_name"] for b in ds.batches(2, rank=0, world=2, deal="reference")]
    r1 = [b["item_name"] for b in ds.batches(2, rank=1, world=2, deal="reference")]
    flat0, flat1 = [n for b in r0 for n in b], [n for b in r1 for n in b]
    assert flat0 == ["fake_%03d" % i for i in range(0, 6, 2)] and flat1 == ["fake_%03d" % i for i in range(1, 6, 2)]
    with pytest.raises(ValueError):
        next(ds.batches(2, deal="nope"))


def test_value_that_is_the_key_stays_one_tensor(tmp_path):
    """The reference binarizer stores one feature tensor as both 'key' and 'value'; pickle keeps that identity, the reader
    and the collate keep it, so the pipeline uploads it once (values_dev may alias keys_dev in the C ABI)."""
    one = fake_exp.write(str(tmp_path / "one"), n_items=4, same_key_value=True)
    two = fake_exp.write(str(tmp_path / "two"), n_items=4, same_key_value=False)
    b1 = next(DictTTSTestSet(one["hparams"]).batches(max_sentences=4))
    b2 = next(DictTTSTestSet(two["hparams"]).batches(max_sentences=4))
    assert b1["values"] is b1["keys"]
    assert b2["values"] is not b2["keys"] and torch.equal(b2["values"], b2["keys"])
    for k in ("keys", "key_map", "pinyin", "pinyin_map", "word_tokens"):
        assert torch.equal(b1[k], b2[k]), k
