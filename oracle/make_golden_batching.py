"""TEST INFRASTRUCTURE ONLY -- run in the build container (needs /root/reference).

Runs the reference's own ``utils.batch_by_size`` (utils/__init__.py:180-234) and the ``x[rank::world]`` dealing of
``build_dataloader`` (tasks/tts/tts_base.py:113-155) on seeded length lists and writes tests/golden/batching.json.

    python -m oracle.make_golden_batching
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

CASES = [
    # seed, n, max_len, max_tokens, max_sentences, multiple
    (1, 50, 400, 2000, None, 1),
    (2, 97, 900, 6000, 12, 1),
    (3, 64, 500, 4000, 8, 4),
    (4, 33, 300, None, 5, 2),
    (5, 120, 1500, 28000, 60, 8),
    (6, 7, 100, 100, 3, 2),
    (7, 1, 50, 64, 1, 1),
]


def lengths(seed, n, max_len):
    return np.random.RandomState(seed).randint(max(1, max_len // 8), max_len + 1, size=n).tolist()


def main():
    ref_loader.load()
    import utils as ref_utils                                   # the reference's utils package (sys.path set by load())
    out = []
    for seed, n, max_len, mt, ms, mult in CASES:
        sizes = lengths(seed, n, max_len)
        order = np.arange(n)
        batches = ref_utils.batch_by_size(order, lambda i: sizes[i], max_tokens=mt, max_sentences=ms,
                                          required_batch_size_multiple=mult)
        rec = dict(seed=seed, n=n, max_len=max_len, max_tokens=mt, max_sentences=ms, multiple=mult,
                   batches=[[int(i) for i in b] for b in batches])
        # DDP dealing exactly as tts_base.py:148-151 does it, for world = 2
        rec["dealt_w2"] = [[[int(i) for i in x[r::2]] for x in batches if len(x) % 2 == 0] for r in range(2)]
        out.append(rec)
    path = os.path.join(ROOT, "tests", "golden", "batching.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, sum(len(r["batches"]) for r in out), "batches")


if __name__ == "__main__":
    main()
