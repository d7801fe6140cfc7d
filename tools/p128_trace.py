"""Timeline of rb_pair128_kernel's pipeline (developer tool): runs one valid-length vocode pass on the trace build of the
library (make -C dict_tts_b200/csrc trace) and prints, for CTA 0 and the launch selected by --k / --dil, when the MMA
thread and epilogue warp 3 passed each pipeline event (microseconds since the first event, clock64 / --ghz)."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dict_tts_b200.binding as binding  # noqa: E402

binding.LIB_PATH = os.path.join(ROOT, "dict_tts_b200", "libdtts_trace.so")
from dict_tts_b200 import synth  # noqa: E402
from dict_tts_b200.engine import HifiGanEngine  # noqa: E402

EV = ["c1.enter", "c1.acc1empty", "c1.issued", "c2.enter", "c2.tfull", "c2.acc2empty", "c2.issued", "e1a.acc1full",
      "e1b.enter", "e1b.tempty", "e1b.done", "e2.acc2full", "e2.acc2released", "e2.done", "c1.wait_A", "c1.wait_W",
      "c2.wait_W", "w.slot_free", "w.own_full", "w.peer_full"]

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=7)
    ap.add_argument("--dil", type=int, default=3)
    ap.add_argument("--ghz", type=float, default=1.7)
    ap.add_argument("--tiles", type=int, default=8)
    a = ap.parse_args()
    lib = binding.load()
    lib.dtts_debug_p128_trace_select.argtypes = [C.c_int, C.c_int]
    lib.dtts_debug_p128_trace.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    eng = HifiGanEngine(synth.make_vocoder_state_dict(4321), precision=6)
    mel = synth.make_mel(7, 60, 400).cuda()
    ml = synth.make_batch(seed=1234, B=60, min_chars=12, max_chars=20, max_frames=400, Lk_cap=8)["mel_lengths"]
    lens = ml.int().cuda()
    eng(mel, lens)
    assert lib.dtts_debug_p128_trace_select(a.k, a.dil) == 0
    eng(mel, lens)
    torch.cuda.synchronize()
    n_ev, n_t = 20, 24
    buf = (C.c_longlong * (n_ev * n_t))()
    assert lib.dtts_debug_p128_trace(buf, n_ev * n_t) == 0
    v = [[buf[t * n_ev + e] for e in range(n_ev)] for t in range(n_t)]
    t0 = min(x for row in v for x in row[:14] if x > 0)
    us = lambda c: (c - t0) / (a.ghz * 1e3)
    print("k=%d dil=%d, CTA 0, microseconds at %.2f GHz" % (a.k, a.dil, a.ghz))
    print("tile " + " ".join("%15s" % e for e in EV[:17]))
    for t in range(a.tiles):
        row = v[t]
        print("%4d " % t + " ".join(("%15.2f" % us(row[e])) if e < 14 else ("%15.2f" % (row[e] / (a.ghz * 1e3)))
                                     for e in range(17)))
    # steady-state summary over tiles 2..: period, MMA-thread waits
    per = [(v[t + 1][6] - v[t][6]) / (a.ghz * 1e3) for t in range(2, a.tiles + 6)]
    print("tile period (c2.issued to c2.issued), tiles 2..: " + " ".join("%.1f" % x for x in per))
    for name, f in (("c1: wait acc1empty", lambda r: r[1] - r[0]), ("c1: issue span", lambda r: r[2] - r[1]),
                    ("c1: of which wait A", lambda r: r[14]), ("c1: of which wait W", lambda r: r[15]),
                    ("c2: wait tfull", lambda r: r[4] - r[3]), ("c2: wait acc2empty", lambda r: r[5] - r[4]),
                    ("c2: issue span", lambda r: r[6] - r[5]), ("c2: of which wait W", lambda r: r[16]),
                    ("e1a: acc1full -> e1b.enter", lambda r: r[8] - r[7]), ("e1b: wait tempty", lambda r: r[9] - r[8]),
                    ("e1b: store", lambda r: r[10] - r[9]), ("e2: acc2full -> released", lambda r: r[12] - r[11]),
                    ("e2: acc2full -> done", lambda r: r[13] - r[11]),
                    ("W stage (conv2, chunk 2, first): slot seen free -> own half landed", lambda r: r[18] - r[17]),
                    ("   own half seen -> peer's half relayed", lambda r: r[19] - r[18]),
                    ("   c2 issue start -> slot seen free by the producer", lambda r: r[17] - r[5]),
                    ("(DTTS_P128_COPYLAT=1 only) weight copy issued -> landed", lambda r: r[13] if os.environ.get("DTTS_P128_COPYLAT") else 0)):
        xs = [f(v[t]) / (a.ghz * 1e3) for t in range(2, a.tiles + 6)]
        print("%-30s mean %6.2f us   " % (name, sum(xs) / len(xs)) + " ".join("%.1f" % x for x in xs[:10]))
