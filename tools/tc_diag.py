"""Quick GPU diagnostic for the tcgen05 convolution: prints the max error per shape (no asserts)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.test_gpu_tensorcore import TC_CONV_SHAPES, _run_tc_conv  # noqa: E402

if __name__ == "__main__":
    print("variant", os.environ.get("DTTS_TC_VARIANT", "0"), flush=True)
    which = [int(a) for a in sys.argv[1:]] or range(len(TC_CONV_SHAPES))
    for i in which:
        shape = TC_CONV_SHAPES[i]
        try:
            out, act, ref = _run_tc_conv(shape, int(os.environ.get("DTTS_TC_PRECISION", "1")), False)
            err = (out - ref).abs()
            print(i, shape, "max err %.3e  mean err %.3e  ref max %.2f  nan %d" %
                  (err.max().item(), err.mean().item(), ref.abs().max().item(), int(torch.isnan(out).sum())), flush=True)
        except Exception as e:  # noqa: BLE001
            print(i, shape, "FAILED:", e, flush=True)
            break
