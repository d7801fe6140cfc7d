"""Summarises an .ncu-rep (read on the CPU box with `ncu -i`): headline metrics per launch and, with --source,
where the warps of the kernel spend their samples (top SASS lines).  Used to write profiles/*.md."""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "sm cycles"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed cycles"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "tensor pipe active %, _realtime flavour (varies between captures of the same launch: do not quote)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem->tensor-core wavefronts % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "bulk-copy (TMA) bytes L2->smem"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__cluster_dim_x", "cluster"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield {h: (v, u) for h, u, v in zip(hdr, units, r)}


def source(path, top=14):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for si, st in enumerate(starts):
        hdr = rows[st + 1]
        body = rows[st + 2: starts[si + 1] if si + 1 < len(starts) else None]
        ci = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ci["# Samples"]]) for r in body)
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        print(f"  samples {tot}, SASS instructions {len(body)}")
        for r in sorted(body, key=lambda r: -int(r[ci["# Samples"]]))[:top]:
            st2 = sorted(((int(r[ci[c]]), c) for c in stall_cols), reverse=True)[0]
            print("   %5.1f%%  x%-9s %-60s %s" % (100.0 * int(r[ci["# Samples"]]) / max(tot, 1),
                                                r[ci["Instructions Executed"]], r[1].strip()[:60], st2[1]))


if __name__ == "__main__":
    for path in [a for a in sys.argv[1:] if not a.startswith("--")]:
        print("==", path)
        for m in raw(path):
            print(" kernel:", m.get("Kernel Name", ("?", ""))[0][-60:])
            for k, label in KEYS:
                hit = [h for h in m if h.endswith(k)]
                if hit:
                    v, u = m[hit[0]]
                    print("  %-44s %s %s" % (label, v, u))
        if "--source" in sys.argv:
            source(path)
