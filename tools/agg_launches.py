"""Aggregates an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list by (kernel, grid) over the LAST pass of the hot path found in the log (a pass starts at `first_kernel`)."""
import csv
import sys


def main(path, first_kernel="embed"):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    per_id = {}
    order = []
    for x in csv.DictReader(lines):
        i = int(x["ID"])
        if i not in per_id:
            per_id[i] = dict(name=x["Kernel Name"].split("(")[0][-34:], grid=x["Grid Size"], t=0.0, rd=0.0, wr=0.0)
            order.append(i)
        v = float(x["Metric Value"].replace(",", ""))
        unit = x["Metric Unit"]
        m = x["Metric Name"]
        if m == "gpu__time_duration.sum":
            per_id[i]["t"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per_id[i]["rd" if "read" in m else "wr"] = v * scale
    rows = [per_id[i] for i in order]
    idx = [] if first_kernel == "ALL" else [i for i, x in enumerate(rows) if first_kernel in x["name"]]
    start = idx[-1] if idx else 0
    agg, tot = {}, 0.0
    for r in rows[start:]:
        tot += r["t"]
        a = agg.setdefault((r["name"], r["grid"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += r["t"]
        a[2] += r["rd"]
        a[3] += r["wr"]
    print("launches %d  total %.1f us (first kernel of the pass: %s)" % (len(rows) - start, tot, rows[start]["name"]))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-36s %-16s n=%-3d total %9.1f us  avg %8.1f us  %5.1f%%  dram rd %8.1f MB wr %8.1f MB" % (
            k[0], k[1], v[0], v[1], v[1] / v[0], 100 * v[1] / tot, v[2] / 1e6, v[3] / 1e6))


if __name__ == "__main__":
    main(*sys.argv[1:])
