"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import csv
import sys


def main(path, last_pass_marker="to_planes"):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = [(x["Kernel Name"].split("(")[0][-34:], x["Grid Size"], float(x["Metric Value"]) / 1e3)
            for x in csv.DictReader(lines)]
    idx = [i for i, x in enumerate(rows) if last_pass_marker in x[0]]
    start = idx[-1] - 1 if idx else 0
    agg, tot = {}, 0.0
    for n, g, t in rows[start:]:
        tot += t
        a = agg.setdefault((n, g), [0, 0.0])
        a[0] += 1
        a[1] += t
    print("launches %d  total %.1f us" % (len(rows) - start, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-36s %-16s n=%-3d total %9.1f us  avg %8.1f us  %5.1f%%" % (k[0], k[1], v[0], v[1], v[1] / v[0],
                                                                          100 * v[1] / tot))


if __name__ == "__main__":
    main(*sys.argv[1:])
