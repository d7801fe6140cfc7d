"""Aggregate the warp-stall samples of an ncu capture by SOURCE LINE (developer tool).

  python tools/ncu_lines.py <report.ncu-rep> <kernel-name-substring> [top]

The SASS page of the report gives samples per instruction address; nvdisasm --print-line-info on the cubin of the same
build (extracted from dict_tts_b200/libdtts.so) maps addresses to file:line (inlined helpers are attributed to the line
of the helper).  The build must be the one that was profiled."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(here, "dict_tts_b200", "libdtts.so")], cwd=tmp, check=True,
               stdout=subprocess.DEVNULL)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, f)], capture_output=True,
                         text=True).stdout
    cur_line, inside, marks = None, False, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            inside = kern in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            marks.append(m.groups())
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if marks:
                # innermost location inside the kernel's own file: the call site of the first helper frame, else the line
                main = os.path.basename(marks[-1][0])
                loc = None
                for f0, l0, f1, l1 in marks:
                    if os.path.basename(f0) != main and f1 and os.path.basename(f1) == main:
                        loc = (main, int(l1))
                        break
                if loc is None:
                    f0, l0 = next(((a, b) for a, b, _, _ in marks if os.path.basename(a) == main), marks[0][:2])
                    loc = (os.path.basename(f0), int(l0))
                cur_line = loc
                marks = []
            if cur_line:
                addr2line[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]
ia, isrc, isamp = H.index("Address"), H.index("Source"), H.index("# Samples")
base = None
agg = collections.Counter()
ins = {}
total = 0
for r in rows[hi + 1:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") or not r[ia].isdigit() else int(r[ia])
        n = int(r[isamp])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    off = a - base
    key = addr2line.get(off, (("?", 0), ""))[0]
    agg[key] += n
    total += n
    if n:
        ins.setdefault(key, []).append((n, r[isrc][:70]))
print("total samples", total, " mapped instructions", len(addr2line))
for key, n in agg.most_common(top):
    print("%6d %5.1f%%  %s:%d" % (n, 100.0 * n / max(total, 1), key[0], key[1]))
    for m, s in sorted(ins.get(key, []), reverse=True)[:2]:
        print("           %6d  %s" % (m, s))
