"""Stall samples of an ncu SASS-source export grouped by line ranges of one source file.
usage: ncu_regions.py <src.csv> <nvdisasm -g output of the kernel> <section index> <file> name:lo-hi ...
(`ncu --page source --csv` prints TWO sections per launch: launch i is section 2 i)"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kidx, fname = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
regions = []
for a in sys.argv[5:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
line_of = {}
cur = None
for ln in open(sass):
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
sec = rows[starts[kidx]:starts[kidx + 1]]
hdr = sec[1]
cols = ["stall_wait", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_selected", "stall_not_selected", "stall_math", "stall_dispatch", "stall_mio",
        "stall_branch_resolving", "stall_no_inst"]
idx = {c: hdr.index(c) for c in cols}
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")


def region(l):
    if l is None:
        return "?"
    f, n = l
    if f != fname:
        return f
    for name, lo, hi in regions:
        if lo <= n <= hi:
            return name
    return "other"


per = defaultdict(lambda: defaultdict(int))
base = None
tot = 0
for r in sec[2:]:
    if not r or not r[ia]:
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = region(line_of.get(a - base))
    for c in cols:
        per[key][c] += int(float(r[idx[c]] or 0))
    per[key]["instr"] += int(float(r[ie] or 0))
    per[key]["samples"] += int(float(r[isamp] or 0))
    tot += int(float(r[isamp] or 0))
for k, v in sorted(per.items(), key=lambda kv: -kv[1]["samples"]):
    print(f"{str(k):22s} {100*v['samples']/tot:5.1f}% samples {v['samples']:6d} instr {v['instr']/1e6:7.1f}M | " + " ".join(f"{c[6:]}={v[c]}" for c in cols))
