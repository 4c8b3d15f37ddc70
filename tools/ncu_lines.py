"""Join an `ncu --page source --csv` (SASS view) export with `nvdisasm -g` line info: stall samples / instructions per CUDA source line.
usage: ncu_lines.py <src.csv> <nvdisasm -g output> <kernel index> [top]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, kidx = sys.argv[1], sys.argv[2], int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of = {}
cur = None
for ln in open(sass):
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(src_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
sec = rows[starts[kidx]:starts[kidx + 1]]
hdr = sec[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
base = None
per = defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r in sec[2:]:
    if len(r) <= iex or not r[ia]:
        continue
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    key = line_of.get(a - base, ((None, 0), ""))[0]
    vals = [int(float(r[isamp] or 0)), int(float(r[iex] or 0)), int(float(r[iw] or 0)) if iw is not None else 0]
    for j in range(3):
        per[key][j] += vals[j]
        tot[j] += vals[j]
print("kernel:", sec[0][1][:80], "total samples", tot[0], "instr", tot[1], "smem wavefronts", tot[2])
for key, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):40s} samples {v[0]:7d} ({100*v[0]/max(tot[0],1):5.1f}%)  instr {v[1]:11d} ({100*v[1]/max(tot[1],1):5.1f}%)  smem_wf {v[2]:11d}")
