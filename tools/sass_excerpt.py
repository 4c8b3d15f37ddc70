"""SASS evidence for profiles/: per hot kernel the opcode histogram of `cuobjdump -sass` on the built library (sm_100a cubins
only), with the Blackwell-era mnemonics called out: UBLKCP (cp.async.bulk), LDGSTS (cp.async), DFMA / DMMA (fp64), REDG, SYNCS."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "smearfem.jl_b200", "libsmearfem_b200.so")
want = sys.argv[1:] or ["k_values_tile2", "k_values_tile", "k_values_mma", "k_spmv_group", "k_spmv_tma", "k_spmv_stream", "k_matfree_color2",
                        "k_struct_colind_side", "k_pcg_update", "k_border"]
print("# cuobjdump -lelf:", ", ".join(l.split(":")[-1].strip() for l in subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout.splitlines()))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, per = None, collections.defaultdict(collections.Counter)
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and cur:
        ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(1).strip())
        per[cur][ins.split()[0]] += 1
special = ("UBLKCP", "LDGSTS", "DFMA", "DMMA", "REDG", "SYNCS", "UTMA", "STG", "LDS", "STS", "SHFL", "FSEL")
for key in want:
    for fn, c in sorted(per.items()):
        if key not in fn:
            continue
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:150]
        tot = sum(c.values())
        fam = collections.Counter()
        for op, n in c.items():
            for s in special:
                if op.startswith(s):
                    fam[s] += n
        print(f"\n== {demangled}\n   {tot} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in fam.most_common()))
        print("   top opcodes: " + ", ".join(f"{op} {n}" for op, n in c.most_common(12)))
