"""Per-kernel resource usage of the built library (`cuobjdump --dump-resource-usage`, sm_100a cubins): registers, stack (spill)
bytes, static shared memory, sorted by register count.  No GPU needed; the output is committed as profiles/r2_resource_usage.txt."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "smearfem.jl_b200", "libsmearfem_b200.so")
out = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
rows, fn = [], None
for ln in out.splitlines():
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and fn:
        rows.append((int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4)), fn))
        fn = None
names = subprocess.run(["c++filt"], input="\n".join(r[4] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# {len(rows)} kernels in libsmearfem_b200.so (sm_100a); REG = registers per thread, STACK = bytes of local stack per thread (spills /")
print("# local arrays), SHARED = static shared memory (dynamic shared memory is set at launch: see SMEM_BYTES in the sources)")
print(f"{'REG':>4} {'STACK':>6} {'SHARED':>7}  kernel")
for (reg, stack, shared, _local, _), name in sorted(zip(rows, names), key=lambda t: (-t[0][0], t[1])):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    print(f"{reg:>4} {stack:>6} {shared:>7}  {name[:170]}")
