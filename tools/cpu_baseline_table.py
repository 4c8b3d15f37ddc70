"""the C restatement of the reference algorithm (element loop -> COO triplets -> sparse(), src/fem.jl:135-256)
timed on the host cores at 20^3, 50^3 and 100^3, single-threaded (the reference's execution model) and with OpenMP on all cores.
Prints one JSON line per (size, threads).  CPU only (test infrastructure: oracle/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import c_oracle, fem_oracle as o

sizes = [int(a) for a in sys.argv[1:]] or [20, 50, 100]
allc = c_oracle.max_threads()
model = ""
try:
    model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass
for ne in sizes:
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    for th in (allc, 1):
        t = time.perf_counter()
        K = c_oracle.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=th)
        dt = time.perf_counter() - t
        assert K.nnz == 9 * (3 * (ne + 1) - 2) ** 3
        print(json.dumps({"ne": ne, "elements": ne**3, "threads": th, "seconds": dt, "elements_per_s": ne**3 / dt, "nnz": int(K.nnz),
                          "coo_bytes": 576 * 24 * ne**3, "cpu": model, "cores_available": allc,
                          "what": "restated reference algorithm (C port), not a Julia measurement"}), flush=True)
        del K
