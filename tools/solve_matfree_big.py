"""One GPU, no assembled matrix: the example problem on an ne^3 lattice through smfem_matfree_operator + multigrid-PCG.
usage: solve_matfree_big.py [ne ...]   (default 200 300)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import smearfem_b200 as sf

ctx = sf.context()
for ne in [int(a) for a in sys.argv[1:]] or [200, 300]:
    n1 = ne + 1
    t0 = time.perf_counter()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    F = sf.SparseMatrixB200.matrix_free(ctx, mesh, 40, 0.4).add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    F.use_multigrid(True)
    ctx.sync()
    t1 = time.perf_counter()
    _, it, rr = F.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
    ms_first = F.pcg_stats()["ms_total"]
    _, it2, rr2 = F.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
    ms = F.pcg_stats()["ms_total"]
    free, total = torch.cuda.mem_get_info()
    nnz = 9 * (3 * n1 - 2) ** 3
    print(f"ne={ne}: {ne**3 / 1e6:.1f} M elements, {3 * n1**3 / 1e6:.1f} M dofs; assembled K would take {12 * nnz / 1e9:.1f} GB; "
          f"set-up {1e3 * (t1 - t0):.0f} ms; multigrid-PCG {it2} iterations, {ms:.0f} ms (first solve incl. hierarchy {ms_first:.0f} ms), relres {rr2:.1e}; "
          f"device memory in use {(total - free) / 1e9:.1f} GB", flush=True)
    F.free()
    mesh.free()
