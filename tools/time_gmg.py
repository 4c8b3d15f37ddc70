"""Jacobi-PCG vs multigrid-PCG on the example problem at ne (default 100), one GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rtol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-10
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
qj, itj, relj = K.pcg_solve(rtol=rtol, maxit=20000)
qj, itj, relj = K.pcg_solve(rtol=rtol, maxit=20000)
msj = K.pcg_stats()["ms_total"]
K.use_multigrid(True)
qg, itg, relg = K.pcg_solve(rtol=rtol, maxit=500)   # builds the hierarchy
ms_first = K.pcg_stats()["ms_total"]
qg, itg, relg = K.pcg_solve(rtol=rtol, maxit=500)
msg = K.pcg_stats()["ms_total"]
d = np.linalg.norm(qg - qj) / np.linalg.norm(qj)
print(f"ne={ne} rtol={rtol:g}: Jacobi-PCG {itj} iterations {msj:.1f} ms (relres {relj:.1e}) | multigrid-PCG {itg} iterations {msg:.1f} ms "
      f"(first call incl. hierarchy {ms_first:.1f} ms, relres {relg:.1e}) | speed-up {msj / msg:.1f}x | ||q_mg - q_j||/||q_j|| = {d:.1e}")
