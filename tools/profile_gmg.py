"""Short single-GPU run for ncu: one multigrid-preconditioned CG solve of the example problem at ne (default 100)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
K.use_multigrid(True)
q, it, rel = K.pcg_solve(rtol=1e-10, maxit=500)
print(f"multigrid-PCG ne={ne}: {it} iterations, relres {rel:.1e}, {K.pcg_stats()['ms_total']:.1f} ms (under ncu: not a timing)")
