"""Short PCG run for ncu launch lists: assemble ne^3, 1 solve capped at maxit iterations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40.0, 0.4)
K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
q, it, rel = K.pcg_solve(rtol=1e-10, maxit=maxit, want_q=False)
st = K.pcg_stats()
print(f"pcg: {it} iters, {st['ms_total']:.2f} ms total, {st['ms_total']/max(it,1):.4f} ms/iter, relres {rel:.2e}")
