"""Short run for ncu: a few applications of the matrix-free operator at ne (default 100)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
K.use_matrix_free(True).use_matrix_free(False)
print(K.bench_spmv(reps=2, variant=5))
