import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf
ne = 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40.0, 0.4)
K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
for mode in ("0", "2", "3", "7", "0"):
    os.environ["SMFEM_BENCH_MODE"] = mode
    print(f"mode {mode}: {K.bench_spmv(reps=30, variant=4):.4f} ms")
