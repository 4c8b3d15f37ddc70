"""Short single-GPU run for ncu: one assembly + a few SpMVs (per variant) at ne (default 100)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
for _ in range(2):
    K.pattern_rebuild()
    K.assemble_values(40.0, 0.4)   # k_values_tile, values only
    K.reassemble(40.0, 0.4)        # k_struct_rowptr + k_values_tile writing colind too (fused assembly step)
K.add_surface_mass(100.0)
i = K.info()
b = 12 * i["nnz_local"] + 24 * i["nrows_local"]
for v in (4, 3, 2, 1, 0):
    ms = K.bench_spmv(reps=reps, variant=v)
    print(f"spmv variant {v}: {ms:.4f} ms  {b / ms / 1e6:.1f} GB/s")
