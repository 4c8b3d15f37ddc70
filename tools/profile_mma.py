"""Short single-GPU run for ncu: values-only assembly with the tile variants given on the command line."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mma84"]
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
for v in variants:
    os.environ["SMFEM_TILE"] = v
    for _ in range(2):
        K.assemble_values(40.0, 0.4)
