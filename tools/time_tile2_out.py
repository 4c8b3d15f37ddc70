"""Output-path experiments for the layer-march kernel: ablation masks with / without the column indices."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
os.environ["SMFEM_TILE"] = sys.argv[2] if len(sys.argv) > 2 else "v2"


def run(fn, n=8):
    for _ in range(2):
        fn()
    ctx.timer_start()
    for _ in range(n):
        fn()
    return ctx.timer_stop() / n


for skip in (0, 7, 8, 15, 64 + 7, 128 + 7, 64 + 128 + 7, 64, 128):
    os.environ["SMFEM_TILE_SKIP"] = str(skip)
    v = run(lambda: K.assemble_values(40.0, 0.4))
    f = run(lambda: K.reassemble(40.0, 0.4))
    print(f"skip={skip:3d}: values {v:.3f} ms  fused {f:.3f} ms", flush=True)
