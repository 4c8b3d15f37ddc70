"""Time the assembly kernels for the tile variants (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
for tile in ("4x4", "8x4"):
    os.environ["SMFEM_TILE"] = tile
    for _ in range(3):
        K.assemble_values(40.0, 0.4)
    ctx.timer_start()
    for _ in range(10):
        K.assemble_values(40.0, 0.4)
    ms = ctx.timer_stop() / 10
    print(f"tile {tile}: values {ms:.3f} ms  -> {ne**3/ms/1e3:.1f} M el/s")
ctx.timer_start()
for _ in range(10):
    K.pattern_rebuild()
print(f"pattern {ctx.timer_stop()/10:.3f} ms")
for tile in ("4x4", "8x4"):
    os.environ["SMFEM_TILE"] = tile
    for _ in range(3):
        K.reassemble(40.0, 0.4)
    ctx.timer_start()
    for _ in range(10):
        K.reassemble(40.0, 0.4)
    ms = ctx.timer_stop() / 10
    print(f"tile {tile}: fused pattern+values {ms:.3f} ms  -> {ne**3/ms/1e3:.1f} M el/s")
