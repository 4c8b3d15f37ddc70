"""Ablation of the tile kernel's phases (profiling aid; results are wrong when something is skipped)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
os.environ.setdefault("SMFEM_TILE", "4x4")  # this script ablates the first tile kernel; tools/time_tile2.py the layer-march one
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
for skip, name in [(0, "full"), (1, "no phase1"), (2, "no main loop"), (4, "no combine"), (8, "no output"), (14, "phase1 only"), (13, "main only"), (11, "combine only"), (7, "output only"), (15, "skeleton")]:
    os.environ["SMFEM_TILE_SKIP"] = str(skip)
    for _ in range(2):
        K.assemble_values(40.0, 0.4)
    ctx.timer_start()
    for _ in range(5):
        K.assemble_values(40.0, 0.4)
    print(f"skip={skip:2d} {name:14s}: {ctx.timer_stop()/5:.3f} ms")
# store ablation: 16 / 32 collapse the column-index / value stores onto a small cache-resident window (same instructions, no DRAM)
for skip in (0, 16, 32, 48, 8):
    os.environ["SMFEM_TILE_SKIP"] = str(skip)
    for _ in range(2):
        K.reassemble(40.0, 0.4)
    ctx.timer_start()
    for _ in range(5):
        K.reassemble(40.0, 0.4)
    f = ctx.timer_stop() / 5
    ctx.timer_start()
    for _ in range(5):
        K.assemble_values(40.0, 0.4)
    print(f"skip={skip:2d}: fused {f:.3f} ms, values {ctx.timer_stop()/5:.3f} ms")
