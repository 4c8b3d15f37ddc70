"""Time the tile kernel for output modes / chunk plans (CUDA events) and check each variant bit-for-bit against mode 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smearfem_b200 as sf
ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)


def run(fn, n=10):
    for _ in range(3):
        fn()
    ctx.timer_start()
    for _ in range(n):
        fn()
    return ctx.timer_stop() / n


def csc():
    os.environ["SMFEM_DEBUG_CLEAR"] = "1"  # every buffer is set to 0xFF first: proves each entry is rewritten
    K.reassemble(40.0, 0.4)
    os.environ.pop("SMFEM_DEBUG_CLEAR")
    return K.to_csc()


ref = None
chunk_plans = [None] + sys.argv[2:]
for plan in chunk_plans:
    if plan is None:
        os.environ.pop("SMFEM_TILE_CHUNKS", None)
    else:
        os.environ["SMFEM_TILE_CHUNKS"] = plan
    for out in ("0", "2", "3"):
        os.environ["SMFEM_TILE_OUT"] = out
        v = run(lambda: K.assemble_values(40.0, 0.4))
        f = run(lambda: K.reassemble(40.0, 0.4))
        line = f"chunks={plan or 'auto':>24s} out={out}: values {v:.3f} ms, fused {f:.3f} ms -> {ne**3/f/1e3:.1f} M el/s"
        if ne <= 64 or plan is None:
            got = csc()
            if ref is None:
                ref = got
            same = all(np.array_equal(a, b) for a, b in zip(ref, got))
            line += f"  bit-identical to first: {same}"
        print(line, flush=True)
