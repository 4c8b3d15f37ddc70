"""DMMA value kernel (SMFEM_TILE=mma*) against the scalar tile kernel: parity on small meshes, then timing at ne (default 100)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

VARIANTS = sys.argv[2].split(",") if len(sys.argv) > 2 else ["mma75", "mma84", "mma44"]
ne_big = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


ok = True
for ne in (3, 13, 20):
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    os.environ.pop("SMFEM_TILE", None)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    cp0, rv0, nz0 = K.to_csc()
    d0 = K.diag()
    for v in VARIANTS:
        for chunks in (None, "2,1," + str(ne + 1 - 3)) if ne > 3 else (None,):
            os.environ["SMFEM_TILE"] = v
            os.environ["SMFEM_DEBUG_CLEAR"] = "1"
            if chunks:
                os.environ["SMFEM_TILE_CHUNKS"] = chunks
            K.reassemble(40, 0.4)
            cp, rv, nz = K.to_csc()
            d = K.diag()
            K.reassemble(40, 0.4)
            nz2 = K.to_csc()[2]
            os.environ.pop("SMFEM_TILE_CHUNKS", None)
            os.environ.pop("SMFEM_DEBUG_CLEAR", None)
            pat = np.array_equal(cp, cp0) and np.array_equal(rv, rv0)
            finite = bool(np.all(np.isfinite(nz)))
            rv_ = rel(nz, nz0) if finite else float("nan")
            det = np.array_equal(nz, nz2)
            good = pat and finite and rv_ <= 1e-13 and det and rel(d, d0) <= 1e-13
            ok &= good
            print(f"ne={ne:3d} {v:9s} chunks={chunks}: pattern={pat} finite={finite} relK={rv_:.2e} reldiag={rel(d, d0):.2e} deterministic={det}"
                  f" maxabs={np.max(np.abs(nz - nz0)):.2e}  {'OK' if good else 'FAIL'}", flush=True)
    os.environ.pop("SMFEM_TILE", None)
    del K, mesh
print("PARITY", "OK" if ok else "FAIL", flush=True)

ne = ne_big
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
for v in ["4x4"] + VARIANTS:
    os.environ["SMFEM_TILE"] = v
    for fused in (False, True):
        f = (lambda: K.reassemble(40.0, 0.4)) if fused else (lambda: K.assemble_values(40.0, 0.4))
        for _ in range(3):
            f()
        ctx.timer_start()
        for _ in range(10):
            f()
        ms = ctx.timer_stop() / 10
        print(f"ne={ne} tile {v:9s} {'fused ' if fused else 'values'}: {ms:.3f} ms -> {ne**3 / ms / 1e3:.1f} M el/s", flush=True)
if len(sys.argv) > 3:  # ablation of the first variant
    os.environ["SMFEM_TILE"] = VARIANTS[0]
    for skip in (0, 1, 2, 4, 8, 3, 7, 15):
        os.environ["SMFEM_TILE_SKIP"] = str(skip)
        for _ in range(2):
            K.assemble_values(40.0, 0.4)
        ctx.timer_start()
        for _ in range(5):
            K.assemble_values(40.0, 0.4)
        print(f"skip={skip:2d}: {ctx.timer_stop() / 5:.3f} ms", flush=True)
    os.environ.pop("SMFEM_TILE_SKIP")
