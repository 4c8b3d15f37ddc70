"""Time the layer-march kernel (SMFEM_TILE=v2) beside the first tile kernel, with the ablation masks of both
(SMFEM_TILE_SKIP: 1 phase 1, 2 sweeps / main loop, 4 combine, 8 output, 32 value stores collapsed) and a value check."""
import os
import sys

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


os.environ["SMFEM_TILE"] = "4x4"
K.reassemble(40.0, 0.4)
d_ref = K.diag()
v1 = run(lambda: K.assemble_values(40.0, 0.4))
f1 = run(lambda: K.reassemble(40.0, 0.4))
print(f"tile v1: values {v1:.3f} ms, fused {f1:.3f} ms", flush=True)
os.environ["SMFEM_TILE"] = "v2"
os.environ["SMFEM_DEBUG_CLEAR"] = "1"
K.reassemble(40.0, 0.4)
os.environ.pop("SMFEM_DEBUG_CLEAR")
d2 = K.diag()
print("diag rel diff v2 vs v1:", float(np.linalg.norm(d2 - d_ref) / np.linalg.norm(d_ref)), flush=True)
v2 = run(lambda: K.assemble_values(40.0, 0.4))
f2 = run(lambda: K.reassemble(40.0, 0.4))
print(f"tile v2: values {v2:.3f} ms, fused {f2:.3f} ms -> {ne**3 / f2 / 1e3:.1f} M el/s", flush=True)
for skip, name in [(1, "no phase1"), (2, "no sweeps"), (4, "no shuffles"), (8, "no output"), (14, "phase1 only"), (13, "sweeps only"),
                   (11, "combine only"), (7, "output only"), (15, "skeleton"), (32, "stores collapsed")]:
    os.environ["SMFEM_TILE_SKIP"] = str(skip)
    t = run(lambda: K.reassemble(40.0, 0.4), 5)
    print(f"v2 skip={skip:2d} {name:16s}: fused {t:.3f} ms", flush=True)
os.environ.pop("SMFEM_TILE_SKIP")
for variant in ("v2base", "v2l", "v2i", "v2li", "v2s", "v2g", "v2p", "v2", "v2e", "v2all", "v3", "v3b"):
    os.environ["SMFEM_TILE"] = variant
    K.reassemble(40.0, 0.4)
    ok = bool(np.array_equal(K.diag(), d2))
    v = run(lambda: K.assemble_values(40.0, 0.4))
    f = run(lambda: K.reassemble(40.0, 0.4))
    print(f"variant {variant:7s}: values {v:.3f} ms, fused {f:.3f} ms, same diag bits as v2: {ok}", flush=True)
for variant in ("v2base", "v2", "v2s"):
    os.environ["SMFEM_TILE"] = variant
    os.environ["SMFEM_COLIND_SIDE"] = "1"
    f = run(lambda: K.reassemble(40.0, 0.4))
    print(f"variant {variant:7s} + side-stream colind kernel: step {f:.3f} ms", flush=True)
os.environ.pop("SMFEM_COLIND_SIDE")
os.environ["SMFEM_TILE"] = "v2"
# occupancy sensitivity: one CTA (4 warps) per SM instead of two
os.environ["SMFEM_TILE_SMEM_PAD"] = "40000"
print(f"v2 with ONE resident CTA per SM (4 warps): values {run(lambda: K.assemble_values(40.0, 0.4)):.3f} ms, fused {run(lambda: K.reassemble(40.0, 0.4)):.3f} ms", flush=True)
os.environ["SMFEM_TILE"] = "v2s"
os.environ["SMFEM_TILE_SMEM_PAD"] = "20000"
print(f"v2s (64-thread CTAs) with TWO resident CTAs per SM (4 warps): fused {run(lambda: K.reassemble(40.0, 0.4)):.3f} ms", flush=True)
os.environ.pop("SMFEM_TILE_SMEM_PAD")
os.environ["SMFEM_TILE"] = "v2"
for plan in sys.argv[2:]:
    os.environ["SMFEM_TILE_CHUNKS"] = plan
    print(f"v2 chunks={plan}: fused {run(lambda: K.reassemble(40.0, 0.4)):.3f} ms", flush=True)
