"""General (unstructured) path timing: 64^3 lattice with shuffled element order, standard vs permuted dof map."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smearfem_b200 as sf
from oracle import fem_oracle as o

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
o.inflate_sphere(NL, 0, 1, 0, 1)
rng = np.random.default_rng(0)
IENp = IEN[rng.permutation(IEN.shape[0])]
ctx = sf.context()
for name, ids in (("standard ID", ID), ("permuted ID", (rng.permutation(ID.size) + 1).reshape(ID.shape))):
    mesh = sf.Mesh.from_host(ctx, NL, IENp, ids, 3, 3, ne)
    ctx.timer_start()
    K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
    tp = ctx.timer_stop()
    ctx.timer_start()
    nc, sizes = mesh.colors()
    tc = ctx.timer_stop()
    res = {}
    for mode in ("colored", "atomic"):
        os.environ["SMFEM_VALUES"] = mode
        K.assemble_values(40, 0.4)
        ctx.timer_start()
        for _ in range(3):
            K.assemble_values(40, 0.4)
        res[mode] = ctx.timer_stop() / 3
    os.environ.pop("SMFEM_VALUES")
    K.assemble_values(40, 0.4)   # default: gather form for the standard dof map (first call builds the plan), coloured scatter otherwise
    ctx.timer_start()
    for _ in range(3):
        K.assemble_values(40, 0.4)
    res["default"] = ctx.timer_stop() / 3
    ms = res["default"]
    print(f"  colouring {tc:.1f} ms, {nc} colours (sizes {sizes.min()}..{sizes.max()}); values coloured {res['colored']:.2f} ms, atomic {res['atomic']:.2f} ms, default (gather where possible) {res['default']:.2f} ms")
    t = K.bench_spmv(reps=10, variant=4)
    i = K.info()
    gbs = (12 * i["nnz_local"] + 24 * i["nrows_local"]) / t / 1e6
    print(f"general path, {name}: structured={mesh.info()['structured']} pattern {tp:.1f} ms, values {ms:.2f} ms "
          f"({ne**3 / ms / 1e3:.0f} M el/s), spmv {t:.3f} ms = {gbs:.0f} GB/s")
