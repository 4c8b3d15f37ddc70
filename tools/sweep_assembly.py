"""BASELINE config 5: assembly-only sweep (pattern build + values, timed separately and fused) + SpMV, 1 GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

sizes = [int(a) for a in sys.argv[1:]] or [50, 100, 150, 200, 250, 300]
ctx = sf.context()
out = []
for ne in sizes:
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
    K.assemble_values(40.0, 0.4)
    reps = 5 if ne <= 200 else 3
    def t(f):
        f(); ctx.timer_start()
        for _ in range(reps): f()
        return ctx.timer_stop() / reps
    tp = t(lambda: K.pattern_rebuild())
    tv = t(lambda: K.assemble_values(40.0, 0.4))
    tf = t(lambda: K.reassemble(40.0, 0.4))
    i = K.info()
    assert i["nnz"] == 9 * (3 * (ne + 1) - 2) ** 3
    ts = K.bench_spmv(reps=10, variant=4)
    b_spmv = 12 * i["nnz_local"] + 24 * i["nrows_local"]
    n1 = ne + 1
    b_val = 8 * i["nnz"] + 48 * n1**3 + 64 * ne**3
    b_tot = b_val + 4 * i["nnz"] + 8 * (3 * n1**3 + 1)
    r = dict(ne=ne, elements=ne**3, nnz=i["nnz"], pattern_ms=tp, values_ms=tv, fused_ms=tf,
             el_per_s_values=ne**3 / tv * 1e3, el_per_s_fused=ne**3 / tf * 1e3,
             values_GBs=b_val / tv / 1e6, fused_GBs=b_tot / tf / 1e6, spmv_ms=ts, spmv_GBs=b_spmv / ts / 1e6)
    out.append(r)
    print(json.dumps(r), flush=True)
    K.free(); mesh.free()
