"""Assembly of ONE rank's z-slab of a larger mesh, on one GPU (assembly needs no communication): time_slab.py ne nranks [rank...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf
ne, nranks = int(sys.argv[1]), int(sys.argv[2])
ranks = [int(a) for a in sys.argv[3:]] or [0, nranks // 2, nranks - 1]
for r in ranks:
    ctx = sf.Context(device=0, rank=r, nranks=nranks)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
    def t(f, n=10):
        for _ in range(3): f()
        ctx.timer_start()
        for _ in range(n): f()
        return ctx.timer_stop() / n
    tv = t(lambda: K.assemble_values(40.0, 0.4))
    tf = t(lambda: K.reassemble(40.0, 0.4))
    i = K.info()
    print(f"ne={ne} rank {r}/{nranks}: rows {i['nrows_local']} nnz {i['nnz_local']}  values {tv:.3f} ms  fused {tf:.3f} ms", flush=True)
    del K, mesh
