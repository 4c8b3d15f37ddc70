"""Matrix-free operator vs the assembled row-triple SpMV: time per application, Jacobi-PCG and multigrid-PCG with either."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smearfem_b200 as sf

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
K.use_matrix_free(True).use_matrix_free(False)   # names the mesh; bench_spmv variant 5 switches the operator itself
for v, name, env in ((4, "CSR row-triple", None), (5, "matrix-free (corner form, v1)", "v1"), (5, "matrix-free (monomial form)", None),
                     (5, "monomial, 128 thr x 4/SM (128 regs)", "c1"), (5, "monomial, 64 thr x 6/SM", "c2"), (5, "monomial, 64 thr x 8/SM", "c3")):
    if env:
        os.environ["SMFEM_MATFREE"] = env
    else:
        os.environ.pop("SMFEM_MATFREE", None)
    ms = K.bench_spmv(reps=30, variant=v)
    print(f"{name:36s}: {ms:.4f} ms per application", flush=True)
os.environ.pop("SMFEM_MATFREE", None)
for mf in (False, True):
    K.use_matrix_free(mf)
    _, it, rr = K.pcg_solve(rtol=1e-10, maxit=20000, want_q=False)
    st = K.pcg_stats()
    print(f"Jacobi-PCG    matrix_free={mf}: {it} iterations, {st['ms_total']:.1f} ms ({st['ms_total'] / it:.4f} ms/iter), relres {rr:.2e}", flush=True)
    K.use_multigrid(True)
    K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
    _, it, rr = K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
    print(f"multigrid-PCG matrix_free={mf}: {it} iterations, {K.pcg_stats()['ms_total']:.1f} ms, relres {rr:.2e}", flush=True)
    K.use_multigrid(False)
