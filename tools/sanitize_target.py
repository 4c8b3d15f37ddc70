"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): structured + general paths, all SpMV variants, PCG."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import smearfem_b200 as sf
from oracle import fem_oracle as o

ctx = sf.context()
for ne in (5, 9):
    r = o.example_problem(ne)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    for tile in ("4x4", "8x4"):
        os.environ["SMFEM_TILE"] = tile
        K.reassemble(40, 0.4)
        nz = K.to_csc()[2]
        assert np.linalg.norm(nz - r["K"].nzval) <= 1e-12 * np.linalg.norm(nz), tile
    K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    for v in (4, 3, 2, 1, 0):
        K.set_spmv_variant(v)
        q, it, rel = K.pcg_solve(rtol=1e-12, maxit=2000)
        assert np.linalg.norm(q - r["q"]) <= 1e-9 * np.linalg.norm(q), v
    K.set_spmv_variant(4)
    K.set_dirichlet_zplanes(0.002)
    q, it, rel = K.pcg_solve(rtol=1e-12, maxit=2000, warm_scale=2.0)
    assert it <= 25
    K.free(); mesh.free()
# general path (permuted ids), 2-D, scalar
NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 4, 3)
rng = np.random.default_rng(1)
IDp = (rng.permutation(ID.size) + 1).reshape(ID.shape)
Kg = sf.assemble_system(4, NL, IEN[rng.permutation(64)], 3, "Q1", 3, IDp, 40, 0.4)
y = Kg.spmv(np.ones(Kg.shape[0]))
NL2, IEN2, ID2, *_ = sf.meshgrid(0, 1, 0, 1, 0, 1, 6, 2)
K2 = sf.assemble_system(6, NL2, IEN2, 2, "Q1", 2, ID2, 40, 0.4)
Ks = sf.assemble_system(6, NL2, IEN2, 2)
print("sanitize target ok")
