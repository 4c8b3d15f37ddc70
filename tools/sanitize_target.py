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
        for out in ("0", "2", "3"):   # output routes: direct stores, all-TMA in place, values direct + column indices by TMA
            os.environ["SMFEM_TILE"] = tile
            os.environ["SMFEM_TILE_OUT"] = out
            K.reassemble(40, 0.4)
            nz = K.to_csc()[2]
            assert np.linalg.norm(nz - r["K"].nzval) <= 1e-12 * np.linalg.norm(nz), (tile, out)
    os.environ.pop("SMFEM_TILE_OUT")
    for tile in ("mma75", "mma84", "mma44"):   # fp64 tensor-core (DMMA) kernel, opt-in
        os.environ["SMFEM_TILE"] = tile
        K.reassemble(40, 0.4)
        nzm = K.to_csc()[2]
        assert np.linalg.norm(nzm - r["K"].nzval) <= 1e-12 * np.linalg.norm(nzm), tile
    for tile in ("v2base", "v2l", "v2i", "v2s", "v2g", "v2p", "v2e", "v2all", "v2"):   # layer-march kernel: all instantiations (round 2)
        os.environ["SMFEM_TILE"] = tile
        K.reassemble(40, 0.4)
        nzv = K.to_csc()[2]
        assert np.linalg.norm(nzv - r["K"].nzval) <= 1e-12 * np.linalg.norm(nzv), tile
    os.environ.pop("SMFEM_TILE")
    os.environ["SMFEM_COLIND_SIDE"] = "1"   # persistent closed-form colind kernel on the side stream beside the value kernel
    K.reassemble(40, 0.4)
    assert np.array_equal(K.to_csc()[2], nzv)
    os.environ.pop("SMFEM_COLIND_SIDE")
    nz = K.to_csc()[2]
    # the one-call host route: streamed NodeList (watermark polling), hybrid lattice check on the copy stream
    NLh, IENh, IDh, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NLh, 0, 1, 0, 1)
    for threads in ("0", "2"):
        os.environ["SMFEM_HOST_THREADS"] = threads
        Kh = sf.assemble_system(ne, NLh, IENh, 3, "Q1", 3, IDh, 40, 0.4)
        assert Kh.mesh.info()["structured"]
        assert np.array_equal(Kh.to_csc()[2], nz)
        Kh.free()
    os.environ.pop("SMFEM_HOST_THREADS")
    K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    for v in (4, 3, 2, 1, 0):
        K.set_spmv_variant(v)
        q, it, rel = K.pcg_solve(rtol=1e-12, maxit=2000)
        assert np.linalg.norm(q - r["q"]) <= 1e-9 * np.linalg.norm(q), v
    K.set_spmv_variant(4)
    K.set_dirichlet_zplanes(0.002)
    q, it, rel = K.pcg_solve(rtol=1e-12, maxit=2000, warm_scale=2.0)
    assert it <= 25
    K.use_multigrid(True)   # multigrid-preconditioned CG (odd ne coarsens too: 5 -> 3, 9 -> 5 -> 3)
    qg, itg, relg = K.pcg_solve(rtol=1e-12, maxit=200)
    assert np.linalg.norm(qg - 2 * r["q"]) <= 1e-9 * np.linalg.norm(qg) and itg < 60, itg
    K.use_multigrid(False)
    # matrix-free operator (both element kernels), inside Jacobi-PCG and the multigrid cycle; clone; border extraction
    for ver in ("v1", None):
        if ver:
            os.environ["SMFEM_MATFREE"] = ver
        else:
            os.environ.pop("SMFEM_MATFREE", None)
        K.use_matrix_free(True)
        qm, itm, relm = K.pcg_solve(rtol=1e-12, maxit=2000)
        assert np.linalg.norm(qm - 2 * r["q"]) <= 1e-9 * np.linalg.norm(qm), ver
        K.use_multigrid(True)
        qm, itm, relm = K.pcg_solve(rtol=1e-12, maxit=200)
        assert np.linalg.norm(qm - 2 * r["q"]) <= 1e-9 * np.linalg.norm(qm), ver
        K.use_multigrid(False)
        K.use_matrix_free(False)
    Kc = K.clone()
    assert np.array_equal(Kc.to_csc()[2], K.to_csc()[2])
    Kc.free()
    borders = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)[5]
    CM = np.array([[8 * 2048 / 7.07, 0.0, 2048 / 2], [0.0, 8 * 1536 / 5.3, 1536 / 2], [0.0, 0.0, 1.0]]).T
    for state in ("init", "update"):
        B, S = K.extract_borders(CM, borders, state, ne)
        assert B.shape[1] > 4
    K.free(); mesh.free()
# general path (permuted ids), 2-D, scalar
NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 4, 3)
rng = np.random.default_rng(1)
IDp = (rng.permutation(ID.size) + 1).reshape(ID.shape)
Kg = sf.assemble_system(4, NL, IEN[rng.permutation(64)], 3, "Q1", 3, IDp, 40, 0.4)
y = Kg.spmv(np.ones(Kg.shape[0]))
assert Kg.mesh.colors()[0] >= 8   # coloured scatter (default on general meshes); the atomic one:
os.environ["SMFEM_VALUES"] = "atomic"; Kg.assemble_values(40, 0.4); os.environ.pop("SMFEM_VALUES")
NL2, IEN2, ID2, *_ = sf.meshgrid(0, 1, 0, 1, 0, 1, 6, 2)
K2 = sf.assemble_system(6, NL2, IEN2, 2, "Q1", 2, ID2, 40, 0.4)
Ks = sf.assemble_system(6, NL2, IEN2, 2)
# Q2 9-node scalar quads (general path)
n = 2 * 3 + 1
xs = np.arange(n) / (n - 1)
NLq = np.array([[x, y] for y in xs for x in xs]).T.copy()
IENq = np.array([[2 * j * n + 2 * i + a + 1 for a in (0, 2, 2 * n + 2, 2 * n, 1, n + 2, 2 * n + 1, n, n + 1)] for j in range(3) for i in range(3)])
Kq = sf.assemble_system(3, NLq, IENq, 2, "Q2", 1)
print("sanitize target ok")
