"""Parity of the CUDA path (through the C ABI) with the oracle.  Run on the GPU box: -m gpu.

Bars (BASELINE.json north_star): sparsity pattern bit-exact; ||K-K_ref||_F / ||K_ref||_F <= 1e-10;
||u-u_ref||_2 / ||u_ref||_2 <= 1e-10 (fp64).  Mesh generation / inflate are bit-exact."""
import numpy as np
import pytest

import smearfem_b200 as sf
from oracle import fem_oracle as o

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def _load_csc(g, p):
    return o.JuliaCSC(int(g[p + "_m"]), int(g[p + "_n"]), g[p + "_colptr"], g[p + "_rowval"], g[p + "_nzval"])


def assert_csc_parity(K, Ko, tol=TOL):
    colptr, rowval, nzval = K.to_csc()
    assert K.shape == (Ko.m, Ko.n) and K.nnz == Ko.nnz
    assert np.array_equal(colptr, Ko.colptr), "colptr differs (pattern must be bit-exact)"
    assert np.array_equal(rowval, Ko.rowval), "rowval differs (pattern must be bit-exact)"
    assert rel(nzval, Ko.nzval) <= tol
    # explicit zeros are structural, as in Julia's sparse()
    assert (nzval == 0.0).sum() == (Ko.nzval == 0.0).sum() or rel(nzval, Ko.nzval) <= tol


# ---- mesh --------------------------------------------------------------------------------------
@pytest.mark.parametrize("ne", [1, 2, 7, 20])
def test_meshgrid_and_inflate_bit_exact(ne):
    NL, IEN, ID, top, btm, borders = sf.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    NLo, IENo, IDo, topo, btmo, borderso = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    assert np.array_equal(NL, NLo) and np.array_equal(IEN, IENo) and np.array_equal(ID, IDo)
    assert np.array_equal(top, topo) and np.array_equal(btm, btmo)
    assert all(list(a) == list(b) for a, b in zip(borders, borderso))
    out = sf.inflate_sphere(NL, 0, 1, 0, 1)
    assert out is NL  # in place, like the reference
    assert np.array_equal(NL, o.inflate_sphere(NLo, 0, 1, 0, 1))


def test_meshgrid_general_box_bit_exact():
    NL, *_ = sf.meshgrid(-0.3, 1.7, 0.25, 0.5, 2.0, 5.0, 6, 3)
    NLo, *_ = o.meshgrid(-0.3, 1.7, 0.25, 0.5, 2.0, 5.0, 6, 3)
    assert np.array_equal(NL, NLo)


# ---- assembly: structured fast path (device mesh) ---------------------------------------------------
@pytest.mark.parametrize("ne,inflate", [(1, False), (2, False), (2, True), (5, True), (8, True), (20, True)])
def test_hex_elasticity_structured(ne, inflate):
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3)
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    if inflate:
        mesh.inflate_sphere(0, 1, 0, 1)
        o.inflate_sphere(NL, 0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    Ko = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    assert K.nnz == 9 * (3 * (ne + 1) - 2) ** 3
    assert_csc_parity(K, Ko)
    assert rel(K.diag(), Ko.to_scipy().diagonal()) <= TOL
    K.free()
    mesh.free()


@pytest.mark.parametrize("name", ["hex_ne2_cube", "hex_ne2_inflated", "hex_ne4_cube", "hex_ne4_inflated"])
def test_golden_fixtures_through_reference_api(golden_dir, name):
    """Reads like the reference's own (missing) integration test: host arrays in, CSC out."""
    g = np.load(f"{golden_dir}/{name}.npz")
    ne = int(g["ne"])
    K = sf.assemble_system(ne, g["NodeList"], g["IEN"], 3, "Q1", 3, g["ID"], 40, 0.4)
    assert K.mesh.info()["structured"]
    assert_csc_parity(K, _load_csc(g, "K"))
    b = sf.apply_boundary_conditions(ne, g["NodeList"], g["IEN"], g["IEN_top"], g["IEN_btm"], 3, "Q1", g["ID"])
    bo = _load_csc(g, "b")
    colptr, rowval, nzval = b.to_csc()
    assert np.array_equal(colptr, bo.colptr) and np.array_equal(rowval, bo.rowval)
    assert rel(nzval, bo.nzval) <= TOL
    K_bar = K + 100 * b  # examples/vector3D.jl:308
    assert K_bar is not K
    assert_csc_parity(K, _load_csc(g, "K"))  # K itself is untouched, as in the reference (K_bar is a device-side copy)
    q_d, C = sf.setboundaryCond(g["NodeList"], ne, 3, "Q1", 0.001, 3)
    q = sf.solve(K_bar, q_d, C, rtol=1e-13)
    assert rel(q, g["q"]) <= TOL


# ---- assembly: general (unstructured) path ----------------------------------------------------------
def test_general_path_permuted_ids_and_elements():
    """A lattice mesh with shuffled element order and a permuted ID map must take the general path
    and still reproduce Julia's pattern for THAT numbering."""
    ne = 4
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    rng = np.random.default_rng(7)
    IENp = IEN[rng.permutation(IEN.shape[0])]
    IDp = (rng.permutation(ID.size) + 1).reshape(ID.shape).astype(np.int64)
    K = sf.assemble_system(ne, NL, IENp, 3, "Q1", 3, IDp, 40, 0.4)
    assert not K.mesh.info()["structured"]
    Ko = o.assemble_system(ne, NL, IENp, 3, "Q1", 3, IDp, 40, 0.4)
    assert_csc_parity(K, Ko)


def test_general_path_shuffled_elements_standard_ids():
    """Shuffled element order with the standard dof map: general (unstructured) path, but the device recognises the
    node-major ID (one CSR search per node pair, row-triple SpMV); the fold order over elements differs from the
    reference's ascending order only in rounding."""
    ne = 6
    NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    perm = np.random.default_rng(11).permutation(IEN.shape[0])
    K = sf.assemble_system(ne, NL, IEN[perm], 3, "Q1", 3, ID, 40, 0.4)
    assert not K.mesh.info()["structured"]
    Ko = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)  # same matrix: element order does not change K
    assert_csc_parity(K, Ko, tol=1e-13)
    # the whole example pipeline on the general path (explicit face lists, general Dirichlet list)
    b = sf.apply_boundary_conditions(ne, NL, IEN[perm], top, btm, 3, "Q1", ID)
    K_bar = K + 100 * b
    q_d, C = sf.setboundaryCond(NL, ne, 3, "Q1", 0.001, 3)
    q = sf.solve(K_bar, q_d, C, rtol=1e-13)
    assert rel(q, o.example_problem(ne)["q"]) <= TOL


def test_general_path_coloured_scatter_is_deterministic(monkeypatch):
    """General meshes are assembled colour by colour, one add per entry and launch: the colouring is valid (no two
    elements of a colour share a node), covers every element, and two assemblies give the same bits; the atomic scatter
    (SMFEM_VALUES=atomic) agrees to rounding.  Also 2-D Q4 and scalar problems."""
    ne = 7
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=3)
    rng = np.random.default_rng(21)
    IENp = IEN[rng.permutation(IEN.shape[0])]
    IDp = (rng.permutation(ID.size) + 1).reshape(ID.shape).astype(np.int64)
    for ids in (ID, IDp):
        mesh = sf.Mesh.from_host(ctx, NL, IENp, ids, 3, 3, ne)
        assert not mesh.info()["structured"]
        nc, sizes = mesh.colors()
        assert 8 <= nc <= 64 and sizes.sum() == ne**3 and sizes.min() > 0
        K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
        Ko = o.assemble_system(ne, NL, IENp, 3, "Q1", 3, ids, 40, 0.4)
        assert_csc_parity(K, Ko, tol=1e-13)
        nz1 = K.to_csc()[2]
        for _ in range(3):
            K.assemble_values(40, 0.4)
            assert np.array_equal(K.to_csc()[2], nz1), "coloured scatter must be bit-reproducible"
        monkeypatch.setenv("SMFEM_VALUES", "atomic")
        K.assemble_values(40, 0.4)
        monkeypatch.delenv("SMFEM_VALUES")
        assert rel(K.to_csc()[2], nz1) <= 1e-14
    # scalar hex and 2-D plane stress on shuffled connectivities
    Ks = sf.assemble_system(ne, NL, IENp, 3, "Q1", 1)
    assert_csc_parity(Ks, o.assemble_system(ne, NL, IENp, 3, "Q1", 1), tol=1e-13)
    NL2, IEN2, ID2, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 9, 2)
    IEN2p = IEN2[rng.permutation(IEN2.shape[0])]
    K2 = sf.assemble_system(9, NL2, IEN2p, 2, "Q1", 2, ID2, 40, 0.4)
    assert not K2.mesh.info()["structured"] and K2.mesh.colors()[0] >= 4
    assert_csc_parity(K2, o.assemble_system(9, NL2, IEN2p, 2, "Q1", 2, ID2, 40, 0.4), tol=1e-13)
    # lattice meshes need no colouring
    with pytest.raises(sf.SmearFEMError):
        sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, 3, 3).colors()


def test_assemble_system_one_call_matches_two_step():
    """smfem_assemble_system (transfers of IEN / ID overlapped with a speculative lattice assembly) gives the bits of
    smfem_mesh_from_host + smfem_assemble; a mesh with meshgrid's sizes but another numbering falls back correctly."""
    ne = 11
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=5)
    for trial in range(3):  # repeated: the copy stream reuses cached buffers of the previous call
        K1 = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
        assert K1.mesh.info()["structured"]
        mesh = sf.Mesh.from_host(ctx, NL, IEN, ID, 3, 3, ne)
        K2 = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
        for a, b in zip(K1.to_csc(), K2.to_csc()):
            assert np.array_equal(a, b)
        assert np.array_equal(K1.diag(), K2.diag())
    assert_csc_parity(K1, o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4))
    # same sizes, two elements swapped: the lattice check fails after the speculative assembly was launched
    IEN2 = IEN.copy()
    IEN2[[3, 77]] = IEN2[[77, 3]]
    K3 = sf.assemble_system(ne, NL, IEN2, 3, "Q1", 3, ID, 40, 0.4)
    assert not K3.mesh.info()["structured"]
    assert_csc_parity(K3, o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4), tol=1e-13)
    # and a lattice call right after the discarded one
    K4 = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    for a, b in zip(K4.to_csc(), K1.to_csc()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("ne", [2, 4, 9])
def test_plane_stress_quad4(ne, golden_dir):  # config C1 geometry, src/fem.jl:210-217
    NL, IEN, ID, *_ = sf.meshgrid(0, 1, 0, 1, 0, 1, ne, 2)
    K = sf.assemble_system(ne, NL, IEN, 2, "Q1", 2, ID, 40, 0.4)
    Ko = o.assemble_system_literal(ne, NL, IEN, 2, "Q1", 2, ID, 40, 0.4)
    assert K.nnz == 4 * (3 * (ne + 1) - 2) ** 2
    assert_csc_parity(K, Ko)
    if ne in (2, 4):
        assert_csc_parity(K, _load_csc(np.load(f"{golden_dir}/quad_ne{ne}.npz"), "K"))


@pytest.mark.parametrize("ne", [2, 8, 16])
def test_plane_stress_solve_c1(ne, golden_dir):
    """BASELINE config C1 end to end on the device: 2-D Q4 plane stress, bottom edge clamped, u_y = -0.001 on the top edge
    (the reference has no executable 2-D boundary code), solved by the masked Jacobi-PCG; <= 1e-10 vs the oracle's
    direct solve and vs the committed fixture."""
    r = o.plane_stress_problem(ne)
    K = sf.assemble_system(ne, r["NodeList"], r["IEN"], 2, "Q1", 2, r["ID"], 40, 0.4)
    assert_csc_parity(K, r["K"])
    K.set_dirichlet(r["fixed"], r["q_d"][r["fixed"] - 1, 0])
    q, it, relres = K.pcg_solve(rtol=1e-13, maxit=5000)
    assert relres <= 1e-12 and rel(q, r["q"]) <= TOL
    if ne in (8, 16):
        g = np.load(f"{golden_dir}/c1_plane_stress_ne{ne}.npz")
        assert rel(q, g["q"]) <= TOL
    # the reference-facing idiom (setboundaryCond-style q_d + constraint -> solve)
    q2 = sf.solve(K, r["q_d"], sf.Constraint(K.shape[0], r["free"]), rtol=1e-13)
    assert rel(q2, r["q"]) <= TOL
    K.free()


@pytest.mark.parametrize("ne", [50, 64])
def test_hex_elasticity_oracle_parity_multiwave(ne):
    """Oracle parity where the tile kernels run multi-wave grids and multi-chunk plans (the largest oracle
    comparison was ne = 20, one wave): pattern bit-exact and values <= 1e-10 against the C form of the oracle at 50^3 and
    64^3 (inflated + jittered nodes), for the layer-march kernel and for the first tile kernel, device-mesh and host-array routes."""
    import os

    from oracle import c_oracle

    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    o.jitter_nodes(NL, ne, seed=7, amp=0.15)
    Ko = c_oracle.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=c_oracle.max_threads())
    assert Ko.nnz == 9 * (3 * (ne + 1) - 2) ** 3
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    colptr, rowval, nzval = K.to_csc()
    assert np.array_equal(colptr, Ko.colptr) and np.array_equal(rowval, Ko.rowval)
    assert rel(nzval, Ko.nzval) <= TOL
    err_v2 = rel(nzval, Ko.nzval)
    os.environ["SMFEM_TILE"] = "4x4"
    try:
        K.reassemble(40, 0.4)
        assert rel(K.to_csc()[2], Ko.nzval) <= TOL
    finally:
        os.environ.pop("SMFEM_TILE")
    K.free()
    mesh.free()
    del colptr, rowval
    Kh = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)  # host arrays, streamed coordinates
    assert Kh.mesh.info()["structured"]
    assert np.array_equal(Kh.to_csc()[2], nzval), "host-array route must give the same bits as the device-mesh route"
    print(f"ne={ne}: rel ||K - K_oracle|| = {err_v2:.2e}")
    Kh.free()


def test_surface_term_uses_its_own_nodelist():
    """examples/vector3D.jl:306-308: b is integrated over the NodeList passed to apply_boundary_conditions, which need not be
    the one K was assembled from: K from the cube, b from the inflated mesh."""
    ne = 5
    NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    NLi = o.inflate_sphere(NL.copy(order="F"), 0, 1, 0, 1)
    Ko = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    bo = o.apply_boundary_conditions(ne, NLi, IEN, top, btm, 3, "Q1", ID)
    Kbo = o.add_scaled(Ko, bo, 100)
    K = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    b = sf.apply_boundary_conditions(ne, NLi, IEN, top, btm, 3, "Q1", ID)
    K_bar = K + 100 * b
    A = K_bar.to_scipy().toarray()
    assert rel(A, Kbo.to_scipy().toarray()) <= TOL
    # and it differs from the term integrated over K's own (cube) coordinates by far more than the tolerance
    b_cube = o.apply_boundary_conditions(ne, NL, IEN, top, btm, 3, "Q1", ID)
    assert rel(o.add_scaled(Ko, b_cube, 100).to_scipy().toarray(), Kbo.to_scipy().toarray()) > 1e-3


@pytest.mark.parametrize("ndim,ne", [(2, 5), (3, 2), (3, 6)])
def test_scalar_laplace(ndim, ne):  # nDof = 1: raw node ids, no ID (src/fem.jl:199-208)
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, ndim)
    if ndim == 3:
        o.inflate_sphere(NL, 0, 1, 0, 1)
    K = sf.assemble_system(ne, NL, IEN, ndim)  # defaults: "Q1", nDof=1, ID=nothing, as upstream
    Ko = o.assemble_system(ne, NL, IEN, ndim, "Q1", 1)
    assert_csc_parity(K, Ko)
    if ndim == 3:
        assert K.nnz == (3 * (ne + 1) - 2) ** 3
    # constants are in the null space of the Laplacian
    y = K.spmv(np.ones(K.shape[0]))
    assert np.abs(y).max() <= 1e-12 * np.abs(Ko.nzval).max()


def _q2_mesh(ne):
    """Structured 9-node quad mesh on [0,1]^2 in the reference's local ordering (src/fem.jl:80-88: corners CCW,
    then bottom / right / top / left mid-sides, then centre).  The reference ships no Q2 mesh generator."""
    n = 2 * ne + 1
    xs = np.arange(n) / (n - 1)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    NL = np.asfortranarray(np.vstack([xs[ii.ravel()], xs[jj.ravel()]]))
    nid = lambda i, j: j * n + i + 1
    IEN = np.zeros((ne * ne, 9), dtype=np.int64)
    e = 0
    for ej in range(ne):
        for ei in range(ne):
            i0, j0 = 2 * ei, 2 * ej
            IEN[e] = [nid(i0, j0), nid(i0 + 2, j0), nid(i0 + 2, j0 + 2), nid(i0, j0 + 2), nid(i0 + 1, j0), nid(i0 + 2, j0 + 1),
                      nid(i0 + 1, j0 + 2), nid(i0, j0 + 1), nid(i0 + 1, j0 + 1)]
            e += 1
    return NL, IEN


@pytest.mark.parametrize("ne", [1, 3, 6])
def test_q2_scalar_quads(ne):
    """FunctionClass "Q2": 2-D scalar, 9-node quads, the reference's 2x2 (under-)integration (src/fem.jl:77-111, :199-208)."""
    NL, IEN = _q2_mesh(ne)
    rng = np.random.default_rng(ne)
    NL[:, :] += rng.uniform(-0.02, 0.02, NL.shape) / ne  # non-affine elements
    K = sf.assemble_system(ne, NL, IEN, 2, "Q2", 1)
    Ko = o.assemble_system_literal(ne, NL, IEN, 2, "Q2", 1)
    assert_csc_parity(K, Ko)
    assert np.abs(K.spmv(np.ones(K.shape[0]))).max() <= 1e-11 * np.abs(Ko.nzval).max()  # constants in the null space
    with pytest.raises(sf.SmearFEMError):  # upstream Q2 is scalar-only
        sf.assemble_system(ne, NL, IEN, 2, "Q2", 2, np.arange(1, 2 * NL.shape[1] + 1).reshape(2, -1).T.copy())
    with pytest.raises(sf.SmearFEMError):  # Q1 with 9-node connectivity
        sf.assemble_system(ne, NL, IEN, 2, "Q1", 1)


def test_jittered_mesh_no_uniform_shortcut():
    ne = 10
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=1234)
    K = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    assert K.mesh.info()["structured"]
    assert_csc_parity(K, o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4))


# ---- surface term, Dirichlet, solve -------------------------------------------------------------------
@pytest.mark.parametrize("ne", [3, 8])
def test_example_pipeline_device_resident(ne):
    """examples/vector3D.jl:283-322 with everything on the device."""
    ctx = sf.context()
    r = o.example_problem(ne)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    K.add_surface_mass(100.0, keep_b=True)
    cp, rv, bz = K.to_csc(which=1)
    assert abs(bz.sum() - r["b"].nzval.sum()) <= 1e-12 * abs(bz.sum())
    _, _, kb = K.to_csc()
    assert rel(kb, r["K_bar"].nzval) <= TOL
    K.set_dirichlet_zplanes(0.001)
    q, it, relres = K.pcg_solve(rtol=1e-13, maxit=10000)
    assert relres <= 1e-13 and it < 10000
    assert rel(q, r["q"]) <= TOL
    # q is linear in d: second load step of examples/vector3D.jl:310
    K.set_dirichlet_zplanes(0.011)
    q2, *_ = K.pcg_solve(rtol=1e-13, maxit=10000)
    assert rel(q2, 11 * r["q"]) <= TOL


def test_config_c2_ne20_against_golden(golden_dir):
    """BASELINE config 2: examples/vector3D.jl with ne = 20, u vs the oracle's sparse direct solve."""
    g = np.load(f"{golden_dir}/example_summaries.npz")
    ne = 20
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    _, _, nz = K.to_csc()
    s = g["ne20"]
    assert K.nnz == int(s[0]) == 2042829
    assert abs(np.linalg.norm(nz) - s[1]) <= 1e-11 * s[1] and abs(K.diag().sum() - s[2]) <= 1e-11 * s[2]
    K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    q, it, relres = K.pcg_solve(rtol=1e-13, maxit=20000)
    assert rel(q, g["q_ne20"]) <= TOL
    assert abs(np.abs(q[0::3]).max() - s[7]) <= 1e-9 * s[7]


def test_tile_and_atomic_value_kernels_agree(monkeypatch):
    """The structured path has two value kernels (tiled gather = default, atomic scatter = general);
    both must meet the parity bar, and the tiled one must be bit-reproducible run to run."""
    ne = 13  # not a multiple of the 8x4 tile: exercises partial tiles
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=99)
    Ko = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    assert_csc_parity(K, Ko)
    nz1 = K.to_csc()[2]
    d1 = K.diag()
    K.assemble_values(40, 0.4)
    assert np.array_equal(K.to_csc()[2], nz1), "tiled gather kernel must be deterministic"
    # fused reassembly (rowptr closed form + colind written by the value kernel) rewrites EVERY entry of the pattern
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    K.reassemble(40, 0.4)
    monkeypatch.delenv("SMFEM_DEBUG_CLEAR")
    assert_csc_parity(K, Ko)
    assert np.array_equal(K.to_csc()[2], nz1)
    # every tile shape x output route (0 direct block stores, 1 CSR-ordered run + coalesced stores, 2 TMA bulk store) x chunk
    # plan writes every entry of val / colind / diag and gives the same bits
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    monkeypatch.setenv("SMFEM_TILE", "4x4")  # reference bits: the first tile kernel (the default is the layer-march kernel, tested below)
    K.reassemble(40, 0.4)
    nz1, d1 = K.to_csc()[2], K.diag()
    for tile in ("8x4", "4x4"):
        for out in ("0", "2", "3"):
            for chunks in (None, "5,4,5", "14"):
                monkeypatch.setenv("SMFEM_TILE", tile)
                monkeypatch.setenv("SMFEM_TILE_OUT", out)
                if chunks:
                    monkeypatch.setenv("SMFEM_TILE_CHUNKS", chunks)
                K.reassemble(40, 0.4)
                assert_csc_parity(K, Ko)
                assert np.array_equal(K.to_csc()[2], nz1), (tile, out, chunks)
                assert np.array_equal(K.diag(), d1), (tile, out, chunks)
                monkeypatch.delenv("SMFEM_TILE_CHUNKS", raising=False)
    # the fp64 tensor-core (DMMA) kernel, opt-in: same pattern, same values to rounding (the fold order differs),
    # every entry rewritten, bit-reproducible for every tile shape and chunk plan
    monkeypatch.delenv("SMFEM_TILE_OUT")
    for tile in ("v2", "mma75", "mma84", "mma44"):  # v2: the layer-march kernel (assemble_tile2.cu)
        ref = None
        for chunks in (None, "5,4,5", "1,13"):
            monkeypatch.setenv("SMFEM_TILE", tile)
            if chunks:
                monkeypatch.setenv("SMFEM_TILE_CHUNKS", chunks)
            K.reassemble(40, 0.4)
            assert_csc_parity(K, Ko)
            nz = K.to_csc()[2]
            assert rel(nz, nz1) <= 1e-14 and rel(K.diag(), d1) <= 1e-14, (tile, chunks)
            if ref is None:
                ref = nz
            assert np.array_equal(nz, ref), (tile, chunks)
            monkeypatch.delenv("SMFEM_TILE_CHUNKS", raising=False)
    for v in ("SMFEM_TILE", "SMFEM_DEBUG_CLEAR"):
        monkeypatch.delenv(v)
    monkeypatch.setenv("SMFEM_VALUES", "atomic")
    K.assemble_values(40, 0.4)
    assert_csc_parity(K, Ko)
    assert rel(K.diag(), d1) <= 1e-14
    monkeypatch.delenv("SMFEM_VALUES")


@pytest.mark.parametrize("ne,chunks", [(1, None), (2, None), (3, "4"), (9, "3,3,4"), (21, None), (37, "20,11,7"), (37, "1,1,36")])
def test_layer_march_kernel_matches_oracle_and_tile_kernel(monkeypatch, ne, chunks):
    """k_values_tile2 (SMFEM_TILE=v2): pattern bit-exact, values <= 1e-10 vs the oracle and <= 1e-13 vs the first tile kernel,
    every entry of val / colind / diag rewritten, bit-reproducible, independent of the chunk plan."""
    from oracle import c_oracle

    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    if ne > 1:
        o.jitter_nodes(NL, ne, seed=5)
    Ko = c_oracle.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=c_oracle.max_threads())
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    monkeypatch.setenv("SMFEM_TILE", "4x4")
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    nz1, d1 = K.to_csc()[2], K.diag()
    monkeypatch.setenv("SMFEM_TILE", "v2")
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    if chunks:
        monkeypatch.setenv("SMFEM_TILE_CHUNKS", chunks)
    K.reassemble(40, 0.4)
    assert_csc_parity(K, Ko)
    nz2, d2 = K.to_csc()[2], K.diag()
    assert rel(nz2, nz1) <= 1e-13 and rel(d2, d1) <= 1e-13
    monkeypatch.delenv("SMFEM_TILE_CHUNKS", raising=False)
    K.reassemble(40, 0.4)  # automatic chunk plan: same bits
    assert np.array_equal(K.to_csc()[2], nz2) and np.array_equal(K.diag(), d2)
    monkeypatch.delenv("SMFEM_DEBUG_CLEAR")
    K.assemble_values(40, 0.4)  # values only (colind untouched)
    assert_csc_parity(K, Ko)
    assert np.array_equal(K.to_csc()[2], nz2)
    K.free()
    mesh.free()


@pytest.mark.parametrize("ne", [5, 22, 37])
def test_layer_march_kernel_variants_same_bits(monkeypatch, ne):
    """The instantiations of k_values_tile2 (shared-memory layout, compile-time section strides, 64-thread CTAs) only move data
    differently: values, column indices and diagonal are BIT-identical to the first layer-march version."""
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    monkeypatch.setenv("SMFEM_TILE", "v2base")
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    ref = K.to_csc() + (K.diag(),)
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    for variant in ("v2", "v2l", "v2i", "v2li", "v2s", "v2g", "v2p", "v2e", "v2all"):
        monkeypatch.setenv("SMFEM_TILE", variant)
        K.reassemble(40, 0.4)
        got = K.to_csc() + (K.diag(),)
        for a, b in zip(ref, got):
            assert np.array_equal(a, b), variant
    K.free()
    mesh.free()


@pytest.mark.parametrize("ne", [1, 2, 7, 30])
def test_side_stream_colind_kernel_same_pattern(monkeypatch, ne):
    """SMFEM_COLIND_SIDE=1: the column indices are written by the persistent closed-form kernel on the side stream while the value
    kernel runs without its colind output: same colptr / rowval / nzval bits as the fused launch (buffers cleared to 0xFF first)."""
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    ref = K.to_csc(which=2) + (K.diag(),)
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    monkeypatch.setenv("SMFEM_COLIND_SIDE", "1")
    K.reassemble(40, 0.4)
    got = K.to_csc(which=2) + (K.diag(),)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)
    K.free()
    mesh.free()


@pytest.mark.parametrize("ne,beta", [(1, 100.0), (2, 0.0), (5, 100.0), (12, 100.0), (33, 7.5)])
def test_matrix_free_operator_matches_assembled_spmv(ne, beta):
    """SURVEY 8(f) row 3 (second half): y = (K + beta b) x applied matrix-free from the coordinates equals the CSR SpMV of the
    assembled K_bar (<= 1e-13 relative, different summation order only), on an inflated + jittered lattice, and is bit-reproducible."""
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    if ne > 1:
        o.jitter_nodes(NL, ne, seed=3)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    if beta:
        K.add_surface_mass(beta)
    x = np.random.default_rng(ne).standard_normal(3 * (ne + 1) ** 3)
    y_csr = K.spmv(x)
    K.use_matrix_free(True)
    y_mf = K.spmv(x)
    assert rel(y_mf, y_csr) <= 1e-13
    assert np.array_equal(K.spmv(x), y_mf)
    K.use_matrix_free(False)
    assert np.array_equal(K.spmv(x), y_csr)
    K.free()
    mesh.free()


@pytest.mark.parametrize("ne", [6, 20])
def test_matrix_free_pcg_matches_oracle(ne):
    """The example problem solved with the matrix-free operator inside Jacobi-PCG and inside the multigrid-preconditioned CG:
    ||u - u_ref|| / ||u_ref|| <= 1e-10 against the oracle's direct solve, same iteration counts (+-2) as with the assembled K."""
    ctx = sf.context()
    r = o.example_problem(ne)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    q0, it0, _ = K.pcg_solve(rtol=1e-13, maxit=20000)
    K.use_matrix_free(True)
    q1, it1, rel1 = K.pcg_solve(rtol=1e-13, maxit=20000)
    assert rel(q1, r["q"]) <= TOL and rel1 <= 1e-12 and abs(it1 - it0) <= 2
    K.use_multigrid(True)
    q2, it2, rel2 = K.pcg_solve(rtol=1e-13, maxit=200)
    assert rel(q2, r["q"]) <= TOL and rel2 <= 1e-12 and it2 <= 40
    K.use_multigrid(False)
    K.use_matrix_free(False)
    q3, it3, _ = K.pcg_solve(rtol=1e-13, maxit=20000)   # back on the assembled operator: the cached iteration graph is per operator
    assert it3 == it0 and np.array_equal(q3, q0)
    K.free()
    mesh.free()


@pytest.mark.parametrize("ne", [3, 12, 33])
def test_matrix_free_operator_without_assembled_matrix(ne):
    """smfem_matfree_operator: the operator handle that never holds K (no rowptr / colind / val).  Its diagonal, its products and the
    solves through it (Jacobi-PCG and multigrid-PCG, incl. the surface term and the Dirichlet data) equal those of the assembled
    K_bar / the oracle; calls that need CSR arrays fail loudly."""
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    o.jitter_nodes(NL, ne, seed=4)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
    F = sf.SparseMatrixB200.matrix_free(ctx, mesh, 40, 0.4)
    assert F.info()["nnz_local"] == 0 and F.info()["nnz"] == K.info()["nnz"] and F.shape == K.shape
    F.add_surface_mass(100.0)
    assert rel(F.diag(), K.diag()) <= 1e-13
    x = np.random.default_rng(ne).standard_normal(3 * (ne + 1) ** 3)
    assert rel(F.spmv(x), K.spmv(x)) <= 1e-13
    for M in (K, F):
        M.set_dirichlet_zplanes(0.001)
    q0, it0, _ = K.pcg_solve(rtol=1e-13, maxit=20000)
    q1, it1, rel1 = F.pcg_solve(rtol=1e-13, maxit=20000)
    assert rel(q1, q0) <= 1e-10 and rel1 <= 1e-12 and abs(it1 - it0) <= max(2, it0 // 100)   # rtol 1e-13 sits at the rounding floor
    F.use_multigrid(True)
    q2, it2, rel2 = F.pcg_solve(rtol=1e-13, maxit=200)
    assert rel(q2, q0) <= 1e-10 and rel2 <= 1e-12 and it2 <= 60   # jittered lattice, rtol 1e-13
    F.use_multigrid(False)
    C2 = F.clone()
    assert np.array_equal(C2.diag(), F.diag())
    C2.free()
    for bad in (lambda: F.to_csc(), lambda: F.reassemble(40, 0.4), lambda: F.use_matrix_free(False)):
        with pytest.raises(sf.SmearFEMError):
            bad()
    F.free()
    K.free()
    mesh.free()


def test_spmv_variants_and_host_spmv():
    ne = 9
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    K = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    A = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4).to_scipy()
    x = np.random.default_rng(3).standard_normal(K.shape[0])
    for v in (0, 1, 2, 3, 4):
        K.set_spmv_variant(v)
        assert rel(K.spmv(x), A @ x) <= 1e-13


@pytest.mark.parametrize("ne", [8, 13, 20])  # 13 -> 7 -> 4: odd sizes coarsen too
def test_multigrid_pcg_matches_oracle(ne):
    """SURVEY 8(f) row 3, opt-in: CG preconditioned by a geometric multigrid V-cycle (re-assembled coarse levels, Chebyshev
    smoothing) reaches the oracle's direct solution within the same 1e-10 and needs far fewer iterations than Jacobi-PCG."""
    ctx = sf.context()
    r = o.example_problem(ne)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    qj, itj, _ = K.pcg_solve(rtol=1e-13, maxit=5000)
    K.use_multigrid(True)
    qg, itg, relg = K.pcg_solve(rtol=1e-13, maxit=500)
    assert rel(qj, r["q"]) <= TOL and rel(qg, r["q"]) <= TOL
    assert relg <= 1e-13 and itg <= 40 and itg < itj / 3, (itg, itj)
    # another load step on the same hierarchy; a re-assembly with another material rebuilds it
    K.set_dirichlet_zplanes(0.011)
    q2, it2, _ = K.pcg_solve(rtol=1e-13, maxit=500)
    assert rel(q2, 11 * r["q"]) <= 1e-9 and it2 <= 40
    # warm-started load step (q is linear in d, examples/vector3D.jl:310-338): converged after a couple of iterations
    K.set_dirichlet_zplanes(0.021)
    q4, it4, _ = K.pcg_solve(rtol=1e-12, maxit=500, warm_scale=21.0 / 11.0)
    assert rel(q4, 21 * r["q"]) <= 1e-9 and it4 <= 3, it4
    K.set_dirichlet_zplanes(0.011)
    K.use_multigrid(False)
    q3, it3, _ = K.pcg_solve(rtol=1e-13, maxit=5000)
    assert rel(q3, q2) <= 1e-10 and it3 > it2
    # Jacobi-PCG warm-started from a multigrid solution (the solvers keep their solutions in different buffers)
    K.use_multigrid(True)
    K.pcg_solve(rtol=1e-13, maxit=500)
    K.use_multigrid(False)
    K.set_dirichlet_zplanes(0.022)
    q5, it5, _ = K.pcg_solve(rtol=1e-12, maxit=5000, warm_scale=2.0)
    assert rel(q5, 22 * r["q"]) <= 1e-9 and it5 <= 25, it5


def test_multigrid_pcg_general_dirichlet_jittered_mesh():
    """Multigrid on a jittered lattice with a general Dirichlet list (bottom face clamped in all components, top z prescribed)
    and no surface term: the coarse masks come from injection of the fine one; same solution as Jacobi-PCG, < 1/3 of its iterations."""
    ne = 12
    ctx = sf.context()
    NL, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=7, amp=0.3)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).set_nodelist(NL)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    n1 = ne + 1
    kz = np.arange(n1**3) // (n1 * n1)
    btm, top = np.where(kz == 0)[0], np.where(kz == ne)[0]
    dofs = np.concatenate([3 * btm + 1, 3 * btm + 2, 3 * btm + 3, 3 * top + 3])  # 1-based
    vals = np.concatenate([np.zeros(3 * btm.size), np.full(top.size, -0.001)])
    order = np.argsort(dofs)
    K.set_dirichlet(dofs[order], vals[order])
    qj, itj, _ = K.pcg_solve(rtol=1e-12, maxit=5000)
    K.use_multigrid(True)
    qg, itg, relg = K.pcg_solve(rtol=1e-12, maxit=300)
    assert relg <= 1e-12 and rel(qg, qj) <= 1e-9 and itg <= 40 and itg < itj / 3, (itg, itj)


@pytest.mark.parametrize("ne", [2, 8, 21])
def test_extract_borders_on_device(ne):
    """extract_borders (src/PostProcess.jl:60-117), states "init" and "update", on the deformed mesh of the last
    solve, against the oracle's restatement: the same border nodes in the same order (values to rounding)."""
    ctx = sf.context()
    NL, IEN, ID, top, btm, borders = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    CM = np.array([[8 * 2048 / 7.07, 0.0, 2048 / 2], [0.0, 8 * 1536 / 5.3, 1536 / 2], [0.0, 0.0, 1.0]]).T  # examples/vector3D.jl:281
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
    # before any solve (the reference's "init" call, examples/vector3D.jl:288)
    B, S = K.extract_borders(CM, borders, "init", ne)
    Bo, So = o.extract_borders(NL, CM, borders, "init", ne)
    assert B.shape == Bo.shape and rel(B, Bo) <= 1e-14 and rel(S, So) <= 1e-14
    # after a load step (examples/vector3D.jl:325-329)
    K.set_dirichlet_zplanes(0.05)
    q, _, _ = K.pcg_solve(rtol=1e-13, maxit=5000)
    new = NL + np.vstack([q[ID[:, c] - 1] for c in range(3)])
    for state in ("init", "update"):
        B, S = K.extract_borders(CM, borders, state, ne)
        Bo, So = o.extract_borders(new, CM, borders, state, ne)
        assert B.shape == Bo.shape, (state, B.shape, Bo.shape)
        assert rel(B, Bo) <= 1e-13 and rel(S, So) <= 1e-13, state
    with pytest.raises(sf.SmearFEMError):
        K.extract_borders(CM, borders, "other", ne)
    with pytest.raises(sf.SmearFEMError):
        K.extract_borders(CM, borders, "init")
    K.free()
    mesh.free()


def test_project_nodes_matches_postprocess_restatement():
    """motion, NodeList_new and back_project of the border nodes on the device
    (examples/vector3D.jl:325-329, src/PostProcess.jl:131-152) against the NumPy restatement."""
    ne = 8
    ctx = sf.context()
    r = o.example_problem(ne)
    NL, IEN, ID, top, btm, borders = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    CM = np.array([[8 * 2048 / 7.07, 0.0, 2048 / 2], [0.0, 8 * 1536 / 5.3, 1536 / 2], [0.0, 0.0, 1.0]]).T  # examples/vector3D.jl:281
    ids = np.concatenate([np.asarray(b, dtype=np.int64).ravel() for b in borders])
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
    p3, p2 = K.project_nodes(ids, CM)  # before any solve: motion = 0
    assert np.array_equal(p3, NL[:, ids - 1]) and rel(p2, o.back_project(NL[:, ids - 1], CM)) <= 1e-15
    K.set_dirichlet_zplanes(0.001)
    for mg in (False, True):
        K.use_multigrid(mg)
        q, _, _ = K.pcg_solve(rtol=1e-13, maxit=5000)
        motion = np.vstack([q[ID[:, c] - 1] for c in range(3)])  # examples/vector3D.jl:325
        new = NL + motion
        p3, p2 = K.project_nodes(ids, CM)
        assert rel(p3, new[:, ids - 1]) <= 1e-15 and rel(p2, o.back_project(new[:, ids - 1], CM)) <= 1e-15
        assert rel(p3, (NL + np.vstack([r["q"][ID[:, c] - 1] for c in range(3)]))[:, ids - 1]) <= 1e-10
    with pytest.raises(sf.SmearFEMError):
        K.project_nodes([0], CM)
    # the reference-shaped call chain with the multigrid switch (general Dirichlet list from setboundaryCond's C)
    Kh = sf.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    K_bar = Kh + 100 * sf.apply_boundary_conditions(ne, NL, IEN, top, btm, 3, "Q1", ID)
    q_d, C = sf.setboundaryCond(NL, ne, 3, "Q1", 0.001, 3)
    qm, info = sf.solve(K_bar, q_d, C, rtol=1e-13, return_info=True, multigrid=True)
    assert rel(qm, r["q"]) <= TOL and info["iters"] <= 40


def test_manufactured_solution_general_dirichlet():
    """u* ~ N(0,1) (seed 4321), rhs = K̄u*; Dirichlet on an arbitrary dof set."""
    ne = 12
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
    n = K.shape[0]
    rng = np.random.default_rng(4321)
    u = rng.standard_normal(n)
    fixed = np.sort(rng.choice(n, size=n // 10, replace=False)) + 1
    K.set_dirichlet(fixed, u[fixed - 1])
    # rhs on free rows: (K̄ u*)_f ; the solver subtracts K̄ q_d itself
    K.set_spmv_variant(0)
    rhs = K.spmv(u)
    q, it, relres = K.pcg_solve(rtol=1e-13, maxit=20000, rhs_extra=rhs)
    assert rel(q, u) <= 1e-9


# ---- error behaviour at the boundary -----------------------------------------------------------------
def test_error_codes():
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 3, 3)
    with pytest.raises(sf.SmearFEMError):  # ne^ndim != number of elements
        sf.assemble_system(4, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    with pytest.raises(sf.SmearFEMError):  # ID missing for nDof > 1
        sf.assemble_system(3, NL, IEN, 3, "Q1", 3, None, 40, 0.4)
    bad = IEN.copy()
    bad[0, 0] = NL.shape[1] + 5
    with pytest.raises(sf.SmearFEMError):  # BoundsError
        sf.assemble_system(3, NL, bad, 3, "Q1", 3, ID, 40, 0.4)
    IDbad = ID.copy()
    IDbad[0, 0] = IDbad[1, 0]
    with pytest.raises(sf.SmearFEMError):  # not a bijection
        sf.assemble_system(3, NL, IEN, 3, "Q1", 3, IDbad, 40, 0.4)
    with pytest.raises(sf.SmearFEMError):  # Q2 is 2-D scalar only
        sf.assemble_system(3, NL, IEN, 3, "Q2", 3, ID, 40, 0.4)
    with pytest.raises(sf.SmearFEMError):  # 1-D branch is not executable upstream either
        sf.assemble_system(3, NL[:1], IEN[:, :2], 1)
    # unconstrained K is singular: PCG on rhs in the null-space complement still runs, but a zero matrix breaks down
    K = sf.assemble_system(3, NL, IEN, 3, "Q1", 3, ID, 0.0, 0.4)
    K.set_dirichlet(np.array([1]), np.array([1.0]))
    q, it, relres = K.pcg_solve(rtol=1e-10, maxit=50)
    assert np.all(np.isfinite(q))


# ---- size-independent properties at the BASELINE single-GPU size (config C3: 100^3) ----------------------------
def test_full_size_100_properties():
    """No oracle can assemble 100^3 in seconds; check properties the domain offers instead:
    nnz formula, symmetry (x'Ky == y'Kx), the 6 rigid-body modes in the null space, diag > 0,
    agreement of the SpMV variants, determinism, and a manufactured-solution solve."""
    ne = 100
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    n = K.shape[0]
    assert n == 3 * 101**3 and K.nnz == 9 * (3 * 101 - 2) ** 3 == 245438109
    d = K.diag()
    assert np.all(d > 0)
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Kx = K.spmv(x)
    Ky = K.spmv(y)
    assert abs(y @ Kx - x @ Ky) <= 1e-12 * abs(y @ Kx)
    for v in (2, 1):
        K.set_spmv_variant(v)
        assert rel(K.spmv(x), Kx) <= 1e-13
    K.set_spmv_variant(4)
    NL = mesh.nodelist()
    X, Y, Z = NL
    zero, one = np.zeros_like(X), np.ones_like(X)
    scale = np.abs(d).max()
    for mx, my, mz in [(one, zero, zero), (zero, one, zero), (zero, zero, one), (-Y, X, zero), (zero, -Z, Y), (Z, zero, -X)]:
        mode = np.column_stack([mx, my, mz]).ravel()
        assert np.abs(K.spmv(mode)).max() <= 1e-10 * scale
    # determinism of the fused assembly at full size (checksum of the diagonal + of K x)
    K.reassemble(40, 0.4)
    assert np.array_equal(K.diag(), d) and np.array_equal(K.spmv(x), Kx)
    # manufactured solution through the example's boundary conditions: q = q_d + x with rhs = K̄ u*
    K.add_surface_mass(100.0).set_dirichlet_zplanes(0.001)
    u = rng.standard_normal(n)
    btm = NL[2] == 0.0
    top = NL[2] == 1.0
    u[2::3][btm] = 0.0
    u[2::3][top] = -0.001
    rhs = K.spmv(u)
    q, it, relres = K.pcg_solve(rtol=1e-12, maxit=8000, rhs_extra=rhs)
    assert relres <= 2e-12 and rel(q, u) <= 1e-8  # relres is the TRUE residual ||b - K q|| / ||b||, recomputed at exit
    # the same manufactured problem through the multigrid-preconditioned CG (6 levels: 100 -> 50 -> 25 -> 13 -> 7 -> 4)
    K.use_multigrid(True)
    qg, itg, relg = K.pcg_solve(rtol=1e-12, maxit=200, rhs_extra=rhs)
    assert relg <= 2e-12 and rel(qg, u) <= 1e-8 and itg <= 60 and itg < it / 10, (itg, it)
    K.use_multigrid(False)
    # a SMOOTH manufactured field (what a displacement solution looks like; white noise above is the worst case for
    # ||u - u*||, which is bounded by cond(K) x residual): the north-star bar ||u - u*|| / ||u*|| <= 1e-10 at 100^3
    us = np.column_stack([0.01 * np.sin(np.pi * X) * np.cos(2 * Y) * Z, 0.01 * np.cos(X) * np.sin(np.pi * Y) * (1 + Z),
                          -0.001 * Z + 0.02 * np.sin(np.pi * Z) * (1 + X * Y)]).ravel()
    rhs_s = K.spmv(us)
    for mg in (False, True):
        K.use_multigrid(mg)
        qs, its, rels = K.pcg_solve(rtol=1e-13, maxit=12000, rhs_extra=rhs_s)
        print(f"100^3 smooth manufactured solution, multigrid={mg}: iters {its}, true relres {rels:.2e}, ||u-u*||/||u*|| = {rel(qs, us):.2e}")
        assert rel(qs, us) <= TOL


def test_load_stepping_warm_start():
    """examples/vector3D.jl:310-338: 50 load steps d = 0.001:0.01:0.5 on one assembled K̄; q is linear in d, so the
    warm-started solves must converge in (almost) no iterations and still match the oracle scaled by d."""
    ne = 8
    ctx = sf.context()
    r = o.example_problem(ne, d=0.001)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
    deltas = np.arange(0.001, 0.5, 0.01)  # Julia's 0.001:0.01:0.5
    assert len(deltas) == 50
    iters = []
    for d, q, it in sf.load_steps(K, deltas, rtol=1e-13):
        assert rel(q, r["q"] * (d / 0.001)) <= TOL
        iters.append(it)
    assert iters[0] > 50 and max(iters[1:]) <= 25  # first solve cold, the rest start converged (<= one graph chunk)


@pytest.mark.parametrize("ne", [1, 3, 10, 23])
def test_general_path_gather_assembly(monkeypatch, ne):
    """General hex meshes with the standard dof map are assembled in gather form (every entry written once by the warp that owns its
    row: no memset, no atomics, no colours): pattern bit-exact, values <= 1e-13 vs the oracle on a jittered lattice with shuffled
    elements, bit-reproducible, every entry rewritten (buffers cleared to 0xFF first), and equal to rounding to the scatter forms."""
    ctx = sf.context()
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    if ne > 1:
        o.jitter_nodes(NL, ne, seed=9)
    perm = np.random.default_rng(ne).permutation(IEN.shape[0])
    mesh = sf.Mesh.from_host(ctx, NL, IEN[perm], ID, 3, 3, ne)
    assert not mesh.info()["structured"] or ne == 1
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    Ko = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    assert_csc_parity(K, Ko, tol=1e-13)
    nz1, d1 = K.to_csc()[2], K.diag()
    assert rel(d1, Ko.to_scipy().diagonal()) <= 1e-13
    for _ in range(2):
        K.assemble_values(40, 0.4)
        assert np.array_equal(K.to_csc()[2], nz1) and np.array_equal(K.diag(), d1)
    for mode in ("colored", "atomic"):
        monkeypatch.setenv("SMFEM_VALUES", mode)
        K.assemble_values(40, 0.4)
        assert rel(K.to_csc()[2], nz1) <= 1e-14, mode
    monkeypatch.delenv("SMFEM_VALUES")
    K.free()
    mesh.free()
