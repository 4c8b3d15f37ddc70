"""Analytic / cross-form known answers that pin the oracle where the reference has no tests
(SURVEY.md 8c table).  Tolerances: 1e-12 relative on norms; patterns bit-exact between forms."""
import numpy as np
import pytest

from oracle import c_oracle, fem_oracle as o


def _load_csc(g, p):
    return o.JuliaCSC(int(g[p + "_m"]), int(g[p + "_n"]), g[p + "_colptr"], g[p + "_rowval"], g[p + "_nzval"])


def test_single_unit_cube_hex_analytic():
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 1, 3)
    Ke = o.element_matrices_literal(1, NL, IEN, 3, "Q1", 3, 1, 0.3)[0]
    E, nu = 1.0, 0.3
    assert Ke[0, 0] == pytest.approx(E / ((1 + nu) * (1 - 2 * nu)) * ((1 - nu) + (1 - 2 * nu)) / 9, rel=1e-14)
    assert Ke[0, 1] == pytest.approx(0.08012820512820512, rel=1e-13)
    assert Ke[0, 3] == pytest.approx(-0.10683760683760685, rel=1e-13)
    assert np.linalg.norm(Ke) == pytest.approx(1.7240292952089151, rel=1e-13)
    ev = np.linalg.eigvalsh(Ke)
    assert (np.abs(ev) < 1e-12).sum() == 6  # rigid-body modes
    assert np.abs(Ke - Ke.T).max() < 1e-15


def test_scalar_laplace_cube_ne2():
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 2, 3)
    K = o.assemble_system_literal(2, NL, IEN, 3, "Q1", 1)
    assert (K.m, K.n, K.nnz) == (27, 27, 343)
    assert K.get(1, 1) == pytest.approx(1 / 6, rel=1e-14)  # h/3
    assert K.get(14, 14) == pytest.approx(4 / 3, rel=1e-14)
    assert np.abs(K.to_scipy().sum(axis=1)).max() < 1e-14  # constants in the null space


def test_hex_elasticity_cube_ne2_known_values():
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 2, 3)
    K = o.assemble_system_literal(2, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
    assert (K.m, K.nnz) == (81, 3087)
    assert K.colptr[:8].tolist() == [1, 25, 49, 73, 109, 145, 181, 205]
    rows_col1 = K.rowval[:24].tolist()
    assert rows_col1 == list(range(1, 7)) + list(range(10, 16)) + list(range(28, 34)) + list(range(37, 43))
    assert np.linalg.norm(K.nzval) == pytest.approx(224.21231906646884, rel=1e-13)
    assert K.to_scipy().diagonal().sum() == pytest.approx(1219.0476190476188, rel=1e-13)
    assert K.get(1, 1) == pytest.approx(6.349206349206349, rel=1e-13)
    assert K.get(40, 40) == pytest.approx(50.7936507936508, rel=1e-13)
    assert K.get(1, 2) == pytest.approx(2.976190476190476, rel=1e-13)
    # explicit zeros are kept (Julia sparse() semantics)
    assert (K.nzval == 0.0).sum() > 0


def test_plane_stress_ne2_known_values():
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 2, 2)
    K = o.assemble_system_literal(2, NL, IEN, 2, "Q1", 2, ID, 40, 0.4)
    assert (K.m, K.nnz) == (18, 196)
    assert np.linalg.norm(K.nzval) == pytest.approx(214.72911210355724, rel=1e-13)
    assert K.get(1, 1) == pytest.approx(20.634920634920633, rel=1e-13)
    assert K.get(1, 2) == pytest.approx(8.333333333333332, rel=1e-13)
    assert K.get(9, 9) == pytest.approx(82.53968253968253, rel=1e-13)


@pytest.mark.parametrize("name", ["hex_ne2_cube", "hex_ne2_inflated", "hex_ne4_cube", "hex_ne4_inflated"])
def test_forms_agree_with_golden(golden_dir, name):
    g = np.load(f"{golden_dir}/{name}.npz")
    ne = int(g["ne"])
    Kg = _load_csc(g, "K")
    NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    if name.endswith("inflated"):
        o.inflate_sphere(NL, 0, 1, 0, 1)
    assert np.array_equal(NL, g["NodeList"]) and np.array_equal(IEN, g["IEN"]) and np.array_equal(ID, g["ID"])
    for K in (o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4),
              c_oracle.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=2)):
        assert np.array_equal(K.colptr, Kg.colptr) and np.array_equal(K.rowval, Kg.rowval)
        assert np.linalg.norm(K.nzval - Kg.nzval) <= 1e-13 * np.linalg.norm(Kg.nzval)
    b = o.apply_boundary_conditions(ne, NL, IEN, top, btm, 3, "Q1", ID)
    bg = _load_csc(g, "b")
    assert np.array_equal(b.rowval, bg.rowval) and np.allclose(b.nzval, bg.nzval, rtol=1e-14, atol=0)
    # b: total = 3 dofs * (top area + bottom area)
    area2 = b.nzval.sum() / 3
    assert area2 == pytest.approx(2.0 if name.endswith("cube") else area2, rel=1e-13)


def test_nnz_formulas_and_rigid_body_null_space():
    for ne in (2, 3, 5):
        NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
        o.inflate_sphere(NL, 0, 1, 0, 1)
        K = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
        assert K.nnz == 9 * (3 * (ne + 1) - 2) ** 3
        A = K.to_scipy()
        # 3 translations + 3 rotations
        x, y, z = NL
        zero, one = np.zeros_like(x), np.ones_like(x)
        modes = [(one, zero, zero), (zero, one, zero), (zero, zero, one), (-y, x, zero), (zero, -z, y), (z, zero, -x)]
        for mx, my, mz in modes:
            v = np.column_stack([mx, my, mz]).ravel()
            assert np.abs(A @ v).max() < 1e-11 * np.abs(K.nzval).max()
        assert abs(A - A.T).max() < 1e-13 * np.abs(K.nzval).max()


def test_example_summaries(golden_dir):
    g = np.load(f"{golden_dir}/example_summaries.npz")
    s8 = g["ne8"]
    expect = [140625, 581.982575824454, 19595.324305251448, 0.11438942801050471, 4.67863726375369, 2025,
              0.016601620624313463, 2.131993307652083e-4]
    assert np.allclose(s8, expect, rtol=1e-11, atol=0)
    r = o.example_problem(4)
    g4 = np.load(f"{golden_dir}/hex_ne4_inflated.npz")
    assert np.linalg.norm(r["q"] - g4["q"]) <= 1e-12 * np.linalg.norm(g4["q"])
    # q is exactly linear in d
    r2 = o.example_problem(4, d=0.011)
    assert np.linalg.norm(r2["q"] - 11 * r["q"]) <= 1e-12 * np.linalg.norm(r2["q"])


def test_setboundarycond_and_patch():
    ne = 3
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    q_d, free = o.setboundaryCond(NL, ne, 3, "Q1", 0.25, 3)
    assert q_d.shape == (3 * 64, 1) and len(free) == 3 * 64 - 2 * 16
    assert np.count_nonzero(q_d) == 16 and np.all(q_d[q_d != 0] == -0.25)
    # linear patch test: with ALL boundary dofs prescribed to a linear field the interior follows it
    K = o.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4).to_scipy().tocsr()
    A = np.array([[0.01, 0.002, -0.003], [0.0, -0.02, 0.004], [0.005, 0.001, 0.03]])
    u = (A @ NL).T.ravel()
    k, j, i = np.meshgrid(np.arange(ne + 1), np.arange(ne + 1), np.arange(ne + 1), indexing="ij")
    interior = ((i > 0) & (i < ne) & (j > 0) & (j < ne) & (k > 0) & (k < ne)).ravel()
    fdof = np.repeat(interior, 3)
    rhs = -(K[fdof][:, ~fdof] @ u[~fdof])
    uf = np.linalg.solve(K[fdof][:, fdof].toarray(), rhs)
    assert np.abs(uf - u[fdof]).max() < 1e-13


def test_back_project_known_answer():
    """src/PostProcess.jl:131-152 with the example's CameraMatrix (examples/vector3D.jl:281): the origin maps to
    (0, -0.5, 2) in the camera frame, i.e. to (cx, cy - fy/4) in the image; a point on the optical axis maps to (cx, cy)."""
    fx, fy, cx, cy = 8 * 2048 / 7.07, 8 * 1536 / 5.3, 2048 / 2, 1536 / 2
    CM = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]]).T
    P = np.array([[0.0, 0.0, 0.0], [0.0, 1.0, 0.5], [0.25, 0.0, 0.5]]).T
    out = o.back_project(P, CM)
    assert out.shape == (2, 3)
    assert np.allclose(out[:, 0], [cx, cy - fy / 4], rtol=0, atol=1e-12)
    assert np.allclose(out[:, 1], [cx, cy], rtol=0, atol=1e-12)          # (0, 1, 0.5) -> camera (0, 0, 1)
    assert np.allclose(out[:, 2], [cx + fx * 0.125, cy], rtol=0, atol=1e-12)  # (0.25, 0, 0.5) -> camera (0.25, 0, 2)


@pytest.mark.parametrize("ne", [2, 8, 16])
def test_plane_stress_problem_c1(ne):
    """Config C1 (2-D Q4 plane stress, assemble + solve on CPU): the solve satisfies the system and the physics."""
    r = o.plane_stress_problem(ne)
    A = r["K"].to_scipy().tocsr()
    q, f = r["q"], r["free"] - 1
    assert np.abs((A @ q)[f]).max() <= 1e-12 * np.abs(A @ r["q_d"][:, 0]).max()     # equilibrium on the free dofs
    assert np.array_equal(q[r["fixed"] - 1], r["q_d"][r["fixed"] - 1, 0])                # prescribed values kept
    uy = q[1::2].reshape(ne + 1, ne + 1)                                               # [j, i]
    assert np.all(np.diff(uy.mean(axis=1)) < 0)                                        # compression grows monotonically in y
    assert np.allclose(uy, uy[:, ::-1], atol=1e-15) and np.allclose(q[0::2].reshape(ne + 1, ne + 1), -q[0::2].reshape(ne + 1, ne + 1)[:, ::-1], atol=1e-15)
    assert r["K"].nnz == 4 * (3 * (ne + 1) - 2) ** 2
    if ne == 8:
        assert (r["K"].m, r["K"].nnz) == (162, 2500)                                   # SURVEY 8(a) sizes for C1


@pytest.mark.parametrize("ne", [8, 16])
def test_plane_stress_problem_matches_golden(golden_dir, ne):
    g = np.load(f"{golden_dir}/c1_plane_stress_ne{ne}.npz")
    r = o.plane_stress_problem(ne)
    assert np.array_equal(r["K"].colptr, g["K_colptr"]) and np.array_equal(r["K"].rowval, g["K_rowval"])
    assert np.array_equal(r["fixed"], g["fixed"]) and np.array_equal(r["free"], g["free"])
    assert np.linalg.norm(r["q"] - g["q"]) <= 1e-13 * np.linalg.norm(g["q"])


def test_extract_borders_restatement_properties():
    """src/PostProcess.jl:60-117 restated (oracle.extract_borders): "init" returns the (ne+1) left extremes, the sorted top arc, the
    reversed right extremes and the reversed sorted bottom arc; "update" returns a strictly convex counter-clockwise polygon that
    starts at the lexicographically smallest point and contains every projected side node (scipy's Qhull gives the same vertex set)."""
    from scipy.spatial import ConvexHull

    ne = 8
    NL, IEN, ID, top, btm, BL = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    CM = np.array([[8 * 2048 / 7.07, 0.0, 2048 / 2], [0.0, 8 * 1536 / 5.3, 1536 / 2], [0.0, 0.0, 1.0]]).T  # examples/vector3D.jl:281
    B, S = o.extract_borders(NL, CM, BL, "init", ne)
    n1 = ne + 1
    assert S.shape == (2, len(BL[0])) and len(BL[0]) % n1 == 0
    sz = S.shape[1] // n1
    for l in range(n1):
        lay = S[:, l * sz:(l + 1) * sz]
        assert B[0, l] == lay[0].min()
    nt = B.shape[1] - 2 * n1
    assert nt > 0
    topseg = B[:, n1:n1 + (S[1, ne * sz:] > B[1, ne]).sum()]
    assert np.all(np.diff(topseg[0]) >= 0)                         # sortslices: ascending x
    right = B[:, n1 + topseg.shape[1]:2 * n1 + topseg.shape[1]]
    for l in range(n1):
        assert right[0, n1 - 1 - l] == S[0, l * sz:(l + 1) * sz].max()   # reverse(RightborderNodes)
    H, S2 = o.extract_borders(NL, CM, BL, "update")
    assert np.array_equal(S, S2)
    assert tuple(H[:, 0]) == min(map(tuple, S.T))
    x, y = H
    cross = (np.roll(x, -1) - x) * (np.roll(y, -2) - np.roll(y, -1)) - (np.roll(y, -1) - y) * (np.roll(x, -2) - np.roll(x, -1))
    assert np.all(cross > 0)                                        # strictly convex, counter-clockwise
    qh = ConvexHull(S.T)
    assert {tuple(p) for p in H.T} <= {tuple(p) for p in S.T[qh.vertices]} | {tuple(p) for p in H.T}
    # every point inside or on the hull
    for k in range(H.shape[1]):
        a, b = H[:, k], H[:, (k + 1) % H.shape[1]]
        side = (b[0] - a[0]) * (S[1] - a[1]) - (b[1] - a[1]) * (S[0] - a[0])
        assert side.min() >= -1e-9 * np.abs(S).max() ** 2
    with pytest.raises(NameError):
        o.extract_borders(NL, CM, BL, "other")


def test_julia_sparse_restatement_semantics():
    """`sparse(I, J, V)` as the call sites src/fem.jl:253 and examples/vector3D.jl:262 rely on it (Julia stdlib SparseArrays, not
    under /root/reference): dims = (max I, max J), duplicates folded with + in INPUT order, numerical zeros stay structural,
    rows ascending per column, empty columns allowed, empty input = the 0 x 0 matrix."""
    # a fold whose result depends on the order: ((1e16 + 1) - 1e16) = 0 in input order, (1e16 - 1e16) + 1 = 1 otherwise
    E = [3, 1, 3, 3, 2, 5, 5]
    J = [2, 1, 2, 2, 4, 4, 4]
    V = [1e16, 7.0, 1.0, -1e16, 0.0, 2.5, -2.5]
    A = o.julia_sparse(E, J, V)
    assert (A.m, A.n) == (5, 4)
    assert list(A.colptr) == [1, 2, 3, 3, 5]  # column 3 is empty
    assert list(A.rowval) == [1, 3, 2, 5]
    assert list(A.nzval) == [7.0, (1e16 + 1.0) - 1e16, 0.0, 0.0]  # explicit zero (2,4) and cancelled pair (5,4) are both stored
    assert A.get(3, 2) == 0.0 and A.get(5, 4) == 0.0 and A.get(4, 4) == 0.0 and A.nnz == 4
    # against a naive dictionary fold on random triplets with many duplicates
    rng = np.random.default_rng(7)
    E = rng.integers(1, 9, 400)
    J = rng.integers(1, 7, 400)
    V = rng.standard_normal(400) * 10.0 ** rng.integers(-8, 9, 400)
    A = o.julia_sparse(E, J, V)
    acc = {}
    for e, j, v in zip(E, J, V):
        acc[(j, e)] = acc.get((j, e), 0.0) + v  # left-to-right
    keys = sorted(acc)
    assert [k[1] for k in keys] == list(A.rowval)
    assert [acc[k] for k in keys] == list(A.nzval)  # bitwise
    assert all(A.colptr[j] - A.colptr[j - 1] == sum(1 for k in keys if k[0] == j) for j in range(1, A.n + 1))
    Z = o.julia_sparse([], [], [])
    assert (Z.m, Z.n, Z.nnz, list(Z.colptr)) == (0, 0, 0, [1])
    with pytest.raises(ValueError):
        o.julia_sparse([0, 1], [1, 1], [1.0, 2.0])
    with pytest.raises(ValueError):
        o.julia_sparse([1, 2], [1], [1.0, 2.0])
