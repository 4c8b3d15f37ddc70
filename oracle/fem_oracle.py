"""CPU oracle for the smearFEM.jl hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module; the product (`smearfem.jl_b200`) never does.

It is a restatement, in NumPy, of the reference's algorithm (all citations relative to
/root/reference):

    gaussian_quadrature          src/fem.jl:21-31
    basis_function               src/fem.jl:48-114
    assemble_system              src/fem.jl:135-256
    sparse(E,J,V)                Julia stdlib SparseArrays (NOT under /root/reference; un-pinned,
                                 Project.toml:15 has no [compat] entry) -- restated from its
                                 documented behaviour: dims = (max(E), max(J)), duplicates summed
                                 in input order, numerical zeros kept, rows ascending per column.
    meshgrid                     examples/vector3D.jl:10-130
    setboundaryCond              examples/vector3D.jl:133-173
    apply_boundary_conditions    examples/vector3D.jl:175-264
    solve idiom                  examples/vector3D.jl:308-322
    inflate_sphere               src/PostProcess.jl:30-44

PARITY STATUS: the reference's own tests pin ONLY basis_function at element corners and the
Gauss nodes/weights (test/runtests.jl:15-35); those are reproduced bit-for-bit by
tests/test_oracle_reference_tests.py.  For Ke / K / b / q the reference has no golden vectors and
Julia is not installed in this image, so for those quantities this oracle is **parity unpinned**:
it is cross-checked only against (i) an independent second restatement (the literal loop form vs
the vectorised form vs the C form in oracle/c/), and (ii) analytic known answers
(tests/test_oracle_known_answers.py).

All array shapes / index bases are Julia's: NodeList (ndim, nNodes) float64; IEN (nEl, nLocal)
int64 1-based; ID (nNodes, nDof) int64 1-based; sparse results are CSC (colptr, rowval, nzval)
with 1-based int64 indices.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# --------------------------------------------------------------------------------------
# src/fem.jl:21-31
# --------------------------------------------------------------------------------------


def gaussian_quadrature(a, b, nGaussPoints=2):
    """src/fem.jl:21-31.  Expression order is kept literally so that the reference's `==`
    tests (test/runtests.jl:34-35) hold bit-for-bit.  Any other nGaussPoints leaves xi undefined
    in the reference (UndefVarError) -> here: NameError-like ValueError."""
    a = float(a)
    b = float(b)
    if nGaussPoints == 2:
        xi = [-(b - a) / (2 * math.sqrt(3)) + (b + a) / 2, (b - a) / (2 * math.sqrt(3)) + (b + a) / 2]
        w = [(b - a) / 2, (b - a) / 2]
    elif nGaussPoints == 3:
        # NOTE the literal 0 midpoint of the reference (src/fem.jl:27), not (a+b)/2.
        xi = [-(b - a) / (2 * math.sqrt(5 / 3)) + (b + a) / 2, 0.0, (b - a) / (2 * math.sqrt(5 / 3)) + (b + a) / 2]
        w = [(b - a) / 2 * 5 / 9, (b - a) / 2 * 8 / 9, (b - a) / 2 * 5 / 9]
    else:
        raise ValueError("UndefVarError: xi not defined (reference supports nGaussPoints in {2,3} only)")
    return np.array(xi, dtype=np.float64), np.array(w, dtype=np.float64)


# --------------------------------------------------------------------------------------
# src/fem.jl:48-114
# --------------------------------------------------------------------------------------


def basis_function(xi, eta=None, zeta=None, FunctionClass="Q1"):
    """src/fem.jl:48-114.  Returns (N, dN) with dN of shape (nnodes, ndim); the 1-D branch
    returns the reference's 1x2 row matrix (src/fem.jl:75)."""
    if FunctionClass == "Q1":
        if zeta is not None:  # src/fem.jl:51-63
            x, e, z = float(xi), float(eta), float(zeta)
            N = [
                (1 - x) * (1 - e) * (1 - z) / 8,
                (1 + x) * (1 - e) * (1 - z) / 8,
                (1 + x) * (1 + e) * (1 - z) / 8,
                (1 - x) * (1 + e) * (1 - z) / 8,
                (1 - x) * (1 - e) * (1 + z) / 8,
                (1 + x) * (1 - e) * (1 + z) / 8,
                (1 + x) * (1 + e) * (1 + z) / 8,
                (1 - x) * (1 + e) * (1 + z) / 8,
            ]
            dx = [
                -(1 - e) * (1 - z) / 8, (1 - e) * (1 - z) / 8, (1 + e) * (1 - z) / 8, -(1 + e) * (1 - z) / 8,
                -(1 - e) * (1 + z) / 8, (1 - e) * (1 + z) / 8, (1 + e) * (1 + z) / 8, -(1 + e) * (1 + z) / 8,
            ]
            dy = [
                -(1 - x) * (1 - z) / 8, -(1 + x) * (1 - z) / 8, (1 + x) * (1 - z) / 8, (1 - x) * (1 - z) / 8,
                -(1 - x) * (1 + z) / 8, -(1 + x) * (1 + z) / 8, (1 + x) * (1 + z) / 8, (1 - x) * (1 + z) / 8,
            ]
            dz = [
                -(1 - x) * (1 - e) / 8, -(1 + x) * (1 - e) / 8, -(1 + x) * (1 + e) / 8, -(1 - x) * (1 + e) / 8,
                (1 - x) * (1 - e) / 8, (1 + x) * (1 - e) / 8, (1 + x) * (1 + e) / 8, (1 - x) * (1 + e) / 8,
            ]
            return np.array(N), np.column_stack([dx, dy, dz])
        elif eta is not None:  # src/fem.jl:64-69
            x, e = float(xi), float(eta)
            N = [(1 - x) * (1 - e) / 4, (x + 1) * (1 - e) / 4, (1 + x) * (e + 1) / 4, (1 - x) * (1 + e) / 4]
            dx = [-(1 - e) / 4, (1 - e) / 4, (e + 1) / 4, -(1 + e) / 4]
            dy = [-(1 - x) / 4, -(x + 1) / 4, (1 + x) / 4, (1 - x) / 4]
            return np.array(N), np.column_stack([dx, dy])
        else:  # src/fem.jl:70-75
            x = float(xi)
            N = [0.5 - 0.5 * x, 0.5 + 0.5 * x]
            return np.array(N), np.array([[-0.5, 0.5]])
    elif FunctionClass == "Q2":
        if eta is not None:  # src/fem.jl:78-110
            x, e = float(xi), float(eta)
            N = [
                (1 - x) * x * (1 - e) * e / 4,
                -x * (1 + x) * (1 - e) * e / 4,
                x * (1 + x) * e * (1 + e) / 4,
                -(1 - x) * x * e * (1 + e) / 4,
                -(1 - x) * (1 + x) * (1 - e) * e / 2,
                x * (1 + x) * (1 - e) * (1 + e) / 2,
                (1 - x) * (1 + x) * e * (1 + e) / 2,
                -(1 - x) * x * (1 - e) * (1 + e) / 2,
                (1 - x) * (1 + x) * (1 - e) * (1 + e),
            ]
            dx = [
                (1 - 2 * x) * (1 - e) * e / 4,
                -(1 + 2 * x) * (1 - e) * e / 4,
                (1 + 2 * x) * e * (1 + e) / 4,
                -(1 - 2 * x) * e * (1 + e) / 4,
                x * (1 - e) * e,
                (1 + 2 * x) * (1 - e) * (1 + e) / 2,
                -x * e * (1 + e),
                -(1 - 2 * x) * (1 - e) * (1 + e) / 2,
                -2 * x * (1 - e) * (1 + e),
            ]
            dy = [
                (1 - x) * x * (1 - 2 * e) / 4,
                -x * (1 + x) * (1 - 2 * e) / 4,
                x * (1 + x) * (1 + 2 * e) / 4,
                -(1 - x) * x * (1 + 2 * e) / 4,
                -(1 - x) * (1 + x) * (1 - 2 * e) / 2,
                -x * (1 + x) * e,
                (1 - x) * (1 + x) * (1 + 2 * e) / 2,
                (1 - x) * x * e,
                -(1 - x) * (1 + x) * 2 * e,
            ]
            return np.array(N), np.column_stack([dx, dy])
    raise ValueError("UndefVarError: N not defined for this FunctionClass / dimension")


# --------------------------------------------------------------------------------------
# Julia stdlib sparse(E, J, V)
# --------------------------------------------------------------------------------------


@dataclass
class JuliaCSC:
    """A SparseMatrixCSC{Float64,Int64} look-alike: 1-based colptr / rowval."""

    m: int
    n: int
    colptr: np.ndarray  # int64, n+1, 1-based
    rowval: np.ndarray  # int64, nnz, 1-based, ascending inside each column
    nzval: np.ndarray  # float64, nnz (explicit zeros kept)

    @property
    def nnz(self):
        return int(self.rowval.shape[0])

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.m, self.n))

    def get(self, i, j):
        """1-based K[i,j] (0.0 if not stored)."""
        lo, hi = self.colptr[j - 1] - 1, self.colptr[j] - 1
        rows = self.rowval[lo:hi]
        p = np.searchsorted(rows, i)
        if p < rows.shape[0] and rows[p] == i:
            return float(self.nzval[lo + p])
        return 0.0


def julia_sparse(E, J, V):
    """Restatement of Julia's `sparse(I, J, V)` (call sites src/fem.jl:253,
    examples/vector3D.jl:262): size (max(I), max(J)); duplicates combined with `+` in input
    order; numerical zeros kept; row indices ascending in each column."""
    E = np.asarray(E, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    V = np.asarray(V, dtype=np.float64)
    if not (E.shape == J.shape == V.shape):
        raise ValueError("sparse(I, J, V): the three vectors must have the same length")  # Julia: ArgumentError
    if E.size == 0:  # sparse(Int[], Int[], Float64[]) is the 0 x 0 matrix
        return JuliaCSC(0, 0, np.ones(1, dtype=np.int64), np.zeros(0, dtype=np.int64), np.zeros(0))
    if E.min() < 1 or J.min() < 1:
        raise ValueError("sparse(I, J, V): indices are 1-based")  # Julia: ArgumentError
    m, n = int(E.max()), int(J.max())
    key = (J - 1) * m + (E - 1)
    order = np.argsort(key, kind="stable")  # stable -> input order kept inside a duplicate group
    ks = key[order]
    first = np.ones(ks.shape[0], dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    slot = np.cumsum(first) - 1  # target slot of every sorted triplet
    nnz = int(slot[-1]) + 1
    nzval = np.zeros(nnz, dtype=np.float64)
    np.add.at(nzval, slot, V[order])  # ufunc.at is sequential: left-to-right fold in input order
    ukeys = ks[first]
    rowval = (ukeys % m) + 1
    cols = ukeys // m
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr) + 1
    return JuliaCSC(m, n, colptr.astype(np.int64), rowval.astype(np.int64), nzval)


# --------------------------------------------------------------------------------------
# examples/vector3D.jl:10-130
# --------------------------------------------------------------------------------------


def _julia_range(a, b, n):
    """range(a, b, length=n): Julia's twice-precision range yields the correctly rounded
    a + i*(b-a)/(n-1); for the unit interval that is i/(n-1) exactly rounded, endpoints exact."""
    a = float(a)
    b = float(b)
    i = np.arange(n, dtype=np.float64)
    x = a + (b - a) * (i / (n - 1))
    x[0] = a
    x[-1] = b
    return x


def meshgrid(x0, x1, y0, y1, z0, z1, ne, ndim):
    """examples/vector3D.jl:10-130.  Returns the reference's 6-tuple
    (NodeList, IEN, ID, IEN_top, IEN_btm, [BorderNodes, BottomBorderNodes, TopBorderNodes])."""
    n1 = ne + 1
    nN = n1**ndim
    NodeList = np.zeros((ndim, nN))
    IEN = np.zeros((ne**ndim, 2**ndim), dtype=np.int64)
    IEN_top = np.zeros((ne ** (ndim - 1), 2 ** (ndim - 1)), dtype=np.int64)
    IEN_btm = np.zeros((ne ** (ndim - 1), 2 ** (ndim - 1)), dtype=np.int64)
    m = np.arange(1, nN + 1, dtype=np.int64)
    ID = (ndim * (m[:, None] - 1) + np.arange(1, ndim + 1, dtype=np.int64)[None, :]).astype(np.int64)  # :33,:74
    Border, Bottom, Top = [], [], []
    if ndim == 2:  # :23-58
        x = _julia_range(x0, x1, n1)
        y = _julia_range(y0, y1, n1)
        jj, ii = np.meshgrid(np.arange(n1), np.arange(n1), indexing="ij")  # j slow, i fast
        NodeList[0] = x[ii.ravel()]
        NodeList[1] = y[jj.ravel()]
        ib = (ii.ravel() == 0) | (ii.ravel() == ne)
        Border = list(m[ib])
        ej, ei = np.meshgrid(np.arange(1, ne + 1), np.arange(1, ne + 1), indexing="ij")
        ej, ei = ej.ravel(), ei.ravel()
        IEN[:, 0] = (ej - 1) * n1 + ei
        IEN[:, 1] = (ej - 1) * n1 + ei + 1
        IEN[:, 2] = ej * n1 + ei + 1
        IEN[:, 3] = ej * n1 + ei
        # :49-55  (elseif: when ne == 1 the top list stays zero, as in the reference)
        sel = ej == 1
        IEN_btm[ei[sel] - 1, 0] = IEN[sel, 0]
        IEN_btm[ei[sel] - 1, 1] = IEN[sel, 1]
        sel = (ej == ne) & (ej != 1)
        IEN_top[ei[sel] - 1, 0] = IEN[sel, 3]
        IEN_top[ei[sel] - 1, 1] = IEN[sel, 2]
    elif ndim == 3:  # :60-127
        x = _julia_range(x0, x1, n1)
        y = _julia_range(y0, y1, n1)
        z = _julia_range(z0, z1, n1)
        kk, jj, ii = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
        ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
        NodeList[0] = x[ii]
        NodeList[1] = y[jj]
        NodeList[2] = z[kk]
        side = (ii == 0) | (ii == ne) | (jj == 0) | (jj == ne)
        Border = list(m[side])
        Bottom = list(m[~side & (kk == 0)])
        Top = list(m[~side & (kk == ne) & (kk != 0)])
        ek, ej, ei = np.meshgrid(np.arange(1, ne + 1), np.arange(1, ne + 1), np.arange(1, ne + 1), indexing="ij")
        ei, ej, ek = ei.ravel(), ej.ravel(), ek.ravel()
        p = n1 * n1
        IEN[:, 0] = (ek - 1) * p + (ej - 1) * n1 + ei
        IEN[:, 1] = (ek - 1) * p + (ej - 1) * n1 + ei + 1
        IEN[:, 2] = (ek - 1) * p + ej * n1 + ei + 1
        IEN[:, 3] = (ek - 1) * p + ej * n1 + ei
        IEN[:, 4] = ek * p + (ej - 1) * n1 + ei
        IEN[:, 5] = ek * p + (ej - 1) * n1 + ei + 1
        IEN[:, 6] = ek * p + ej * n1 + ei + 1
        IEN[:, 7] = ek * p + ej * n1 + ei
        sel = ek == 1  # :102-107
        IEN_btm[:, :] = IEN[sel][:, 0:4]
        sel = (ek == ne) & (ek != 1)  # :108-113 (elseif)
        if sel.any():
            IEN_top[:, :] = IEN[sel][:, 4:8]
    return NodeList, IEN, ID, IEN_top, IEN_btm, [Border, Bottom, Top]


# --------------------------------------------------------------------------------------
# src/PostProcess.jl:30-44
# --------------------------------------------------------------------------------------


def inflate_sphere(NodeList, x0, x1, y0, y1):
    """src/PostProcess.jl:30-44.  IN PLACE, like the reference.  `scale ≈ 0.` with Julia's
    default tolerances (atol = 0) is true only for scale == 0."""
    cx = 0.5 * (x0 + x1)
    cy = 0.5 * (y0 + y1)
    dx = NodeList[0] - cx
    dy = NodeList[1] - cy
    scale = np.maximum(np.abs(dx), np.abs(dy))
    r = np.sqrt(dx * dx + dy * dy)
    zero = scale == 0.0
    rs = np.where(zero, 1.0, r)
    NodeList[0] = np.where(zero, 0.0, scale * dx / rs)
    NodeList[1] = np.where(zero, 0.0, scale * dy / rs)
    return NodeList


# --------------------------------------------------------------------------------------
# src/fem.jl:135-256  -- literal loop form (ground truth; small ne only)
# --------------------------------------------------------------------------------------


def _gp_table(ndim):
    """src/fem.jl:149-177."""
    xi, w = gaussian_quadrature(-1, 1)
    if ndim == 1:
        return [(xi[0],), (xi[1],)], [w[0], w[1]]
    if ndim == 2:
        wp = [w[0] * w[0], w[1] * w[0], w[1] * w[1], w[0] * w[1]]
        x = [xi[0], xi[1], xi[1], xi[0]]
        y = [xi[0], xi[0], xi[1], xi[1]]
        return list(zip(x, y)), wp
    wp = [w[0] * w[0] * w[0], w[1] * w[0] * w[0], w[1] * w[1] * w[0], w[0] * w[1] * w[0],
          w[0] * w[0] * w[1], w[1] * w[0] * w[1], w[1] * w[1] * w[1], w[0] * w[1] * w[1]]
    x = [xi[0], xi[1], xi[1], xi[0], xi[0], xi[1], xi[1], xi[0]]
    y = [xi[0], xi[0], xi[1], xi[1], xi[0], xi[0], xi[1], xi[1]]
    z = [xi[0], xi[0], xi[0], xi[0], xi[1], xi[1], xi[1], xi[1]]
    return list(zip(x, y, z)), wp


def constitutive(nDof, Young, nu):
    """src/fem.jl:217 (plane stress) and :230 (3-D isotropic)."""
    Young = float(Young)
    nu = float(nu)
    if nDof == 2:
        return np.array([
            [Young / (1 - nu**2), nu * Young / (1 - nu**2), 0.0],
            [nu * Young / (1 - nu**2), Young / (1 - nu**2), 0.0],
            [0.0, 0.0, Young / (2 * (1 + nu))],
        ])
    s = (1 - 2 * nu) / 2
    return np.array([
        [1 - nu, nu, nu, 0, 0, 0],
        [nu, 1 - nu, nu, 0, 0, 0],
        [nu, nu, 1 - nu, 0, 0, 0],
        [0, 0, 0, s, 0, 0],
        [0, 0, 0, 0, s, 0],
        [0, 0, 0, 0, 0, s],
    ], dtype=np.float64) * (Young / ((1 + nu) * (1 - 2 * nu)))


def _B_matrix(dNdX, nDof):
    """src/fem.jl:211-215 and :219-228."""
    nn = dNdX.shape[0]
    if nDof == 2:
        B = np.zeros((3, 2 * nn))
        B[0, 0::2] = dNdX[:, 0]
        B[1, 1::2] = dNdX[:, 1]
        B[2, 0::2] = dNdX[:, 1]
        B[2, 1::2] = dNdX[:, 0]
        return B
    B = np.zeros((6, 3 * nn))
    B[0, 0::3] = dNdX[:, 0]
    B[1, 1::3] = dNdX[:, 1]
    B[2, 2::3] = dNdX[:, 2]
    B[3, 1::3] = dNdX[:, 2]
    B[3, 2::3] = dNdX[:, 1]
    B[4, 0::3] = dNdX[:, 2]
    B[4, 2::3] = dNdX[:, 0]
    B[5, 0::3] = dNdX[:, 1]
    B[5, 1::3] = dNdX[:, 0]
    return B


def element_matrices_literal(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, Young=1, nu=0.3):
    """Per-element Ke (sum over Gauss points, src/fem.jl:183-233) -- debugging aid."""
    gps, wp = _gp_table(ndim)
    nEl = ne**ndim
    nn = IEN.shape[1]
    out = np.zeros((nEl, nn * nDof, nn * nDof))
    cMat = constitutive(nDof, Young, nu) if nDof > 1 else None
    for e in range(nEl):
        coords = NodeList[:, IEN[e, :] - 1]
        for gp in range(2**ndim):
            N, dN = basis_function(*gps[gp], *([None] * (3 - ndim)), FunctionClass)
            Jac = coords @ dN
            w = wp[gp] * abs(np.linalg.det(Jac))
            dNdX = dN @ np.linalg.inv(Jac)
            if nDof == 1:
                out[e] += w * (dNdX @ dNdX.T)
            else:
                B = _B_matrix(dNdX, nDof)
                out[e] += B.T @ cMat @ B * w
    return out


def assemble_coo_literal(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3):
    """src/fem.jl:135-252: the (E, J, V) triplets, slot order exactly as `inz` (:203, :242)."""
    nn = IEN.shape[1]
    nEl = ne**ndim
    if nDof == 1:
        L = nEl * nn**2
    else:
        L = nEl * (ID.shape[1] * 2**ndim) ** 2
    E = np.zeros(L, dtype=np.int64)
    J = np.zeros(L, dtype=np.int64)
    V = np.zeros(L, dtype=np.float64)
    gps, wp = _gp_table(ndim)
    cMat = constitutive(nDof, Young, nu) if nDof > 1 else None
    for e in range(1, nEl + 1):
        coords = NodeList[:, IEN[e - 1, :] - 1]
        for gp in range(2**ndim):
            N, dN = basis_function(*gps[gp], *([None] * (3 - ndim)), FunctionClass)
            Jac = coords @ dN
            w = wp[gp] * abs(np.linalg.det(Jac))
            invJ = np.linalg.inv(Jac)
            dNdX = dN @ invJ
            if nDof == 1:
                szN = N.shape[0]
                for i in range(1, szN + 1):
                    for j in range(1, szN + 1):
                        inz = szN**2 * (e - 1) + szN * (i - 1) + j
                        E[inz - 1] = IEN[e - 1, i - 1]
                        J[inz - 1] = IEN[e - 1, j - 1]
                        V[inz - 1] += w * float(np.dot(dNdX[i - 1, :], dNdX[j - 1, :]))
            else:
                B = _B_matrix(dNdX, nDof)
                Ke = B.T @ cMat @ B * w
                nK = Ke.shape[0]
                for iNode in range(1, nK // nDof + 1):
                    for jNode in range(1, nK // nDof + 1):
                        for iDof in range(1, ID.shape[1] + 1):
                            for jDof in range(1, ID.shape[1] + 1):
                                i = (iNode - 1) * nDof + iDof
                                j = (jNode - 1) * nDof + jDof
                                inz = Ke.size * (e - 1) + (iNode - 1) * nDof * nK + (jNode - 1) * nDof**2 + (iDof - 1) * nDof + jDof
                                E[inz - 1] = ID[IEN[e - 1, iNode - 1] - 1, iDof - 1]
                                J[inz - 1] = ID[IEN[e - 1, jNode - 1] - 1, jDof - 1]
                                V[inz - 1] += Ke[i - 1, j - 1]
    return E, J, V


def assemble_system_literal(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3):
    """src/fem.jl:135-256, loop for loop."""
    return julia_sparse(*assemble_coo_literal(ne, NodeList, IEN, ndim, FunctionClass, nDof, ID, Young, nu))


# --------------------------------------------------------------------------------------
# src/fem.jl:135-256 -- vectorised form (independent second restatement; ne up to ~64)
# --------------------------------------------------------------------------------------


def element_matrices(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, Young=1, nu=0.3, chunk=32768):
    """Batched Ke = sum_gp B' D B w  (src/fem.jl:179-233), all elements at once."""
    gps, wp = _gp_table(ndim)
    nEl = ne**ndim
    nn = IEN.shape[1]
    nd = nn * nDof
    Ke = np.zeros((nEl, nd, nd))
    cMat = constitutive(nDof, Young, nu) if nDof > 1 else None
    tabs = [basis_function(*gps[g], *([None] * (3 - ndim)), FunctionClass)[1] for g in range(2**ndim)]
    for s in range(0, nEl, chunk):
        idx = IEN[s:s + chunk] - 1
        coords = NodeList[:, idx]  # (ndim, c, nn)
        coords = np.transpose(coords, (1, 0, 2))  # (c, ndim, nn)
        for g in range(2**ndim):
            dN = tabs[g]
            Jac = coords @ dN  # (c, ndim, ndim)
            w = wp[g] * np.abs(np.linalg.det(Jac))
            dNdX = dN[None] @ np.linalg.inv(Jac)  # (c, nn, ndim)
            if nDof == 1:
                Ke[s:s + chunk] += w[:, None, None] * (dNdX @ np.transpose(dNdX, (0, 2, 1)))
            else:
                c = dNdX.shape[0]
                if nDof == 2:
                    B = np.zeros((c, 3, nd))
                    B[:, 0, 0::2] = dNdX[:, :, 0]
                    B[:, 1, 1::2] = dNdX[:, :, 1]
                    B[:, 2, 0::2] = dNdX[:, :, 1]
                    B[:, 2, 1::2] = dNdX[:, :, 0]
                else:
                    B = np.zeros((c, 6, nd))
                    B[:, 0, 0::3] = dNdX[:, :, 0]
                    B[:, 1, 1::3] = dNdX[:, :, 1]
                    B[:, 2, 2::3] = dNdX[:, :, 2]
                    B[:, 3, 1::3] = dNdX[:, :, 2]
                    B[:, 3, 2::3] = dNdX[:, :, 1]
                    B[:, 4, 0::3] = dNdX[:, :, 2]
                    B[:, 4, 2::3] = dNdX[:, :, 0]
                    B[:, 5, 0::3] = dNdX[:, :, 1]
                    B[:, 5, 1::3] = dNdX[:, :, 0]
                Ke[s:s + chunk] += (np.transpose(B, (0, 2, 1)) @ cMat @ B) * w[:, None, None]
    return Ke


def assemble_system(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3):
    """Vectorised src/fem.jl:135-256 -> JuliaCSC."""
    Ke = element_matrices(ne, NodeList, IEN, ndim, FunctionClass, nDof, Young, nu)
    nEl = ne**ndim
    conn = IEN[:nEl]
    if nDof == 1:
        rows = conn  # raw node ids (src/fem.jl:204-205)
    else:
        rows = ID[conn - 1, :].reshape(nEl, -1)  # node-major local dof order (src/fem.jl:240-244)
    nd = rows.shape[1]
    E = np.repeat(rows[:, :, None], nd, axis=2).ravel()
    J = np.repeat(rows[:, None, :], nd, axis=1).ravel()
    # NB: slot order inside an element differs from `inz`, but duplicates of one (i,j) never
    # come from the same element twice on these meshes, so the fold order (ascending e) is kept.
    return julia_sparse(E, J, Ke.ravel())


# --------------------------------------------------------------------------------------
# examples/vector3D.jl:175-264
# --------------------------------------------------------------------------------------


def apply_boundary_conditions(ne, NodeList, IEN, IEN_top, IEN_btm, ndim, FunctionClass, ID, nDof=3):
    """examples/vector3D.jl:175-264 (3-D branch; the reference's 2-D branch is broken).
    Returns the surface 'slip' matrix b = int_{top+bottom} N'N as JuliaCSC."""
    assert ndim == 3, "reference 2-D branch of apply_boundary_conditions is not executable"
    nf = ne ** (ndim - 1)
    nl = IEN_btm.shape[1]
    nd = ID.shape[1] * nl
    L = nf * nd * nd * 2
    E = np.zeros(L, dtype=np.int64)
    J = np.zeros(L, dtype=np.int64)
    V = np.zeros(L, dtype=np.float64)
    xi, w = gaussian_quadrature(-1, 1)
    wp = [w[0] * w[0], w[1] * w[0], w[1] * w[1], w[0] * w[1]]
    x = [xi[0], xi[1], xi[1], xi[0]]
    y = [xi[0], xi[0], xi[1], xi[1]]
    blk = nd * nd
    for which, conn, base in (("btm", IEN_btm, 0), ("top", IEN_top, blk * nf)):
        coords = np.transpose(NodeList[:, conn - 1], (1, 0, 2))  # (nf, 3, 4)
        acc = np.zeros((nf, nd, nd))
        for gp in range(4):
            N, dN = basis_function(x[gp], y[gp], None, FunctionClass)
            t = coords @ dN  # (nf, 3, 2)
            wgt = wp[gp] * np.linalg.norm(np.cross(t[:, :, 0], t[:, :, 1]), axis=1)
            M = np.zeros((3, nd))
            M[0, 0::nDof] = N
            M[1, 1::nDof] = N
            M[2, 2::nDof] = N
            acc += wgt[:, None, None] * (M.T @ M)[None]
        rows = ID[conn - 1, :].reshape(nf, -1)
        sl = slice(base, base + blk * nf)
        E[sl] = np.repeat(rows[:, :, None], nd, axis=2).ravel()
        J[sl] = np.repeat(rows[:, None, :], nd, axis=1).ravel()
        V[sl] = acc.ravel()
    return julia_sparse(E, J, V)


# --------------------------------------------------------------------------------------
# examples/vector3D.jl:133-173
# --------------------------------------------------------------------------------------


def setboundaryCond(NodeList, ne, ndim, FunctionClass, d, nDof=1):
    """examples/vector3D.jl:133-173.  Returns (q_d  (ndof x 1 matrix),  free  = 1-based dof ids
    of the kept columns of C, i.e. C = I[:, free])."""
    ndof = nDof * (ne + 1) ** ndim
    q_d = np.zeros((ndof, 1))
    z = NodeList[2]
    nodes = np.arange(1, NodeList.shape[1] + 1)
    btm = z == 0
    top = (z == 1) & ~btm
    q_d[3 * nodes[top] - 1, 0] = -d
    rCol = np.concatenate([3 * nodes[btm], 3 * nodes[top]])
    free = np.setdiff1d(np.arange(1, ndim * (ne + 1) ** ndim + 1), rCol)
    return q_d, free


# --------------------------------------------------------------------------------------
# examples/vector3D.jl:308-322
# --------------------------------------------------------------------------------------


def add_scaled(K: JuliaCSC, b: JuliaCSC, beta):
    """K_bar = K + beta*b (examples/vector3D.jl:308), evaluated ON K's STORED PATTERN (Julia's
    sparse `+` would additionally drop numerical zeros from the result; values are identical)."""
    out = JuliaCSC(K.m, K.n, K.colptr.copy(), K.rowval.copy(), K.nzval.copy())
    for j in range(1, b.n + 1):
        lo, hi = b.colptr[j - 1] - 1, b.colptr[j] - 1
        if hi == lo:
            continue
        klo, khi = K.colptr[j - 1] - 1, K.colptr[j] - 1
        pos = klo + np.searchsorted(K.rowval[klo:khi], b.rowval[lo:hi])
        assert np.array_equal(K.rowval[pos], b.rowval[lo:hi]), "pattern(b) must be inside pattern(K)"
        out.nzval[pos] += beta * b.nzval[lo:hi]
    return out


def solve_reference(K_bar: JuliaCSC, q_d, free, dense=None):
    """examples/vector3D.jl:315-322:  K_free = C'K̄C ; q_f = inv(K_free) C'(-K̄ q_d) ; q = q_d + C q_f.
    `dense=True` follows the reference literally (dense inverse); otherwise a sparse direct
    solve (same linear system) is used so that 20^3 stays tractable."""
    import scipy.sparse.linalg as spla

    A = K_bar.to_scipy().tocsr()
    f = np.asarray(free) - 1
    rhs = -(A @ q_d[:, 0])[f]
    Aff = A[f][:, f]
    if dense is None:
        dense = f.shape[0] <= 3000
    if dense:
        q_f = np.linalg.inv(Aff.toarray()) @ rhs
    else:
        q_f = spla.spsolve(Aff.tocsc(), rhs)
    q = q_d[:, 0].copy()
    q[f] += q_f
    return q


def example_problem(ne, d=0.001, Young=40, nu=0.4, beta=100, inflate=True, literal=False):
    """The intended pipeline of examples/vector3D.jl:266-322 for one load step."""
    NodeList, IEN, ID, IEN_top, IEN_btm, borders = meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    if inflate:
        inflate_sphere(NodeList, 0, 1, 0, 1)
    asm = assemble_system_literal if literal else assemble_system
    K = asm(ne, NodeList, IEN, 3, "Q1", 3, ID, Young, nu)
    b = apply_boundary_conditions(ne, NodeList, IEN, IEN_top, IEN_btm, 3, "Q1", ID)
    K_bar = add_scaled(K, b, beta)
    q_d, free = setboundaryCond(NodeList, ne, 3, "Q1", d, 3)
    q = solve_reference(K_bar, q_d, free)
    return dict(NodeList=NodeList, IEN=IEN, ID=ID, IEN_top=IEN_top, IEN_btm=IEN_btm, K=K, b=b, K_bar=K_bar,
                q_d=q_d, free=free, q=q)


def plane_stress_problem(ne, d=0.001, Young=40, nu=0.4):
    """BASELINE config C1: 2-D Q4 plane stress on the unit square.  The reference has no executable 2-D
    boundary-condition / solve code (setboundaryCond indexes coord[3]; apply_boundary_conditions' 2-D branch is broken), so
    the conditions are the survey's: clamp the bottom edge y = 0 (u_x = u_y = 0), prescribe u_y = -d on the top edge
    y = 1, and solve K[free,free] q_f = -(K q_d)[free] with the idiom of examples/vector3D.jl:315-322."""
    NodeList, IEN, ID, *_ = meshgrid(0, 1, 0, 1, 0, 1, ne, 2)
    K = assemble_system_literal(ne, NodeList, IEN, 2, "Q1", 2, ID, Young, nu)
    ndof = 2 * (ne + 1) ** 2
    y = NodeList[1]
    nodes = np.arange(1, y.shape[0] + 1)
    btm, top = nodes[y == 0.0], nodes[y == 1.0]
    q_d = np.zeros((ndof, 1))
    q_d[2 * top - 1, 0] = -d                                    # u_y of the top edge (ID[m, 2] = 2 m)
    fixed = np.concatenate([2 * btm - 1, 2 * btm, 2 * top])     # 1-based dof ids: u_x, u_y bottom; u_y top
    free = np.setdiff1d(np.arange(1, ndof + 1), fixed)
    q = solve_reference(K, q_d, free)
    return dict(NodeList=NodeList, IEN=IEN, ID=ID, K=K, q_d=q_d, fixed=np.sort(fixed), free=free, q=q)


def back_project(NodeList, CameraMatrix):
    """src/PostProcess.jl:131-152: camera-frame transform, perspective divide, CameraMatrix' * p, rows 1:2."""
    R = np.array([[1.0, 0, 0], [0, 0, 1], [0, -1, 0]])  # :134
    t = np.array([[0.0], [-0.5], [2.0]])                 # :135
    NodeListTrans = R @ np.asarray(NodeList, dtype=np.float64) + t  # :137
    NodeListNorm = NodeListTrans / NodeListTrans[2:3, :]            # :141-145
    NodeListProj = np.asarray(CameraMatrix, dtype=np.float64).T @ NodeListNorm  # :147
    return NodeListProj[0:2, :]                                      # :149


def convex_hull_monotone_chain(points):
    """LazySets.convex_hull for 2-D point lists (called at src/PostProcess.jl:106).  LazySets is a third-party dependency that is
    NOT under /root/reference and is un-pinned (Project.toml lists it without [compat], no Manifest); its documented default for
    planar inputs is Andrew's monotone chain: sort lexicographically by (x, y), build the lower hull left to right and the upper
    hull right to left popping while the turn is not strictly counter-clockwise (collinear points are dropped), drop the repeated
    end points -> vertices in counter-clockwise order starting at the lexicographically smallest point.  PARITY UNPINNED."""
    pts = sorted((float(p[0]), float(p[1])) for p in points)
    if len(pts) <= 2:
        out = []
        for p in pts:
            if p not in out:
                out.append(p)
        return np.array(out, dtype=np.float64).reshape(-1, 2).T

    def right_turn(o, a, b):  # cross product (a - o) x (b - o): > 0 for a counter-clockwise turn
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    def build(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and right_turn(h[-2], h[-1], p) <= 0.0:
                h.pop()
            h.append(p)
        return h

    lower, upper = build(pts), build(reversed(pts))
    hull = lower[:-1] + upper[:-1]
    return np.array(hull, dtype=np.float64).T


def extract_borders(NodeList, CameraMatrix, BorderNodesList, state, ne=None):
    """src/PostProcess.jl:60-117.  Returns (BorderPoints 2 x nb, SideNodes2D 2 x nSide)."""
    side = np.asarray(BorderNodesList[0], dtype=np.int64) - 1      # :62 (1-based node ids)
    SideNodes2D = back_project(np.asarray(NodeList)[:, side], CameraMatrix)  # :64
    if state == "init":  # :67-100
        assert ne is not None, "Number of elements must be provided"
        Left = np.zeros((2, ne + 1))
        Right = np.zeros((2, ne + 1))
        TopLayerList, BottomLayerList = [], []
        szSide = SideNodes2D.shape[1] // (ne + 1)
        for Layers in range(1, ne + 2):
            nodes = SideNodes2D[:, (Layers - 1) * szSide:Layers * szSide]
            minNode = (Layers - 1) * szSide + int(np.argmin(nodes[0]))      # first minimum, like Julia's argmin
            maxNode = (Layers - 1) * szSide + int(np.argmax(nodes[0]))
            Left[:, Layers - 1] = SideNodes2D[:, minNode]
            Right[:, Layers - 1] = SideNodes2D[:, maxNode]
            if Layers == ne + 1:                                            # :83-88
                for nodeId in range(nodes.shape[1]):
                    if nodes[1, nodeId] > SideNodes2D[1, minNode]:
                        TopLayerList.append((Layers - 1) * szSide + nodeId)
            elif Layers == 1:                                               # :89-95 (elseif: with ne == 0 only the top list fills)
                for nodeId in range(nodes.shape[1]):
                    if nodes[1, nodeId] < SideNodes2D[1, minNode]:
                        BottomLayerList.append(nodeId)

        def sortslices(M):  # columns in lexicographic order (x, then y)
            if M.shape[1] == 0:
                return M
            return M[:, np.lexsort((M[1], M[0]))]

        Top = sortslices(SideNodes2D[:, TopLayerList])
        Bottom = sortslices(SideNodes2D[:, BottomLayerList])
        BorderPoints = np.hstack([Left, Top, Right[:, ::-1], Bottom[:, ::-1]])  # :99
    elif state == "update":  # :101-114
        BorderPoints = convex_hull_monotone_chain(SideNodes2D.T)
    else:
        raise NameError("BorderPoints not defined")  # the reference falls through to an UndefVarError
    return BorderPoints, SideNodes2D


def jitter_nodes(NodeList, ne, seed=1234, amp=0.2):
    """Robustness input: seeded jitter U(-amp*h, amp*h) of INTERIOR nodes of the unit
    cube lattice (boundary nodes fixed so the z == 0 / z == 1 Dirichlet tests still hit)."""
    rng = np.random.default_rng(seed)
    n1 = ne + 1
    h = 1.0 / ne
    disp = rng.uniform(-amp * h, amp * h, size=NodeList.shape)
    k, j, i = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
    interior = ((i > 0) & (i < ne) & (j > 0) & (j < ne) & (k > 0) & (k < ne)).ravel()
    NodeList[:, interior] += disp[:, interior]
    return NodeList
