"""TEST INFRASTRUCTURE ONLY (like the rest of oracle/): NumPy/SciPy statement of the multigrid-preconditioned CG that
smearfem.jl_b200/csrc/gmg.cu implements on the device, built from the oracle's restatement of the reference's assembly
(src/fem.jl:135-256, examples/vector3D.jl:175-264).  The reference itself has no iterative solver (dense inverse,
examples/vector3D.jl:315-322); this file pins what the opt-in preconditioner is and how fast it must converge.

  levels      ne -> ceil(ne/2) -> ... <= 4; coarse node I sits on fine node min(2 I, ne_f)
  operators   re-assembled on the subsampled (inflated) nodes with the same E, nu, beta; Dirichlet dofs by injection
  smoother    Chebyshev in D^-1 A on [lmax/8, lmax], 2 steps before / after the coarse correction, 30 on the coarsest level
  transfers   trilinear interpolation per component, restriction = transpose
"""
import numpy as np
import scipy.sparse as sp

from . import fem_oracle as o


def fine_of(I, ne_f):
    return np.minimum(2 * np.asarray(I), ne_f)


def prolong_1d(ne_f):
    """(ne_f + 1) x (ne_c + 1) linear interpolation; for odd ne_f the last fine node is a coarse node itself."""
    ne_c = (ne_f + 1) // 2
    P = sp.lil_matrix((ne_f + 1, ne_c + 1))
    for i in range(ne_f + 1):
        if i == ne_f and ne_f % 2 == 1:
            P[i, ne_c] = 1.0
        elif i % 2 == 0:
            P[i, i // 2] = 1.0
        else:
            P[i, i // 2] = 0.5
            P[i, i // 2 + 1] = 0.5
    return P.tocsr()


def level_operator(NL_level, ne, Young, nu, beta):
    _, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    K = o.assemble_system(ne, NL_level, IEN, 3, "Q1", 3, ID, Young, nu)
    if beta:
        K = o.add_scaled(K, o.apply_boundary_conditions(ne, NL_level, IEN, top, btm, 3, "Q1", ID), beta)
    return K.to_scipy().tocsr()


def build(ne, NL, fixed, Young=40, nu=0.4, beta=100.0):
    levels = []
    ne_l, NL_l, fixed_l = ne, NL, fixed
    while True:
        A = level_operator(NL_l, ne_l, Young, nu, beta)
        # lmax(D^-1 A) of the unconstrained operator: power iteration, 10 % margin (bounds every Dirichlet set)
        dfull = 1.0 / A.diagonal()
        x = 1.0 + 0.37 * ((np.arange(A.shape[0]) * 2654435761) % 1000) / 1000.0
        lam = 1.0
        for _ in range(15):
            x = dfull * (A @ x)
            lam = np.linalg.norm(x)
            x /= lam
        levels.append(dict(ne=ne_l, A=A, fixed=fixed_l, dinv=np.where(fixed_l, 0.0, dfull), lmax=1.1 * lam))
        if ne_l <= 4 or len(levels) >= 12:
            break
        ne_c = (ne_l + 1) // 2
        n1f, n1c = ne_l + 1, ne_c + 1
        f = fine_of(np.arange(n1c), ne_l)
        Kc, Jc, Ic = np.meshgrid(f, f, f, indexing="ij")
        ids = ((Kc * n1f + Jc) * n1f + Ic).ravel()
        P1 = prolong_1d(ne_l)
        levels[-1]["P"] = sp.kron(sp.kron(sp.kron(P1, P1), P1), sp.identity(3)).tocsr()
        NL_l = NL_l[:, ids]
        fixed_l = fixed_l.reshape(-1, 3)[ids].ravel()
        ne_l = ne_c
    return levels


def chebyshev(L, b, x, n, x_zero):
    A, dinv, lmax = L["A"], L["dinv"], L["lmax"]
    lmin = lmax / 8.0
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma = theta / delta
    rho = 1.0 / sigma
    d = None
    for k in range(n):
        r = dinv * b if (k == 0 and x_zero) else dinv * (b - A @ x)
        if k == 0:
            d = r / theta
        else:
            rho_n = 1.0 / (2.0 * sigma - rho)
            d = rho_n * rho * d + 2.0 * rho_n / delta * r
            rho = rho_n
        x = d.copy() if (k == 0 and x_zero) else x + d
    return x


def vcycle(levels, l, b):
    L = levels[l]
    if l == len(levels) - 1:
        return chebyshev(L, b, np.zeros_like(b), 2 if len(levels) == 1 else 30, True)
    x = chebyshev(L, b, np.zeros_like(b), 2, True)
    r = np.where(L["fixed"], 0.0, b - L["A"] @ x)
    rc = L["P"].T @ r
    rc[levels[l + 1]["fixed"]] = 0.0
    e = L["P"] @ vcycle(levels, l + 1, rc)
    e[L["fixed"]] = 0.0
    return chebyshev(L, b, x + e, 2, False)


def pcg(levels, q_d, rtol=1e-10, maxit=500, multigrid=True):
    L0 = levels[0]
    A, fixed = L0["A"], L0["fixed"]
    b = np.where(fixed, 0.0, -(A @ q_d))
    M = (lambda r: vcycle(levels, 0, r)) if multigrid else (lambda r: L0["dinv"] * r)
    x = np.zeros_like(b)
    r = b.copy()
    bn = np.linalg.norm(b)
    rz_old, p = 0.0, None
    for it in range(1, maxit + 1):
        z = M(r)
        rz = r @ z
        p = z if it == 1 else z + (rz / rz_old) * p
        rz_old = rz
        Ap = np.where(fixed, 0.0, A @ p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= rtol * bn:
            return q_d + x, it
    return q_d + x, maxit


def example(ne, d=0.001, **kw):
    """The example problem (examples/vector3D.jl) solved with both preconditioners: (q_mg, it_mg, q_jacobi, it_jacobi)."""
    NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    n1 = ne + 1
    kz = np.arange(n1**3) // (n1 * n1)
    fixed = np.zeros(3 * n1**3, bool)
    fixed[3 * np.where((kz == 0) | (kz == ne))[0] + 2] = True
    q_d = np.zeros(3 * n1**3)
    q_d[3 * np.where(kz == ne)[0] + 2] = -d
    levels = build(ne, NL, fixed)
    qm, itm = pcg(levels, q_d, multigrid=True, **kw)
    qj, itj = pcg(levels, q_d, multigrid=False, maxit=5000, **kw)
    return qm, itm, qj, itj, [L["ne"] for L in levels]
