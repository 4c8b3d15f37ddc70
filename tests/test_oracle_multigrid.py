"""CPU statement of the opt-in multigrid-preconditioned CG (oracle/gmg_prototype.py): it converges to the oracle's direct
solution in a mesh-independent number of iterations, for even and odd ne (the device implementation is tested against the
same oracle in tests/test_gpu_parity.py::test_multigrid_pcg_matches_oracle)."""
import numpy as np
import pytest

from oracle import fem_oracle as o
from oracle import gmg_prototype as g


def test_prolongation_rows_sum_to_one_and_hit_coarse_nodes():
    for ne_f in (4, 5, 7, 8, 13):
        P = g.prolong_1d(ne_f).toarray()
        ne_c = (ne_f + 1) // 2
        assert P.shape == (ne_f + 1, ne_c + 1) and np.allclose(P.sum(axis=1), 1.0)
        f = g.fine_of(np.arange(ne_c + 1), ne_f)
        assert np.array_equal(P[f, np.arange(ne_c + 1)], np.ones(ne_c + 1))  # coarse nodes are injected
        assert f[0] == 0 and f[-1] == ne_f                                   # the z = 0 / z = 1 planes stay planes


@pytest.mark.parametrize("ne", [8, 11])
def test_multigrid_pcg_prototype_converges_fast(ne):
    qm, itm, qj, itj, sizes = g.example(ne, rtol=1e-11)
    ref = o.example_problem(ne)["q"]
    assert sizes[-1] <= 4 and len(sizes) >= 2
    assert np.linalg.norm(qm - ref) <= 1e-9 * np.linalg.norm(ref)
    assert np.linalg.norm(qj - ref) <= 1e-9 * np.linalg.norm(ref)
    assert itm <= 25 and itm < itj / 3, (itm, itj)


def _zplane_problem(ne, NL, clamp_bottom=False):
    n1 = ne + 1
    kz = np.arange(n1**3) // (n1 * n1)
    fixed = np.zeros(3 * n1**3, bool)
    fixed[3 * np.where((kz == 0) | (kz == ne))[0] + 2] = True
    if clamp_bottom:
        b = np.where(kz == 0)[0]
        fixed[3 * b] = fixed[3 * b + 1] = True
    q_d = np.zeros(3 * n1**3)
    q_d[3 * np.where(kz == ne)[0] + 2] = -0.001
    return fixed, q_d


def test_multigrid_prototype_is_robust_to_mesh_and_boundary_conditions():
    """Jittered (non-lattice-like) geometry, and a clamped bottom face without the surface term: still ~15 iterations."""
    ne = 8
    NL, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    o.jitter_nodes(NL, ne, seed=7, amp=0.3)
    fixed, q_d = _zplane_problem(ne, NL)
    lev = g.build(ne, NL, fixed)
    qm, itm = g.pcg(lev, q_d, rtol=1e-10)
    qj, itj = g.pcg(lev, q_d, rtol=1e-10, multigrid=False, maxit=5000)
    assert itm <= 25 and itm < itj / 3 and np.linalg.norm(qm - qj) <= 1e-8 * np.linalg.norm(qj)
    NL, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
    fixed, q_d = _zplane_problem(ne, NL, clamp_bottom=True)
    lev = g.build(ne, NL, fixed, beta=0.0)
    qm, itm = g.pcg(lev, q_d, rtol=1e-10)
    qj, itj = g.pcg(lev, q_d, rtol=1e-10, multigrid=False, maxit=5000)
    assert itm <= 25 and itm < itj / 3 and np.linalg.norm(qm - qj) <= 1e-8 * np.linalg.norm(qj)
