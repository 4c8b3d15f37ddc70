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
