"""The reference's own unit tests (test/runtests.jl:15-35), restated against the oracle.
These are the ONLY results the reference pins; all are compared bit-exactly (`==`), as upstream."""
import math

import numpy as np
import pytest

from oracle import c_oracle, fem_oracle as o


def test_basis_1d_corners():  # test/runtests.jl:15-16
    assert o.basis_function(-1)[0].tolist() == [1.0, 0.0]
    assert o.basis_function(1)[0].tolist() == [0.0, 1.0]


@pytest.mark.parametrize("pt,idx", [((-1, -1), 0), ((1, -1), 1), ((1, 1), 2), ((-1, 1), 3)])
def test_basis_2d_corners(pt, idx):  # test/runtests.jl:18-21
    e = [0.0] * 4
    e[idx] = 1.0
    assert o.basis_function(*pt)[0].tolist() == e


@pytest.mark.parametrize("pt,idx", [((-1, -1, -1), 0), ((1, -1, -1), 1), ((1, 1, -1), 2), ((-1, 1, -1), 3),
                                    ((-1, -1, 1), 4), ((1, -1, 1), 5), ((1, 1, 1), 6), ((-1, 1, 1), 7)])
def test_basis_3d_corners(pt, idx):  # test/runtests.jl:23-30
    e = [0.0] * 8
    e[idx] = 1.0
    assert o.basis_function(*pt)[0].tolist() == e


def test_gauss_2pt():  # test/runtests.jl:34
    xi, w = o.gaussian_quadrature(-1, 1, 2)
    assert xi.tolist() == [-1 / math.sqrt(3), 1 / math.sqrt(3)]
    assert w.tolist() == [1.0, 1.0]


def test_gauss_3pt():  # test/runtests.jl:35
    xi, w = o.gaussian_quadrature(-1, 1, 3)
    assert xi.tolist() == [-math.sqrt(3 / 5), 0.0, math.sqrt(3 / 5)]
    assert w.tolist() == [5 / 9, 8 / 9, 5 / 9]


def test_gauss_c_form_bit_exact(golden_dir):
    g = np.load(golden_dir + "/gauss.npz")
    for n in (2, 3):
        xi, w = c_oracle.gaussian_quadrature(-1, 1, n)
        assert np.array_equal(xi, g[f"gq{n}_xi"]) and np.array_equal(w, g[f"gq{n}_w"])
        xi, w = o.gaussian_quadrature(-1, 1, n)
        assert np.array_equal(xi, g[f"gq{n}_xi"]) and np.array_equal(w, g[f"gq{n}_w"])


def test_gauss_other_n_is_an_error():  # src/fem.jl:23-30: xi undefined for any other n
    with pytest.raises(ValueError):
        o.gaussian_quadrature(-1, 1, 4)


def test_basis_gradients_sum_to_zero_and_partition_of_unity():
    for pt in [(0.3, -0.2, 0.7), (0.1, 0.9), ]:
        N, dN = o.basis_function(*pt)
        assert abs(N.sum() - 1) < 1e-15 and np.abs(dN.sum(axis=0)).max() < 1e-15
    N, dN = o.basis_function(0.3, -0.4, None, "Q2")
    assert N.shape == (9,) and dN.shape == (9, 2)
    assert abs(N.sum() - 1) < 1e-14 and np.abs(dN.sum(axis=0)).max() < 1e-14
    # 1-D quirk: dN is the 1x2 row matrix of src/fem.jl:75
    assert o.basis_function(0.0)[1].shape == (1, 2)
