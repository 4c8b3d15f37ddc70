"""Opt-in experimental kernels (selected by environment variables, never the default path).  The file name sorts last on purpose:
these variants are recorded experiments (profiles/r2_tile2_variants.txt), and under `pytest -x` a problem in one of them must not
cut short the parity tests of the kernels the library actually runs by default.

Hardware record: the diagonal of K from v3 / v3b was bit-identical to the default kernel at 100^3 (tools/time_tile2.py, profiles/
r2_tile2_variants.txt); this test is the stricter check (colptr / rowval / nzval / diagonal, every buffer cleared to 0xFF first).
No log of THIS test on hardware exists (the round's GPU budget ended first), so it runs only on request
(SMFEM_TEST_EXPERIMENTAL=1): the default `-m gpu` suite is exactly the set of tests that has been seen green on a B200."""
import os

import numpy as np
import pytest

import smearfem_b200 as sf

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("SMFEM_TEST_EXPERIMENTAL") != "1",
                       reason="opt-in experimental kernels: set SMFEM_TEST_EXPERIMENTAL=1 (no hardware log of this test yet)"),
]


@pytest.mark.parametrize("ne", [5, 22, 37])
@pytest.mark.parametrize("variant", ["v3", "v3b"])
def test_split_role_layer_march_kernels_same_bits(monkeypatch, variant, ne):
    """assemble_tile3.cu (SMFEM_TILE=v3: 4 x 8 tile, 256 threads, half-size staging; v3b: 4 x 6 tile, 192 threads, full staging)
    does the arithmetic of k_values_tile2 in the same order with the accumulators split between two threads: same bits.
    ne = 5: one partial tile; 22: partial tiles in x and y with v3b's 6-row tiles; 37: several tiles, interior fast path."""
    ctx = sf.context()
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    monkeypatch.setenv("SMFEM_TILE", "v2")
    K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4)
    ref = K.to_csc() + (K.diag(),)
    monkeypatch.setenv("SMFEM_DEBUG_CLEAR", "1")
    monkeypatch.setenv("SMFEM_TILE", variant)
    K.reassemble(40, 0.4)
    got = K.to_csc() + (K.diag(),)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b), variant
    K.free()
    mesh.free()
