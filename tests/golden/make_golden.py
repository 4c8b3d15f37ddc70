"""Generates tests/golden/*.npz from the LITERAL loop form of the oracle (oracle/fem_oracle.py).

The reference is Julia and cannot be executed in this image (no julia binary, no network), so these
are NOT outputs of the reference itself: they freeze the oracle's literal restatement so that the
vectorised / C forms, and the CUDA path on the GPU box, are all compared with the same numbers.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fem_oracle as o  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def csc(prefix, K):
    return {prefix + "_m": K.m, prefix + "_n": K.n, prefix + "_colptr": K.colptr, prefix + "_rowval": K.rowval,
            prefix + "_nzval": K.nzval}


def main():
    # reference unit tests, test/runtests.jl:15-35
    d = {}
    d["gq2_xi"], d["gq2_w"] = o.gaussian_quadrature(-1, 1, 2)
    d["gq3_xi"], d["gq3_w"] = o.gaussian_quadrature(-1, 1, 3)
    np.savez_compressed(os.path.join(OUT, "gauss.npz"), **d)

    # hex elasticity, cube and inflated, ne = 2 and 4  (E=40, nu=0.4 as examples/vector3D.jl:275-276)
    for ne in (2, 4):
        for inflate in (False, True):
            NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
            if inflate:
                o.inflate_sphere(NL, 0, 1, 0, 1)
            K = o.assemble_system_literal(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4)
            b = o.apply_boundary_conditions(ne, NL, IEN, top, btm, 3, "Q1", ID)
            Kb = o.add_scaled(K, b, 100)
            q_d, free = o.setboundaryCond(NL, ne, 3, "Q1", 0.001, 3)
            q = o.solve_reference(Kb, q_d, free, dense=True)
            d = dict(ne=ne, NodeList=NL, IEN=IEN, ID=ID, IEN_top=top, IEN_btm=btm, q_d=q_d, free=free, q=q)
            d.update(csc("K", K))
            d.update(csc("b", b))
            np.savez_compressed(os.path.join(OUT, f"hex_ne{ne}_{'inflated' if inflate else 'cube'}.npz"), **d)

    # scalar Laplace 3-D ne=2 and 2-D plane stress / scalar ne=2,4
    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 2, 3)
    K = o.assemble_system_literal(2, NL, IEN, 3, "Q1", 1)
    np.savez_compressed(os.path.join(OUT, "scalar3d_ne2.npz"), NodeList=NL, IEN=IEN, **csc("K", K))
    for ne in (2, 4):
        NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 2)
        K = o.assemble_system_literal(ne, NL, IEN, 2, "Q1", 2, ID, 40, 0.4)
        Ks = o.assemble_system_literal(ne, NL, IEN, 2, "Q1", 1)
        np.savez_compressed(os.path.join(OUT, f"quad_ne{ne}.npz"), NodeList=NL, IEN=IEN, ID=ID, **csc("K", K), **csc("Ks", Ks))

    # config C1: 2-D plane stress solve, ne = 8 (81 nodes, 162 dofs, nnz 2500) and 16
    for ne in (8, 16):
        r = o.plane_stress_problem(ne)
        np.savez_compressed(os.path.join(OUT, f"c1_plane_stress_ne{ne}.npz"), ne=ne, NodeList=r["NodeList"], IEN=r["IEN"], ID=r["ID"],
                            q_d=r["q_d"], fixed=r["fixed"], free=r["free"], q=r["q"], **csc("K", r["K"]))

    # example problem summaries at ne = 8 and 20 (vectorised form; norms only + q at ne=8)
    summ = {}
    for ne in (8, 20):
        r = o.example_problem(ne)
        K, b = r["K"], r["b"]
        summ[f"ne{ne}"] = np.array([K.nnz, np.linalg.norm(K.nzval), K.to_scipy().diagonal().sum(), np.linalg.norm(b.nzval),
                                    b.nzval.sum(), len(r["free"]), np.linalg.norm(r["q"]), np.abs(r["q"][0::3]).max()])
        if ne == 8:
            summ["q_ne8"] = r["q"]
        else:
            summ["q_ne20"] = r["q"]
    np.savez_compressed(os.path.join(OUT, "example_summaries.npz"), **summ)


if __name__ == "__main__":
    main()
