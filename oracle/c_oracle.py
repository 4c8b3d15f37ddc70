"""ctypes front-end of oracle/c/fem_oracle.c (C form of the oracle; CPU baseline of bench.py).
TEST INFRASTRUCTURE ONLY -- see the header of oracle/c/fem_oracle.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfem_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "c", "fem_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_sparse.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
        _lib.oracle_hw_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def max_threads():
    """Threads the CPU baseline should use: every processor this process may run on.  OMP_NUM_THREADS is NOT honoured on
    purpose: torchrun exports OMP_NUM_THREADS=1 to its children, which silently made the N > 1 CPU arm single-threaded."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = int(lib().oracle_hw_threads())
    return max(1, min(n, int(lib().oracle_hw_threads())))


def gaussian_quadrature(a, b, n=2):
    xi = np.zeros(n)
    w = np.zeros(n)
    lib().oracle_gaussian_quadrature(C.c_double(a), C.c_double(b), C.c_int(n), _p(xi, C.c_double), _p(w, C.c_double))
    return xi, w


def assemble_coo(ne, NodeList, IEN, ndim, nDof, ID, Young, nu, nthreads=1, nEl=None):
    """NodeList (ndim,nNodes), IEN (nEl,nn), ID (nNodes,nDof) in Julia shapes (any memory order).
    nEl (optional): IEN holds only nEl elements (a slab of the mesh: bench.py's bounded CPU sample)."""
    NodeList = np.asfortranarray(NodeList, dtype=np.float64)
    IEN = np.asfortranarray(IEN, dtype=np.int64)
    nNodes = NodeList.shape[1]
    if ID is None:
        ID = np.zeros((1, 1), dtype=np.int64)
    ID = np.asfortranarray(ID, dtype=np.int64)
    if nEl is None:
        nEl = ne**ndim
    assert IEN.shape[0] == nEl
    L = nEl * (IEN.shape[1] * nDof) ** 2
    E = np.zeros(L, dtype=np.int64)
    J = np.zeros(L, dtype=np.int64)
    V = np.zeros(L, dtype=np.float64)
    lib().oracle_assemble_coo_n(C.c_int64(nEl), C.c_int(ndim), C.c_int(nDof), _p(NodeList, C.c_double), _p(IEN, C.c_int64),
                                _p(ID, C.c_int64), C.c_int64(nNodes), C.c_double(Young), C.c_double(nu),
                                _p(E, C.c_int64), _p(J, C.c_int64), _p(V, C.c_double), C.c_int(nthreads))
    return E, J, V


def sparse(E, J, V):
    from .fem_oracle import JuliaCSC

    L = E.shape[0]
    n_guess = int(J.max())
    colptr = np.zeros(n_guess + 1, dtype=np.int64)
    rowval = np.zeros(L, dtype=np.int64)
    nzval = np.zeros(L, dtype=np.float64)
    m = C.c_int64()
    n = C.c_int64()
    nnz = lib().oracle_sparse(C.c_int64(L), _p(E, C.c_int64), _p(J, C.c_int64), _p(V, C.c_double), C.byref(m), C.byref(n),
                              _p(colptr, C.c_int64), _p(rowval, C.c_int64), _p(nzval, C.c_double))
    return JuliaCSC(m.value, n.value, colptr, rowval[:nnz].copy(), nzval[:nnz].copy())


def assemble_system(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3, nthreads=1, nEl=None):
    assert FunctionClass == "Q1"
    return sparse(*assemble_coo(ne, NodeList, IEN, ndim, nDof, ID, Young, nu, nthreads, nEl))
