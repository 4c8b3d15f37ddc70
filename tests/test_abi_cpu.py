"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol
include/smearfem_b200.h declares, its host helpers reproduce the reference's own unit tests
(test/runtests.jl:15-35) bit-for-bit, and device entry points fail loudly without a device."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import smearfem_b200 as sf
from smearfem_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "smearfem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smfem_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/smearfem_b200.h but not exported"
    # and the Python binding table covers them all
    bound = set(_lib.SIGNATURES) | set(_lib.NON_STATUS)
    assert set(names) == bound, set(names) ^ bound
    assert L.smfem_abi_version() == 2


def test_no_oracle_import_in_product():
    """The product must never import, link or execute the test oracle."""
    pkg = os.path.join(ROOT, "smearfem.jl_b200")
    bad = re.compile(r"import\s+oracle|from\s+oracle|fem_oracle|c_oracle|oracle/|libfem_oracle")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl", ".cpp", ".h")) or f == "Makefile":
                assert not bad.search(open(os.path.join(dp, f)).read()), os.path.join(dp, f)
    assert not bad.search(open(os.path.join(ROOT, "smearfem_b200.py")).read())


# ---- the reference's unit tests through the ABI (host helpers; no GPU needed) ---------------------
def test_basis_1d():  # test/runtests.jl:15-16
    assert sf.basis_function(-1)[0].tolist() == [1.0, 0.0]
    assert sf.basis_function(1)[0].tolist() == [0.0, 1.0]
    assert sf.basis_function(0.3)[1].shape == (1, 2)  # src/fem.jl:75 quirk


@pytest.mark.parametrize("pt,idx", [((-1, -1), 0), ((1, -1), 1), ((1, 1), 2), ((-1, 1), 3)])
def test_basis_2d(pt, idx):  # test/runtests.jl:18-21
    e = [0.0] * 4
    e[idx] = 1.0
    assert sf.basis_function(*pt)[0].tolist() == e


@pytest.mark.parametrize("pt,idx", [((-1, -1, -1), 0), ((1, -1, -1), 1), ((1, 1, -1), 2), ((-1, 1, -1), 3),
                                    ((-1, -1, 1), 4), ((1, -1, 1), 5), ((1, 1, 1), 6), ((-1, 1, 1), 7)])
def test_basis_3d(pt, idx):  # test/runtests.jl:23-30
    e = [0.0] * 8
    e[idx] = 1.0
    assert sf.basis_function(*pt)[0].tolist() == e


def test_gauss():  # test/runtests.jl:34-35
    xi, w = sf.gaussian_quadrature(-1, 1, 2)
    assert xi.tolist() == [-1 / math.sqrt(3), 1 / math.sqrt(3)] and w.tolist() == [1.0, 1.0]
    xi, w = sf.gaussian_quadrature(-1, 1, 3)
    assert xi.tolist() == [-math.sqrt(3 / 5), 0.0, math.sqrt(3 / 5)] and w.tolist() == [5 / 9, 8 / 9, 5 / 9]
    with pytest.raises(sf.SmearFEMError):
        sf.gaussian_quadrature(-1, 1, 4)


def test_host_helpers_match_oracle_bitwise():
    from oracle import fem_oracle as o

    rng = np.random.default_rng(0)
    for _ in range(20):
        x, e, z = rng.uniform(-1, 1, 3)
        for args in [(x,), (x, e), (x, e, z)]:
            N, dN = sf.basis_function(*args)
            No, dNo = o.basis_function(*args)
            assert np.array_equal(N, No) and np.array_equal(dN, dNo)
        N, dN = sf.basis_function(x, e, None, "Q2")
        No, dNo = o.basis_function(x, e, None, "Q2")
        assert np.array_equal(N, No) and np.array_equal(dN, dNo)
    a, b = rng.uniform(-2, 2, 2)
    for n in (2, 3):
        assert all(np.array_equal(u, v) for u, v in zip(sf.gaussian_quadrature(a, b, n), o.gaussian_quadrature(a, b, n)))
    with pytest.raises(sf.SmearFEMError):
        sf.basis_function(0.1, 0.2, 0.3, "Q2")  # reference defines Q2 in 2-D only
    with pytest.raises(sf.SmearFEMError):
        sf.basis_function(0.1, 0.2, 0.3, "Q7")


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sf.SmearFEMError) as ei:
        sf.Context(device=0, rank=0, nranks=1)
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_multi_gpu_entry_point_fails_loudly_without_gpu():
    """smfem_init_multi (one process, n GPUs) has no CPU fallback either, and rejects impossible rank counts before touching CUDA."""
    import torch

    with pytest.raises(sf.SmearFEMError) as ei:
        sf.MultiContext(0)
    assert ei.value.code == _lib.ERR_INVALID
    with pytest.raises(sf.SmearFEMError):
        sf.MultiContext(9)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sf.SmearFEMError) as ei:
        sf.MultiContext(2)
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_setboundarycond_host_prep_matches_oracle():
    from oracle import fem_oracle as o

    NL, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, 3, 3)
    q_d, Cc = sf.setboundaryCond(NL, 3, 3, "Q1", 0.25, 3)
    q_o, free_o = o.setboundaryCond(NL, 3, 3, "Q1", 0.25, 3)
    assert np.array_equal(q_d, q_o) and np.array_equal(Cc.free, free_o) and Cc.shape == (192, 160)


def test_meshgrid_2d_host_matches_oracle():
    from oracle import fem_oracle as o

    for ne in (1, 2, 5):
        a = sf.meshgrid(0, 1, 0, 1, 0, 1, ne, 2)
        b = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 2)
        for u, v in zip(a[:5], b[:5]):
            assert np.array_equal(u, v)
        assert a[5][0] == b[5][0]


def _header_param_counts():
    src = open(os.path.join(ROOT, "include", "smearfem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(smfem_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_julia_shim_ccalls_match_the_header():
    """Julia is not installed, so the shim cannot be executed here; this is the static check that stands in for it: every
    `ccall((:smfem_x, LIB), ret, (argument types...), ...)` names a function the header declares and passes exactly as many
    argument types as the C prototype has parameters."""
    shim = open(os.path.join(ROOT, "smearfem.jl_b200", "julia", "SmearFEMB200.jl")).read()
    counts = _header_param_counts()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(smfem_[a-z0-9_]+),\s*LIB\),\s*(\w+),\s*\(", shim):
        name = m.group(1)
        assert name in counts, f"{name} is not declared in include/smearfem_b200.h"
        # the type tuple: balanced parentheses starting at the '(' the regex ended on
        i = m.end() - 1
        depth, j = 0, i
        while True:
            depth += shim[j] == "("
            depth -= shim[j] == ")"
            if depth == 0:
                break
            j += 1
        tup = shim[i + 1:j]
        # split on top-level commas (Ptr{Ptr{Cvoid}} contains braces, not commas; a trailing comma makes a 1-tuple)
        parts, depth_b, cur = [], 0, ""
        for ch in tup:
            depth_b += ch in "{("
            depth_b -= ch in "})"
            if ch == "," and depth_b == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        ntypes = len([p for p in parts if p.strip()])
        assert ntypes == counts[name], f"{name}: the shim passes {ntypes} argument types, the header declares {counts[name]} parameters"
        assert (m.group(2) == "Cstring") == (name == "smfem_last_error")
        seen += 1
    assert seen >= 25


def _header_param_types():
    """name -> list of canonical parameter types of the C prototype: 'int', 'i64', 'f64', 'f32*', 'i64*', 'int*', 'f64*', 'ptr'
    (an opaque handle or void*), 'ptr*' (pointer to a handle / to void*)."""
    src = open(os.path.join(ROOT, "include", "smearfem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(smfem_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        types = []
        if args not in ("", "void"):
            for a in args.split(","):
                a = re.sub(r"\bconst\b", "", a).strip()
                stars = a.count("*")
                base = re.match(r"\s*([A-Za-z_0-9]+)", a).group(1)
                scalar = {"int": "int", "int64_t": "i64", "double": "f64", "float": "f32"}.get(base)
                if scalar is None:  # smfem_ctx / smfem_mesh / smfem_matrix / smfem_multi* / void: opaque
                    assert stars >= 1, (m.group(1), a)
                    types.append("ptr" + "*" * (stars - 1))
                else:
                    types.append(scalar + "*" * stars)
        out[m.group(1)] = types
    return out


def test_python_binding_types_match_the_header():
    """Every ctypes argtypes list in _lib.SIGNATURES has the C prototype's parameter types, position by position (a c_int where
    the header says int64_t would pass on x86-64 registers and fail on the stack)."""
    import ctypes as C

    def canon(t):
        if t is C.c_int:
            return "int"
        if t is C.c_int64:
            return "i64"
        if t is C.c_double:
            return "f64"
        if t is C.c_void_p:
            return "ptr"
        if t is C.c_char_p:
            return "ptr"
        if t is C.c_float:
            return "f32"
        if isinstance(t, type) and issubclass(t, C._Pointer):  # POINTER(x)
            return canon(t._type_) + "*"
        raise AssertionError(t)

    hdr = _header_param_types()
    for name, args in _lib.SIGNATURES.items():
        got = [canon(t) for t in args]
        assert got == hdr[name], f"{name}: ctypes {got} vs header {hdr[name]}"
    for name, (_, args) in _lib.NON_STATUS.items():
        assert [canon(t) for t in args] == hdr[name]
    assert set(hdr) == set(_lib.SIGNATURES) | set(_lib.NON_STATUS)


def _split_top_level(tup):
    parts, depth, cur = [], 0, ""
    for ch in tup:
        depth += ch in "{("
        depth -= ch in "})"
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


def test_julia_shim_ccall_types_match_the_header():
    """Type-level form of the static shim check: each entry of a ccall's argument-type tuple is the Julia spelling of the C
    parameter type at that position (Cint / Int64 / Cdouble / Ptr{...} / Ref{...}; handles are Ptr{Cvoid})."""
    julia = {
        "int": {"Cint"}, "i64": {"Int64", "Clonglong"}, "f64": {"Cdouble", "Float64"}, "f32": {"Cfloat"},
        "int*": {"Ptr{Cint}", "Ref{Cint}"}, "i64*": {"Ptr{Int64}", "Ref{Int64}"}, "f64*": {"Ptr{Cdouble}", "Ref{Cdouble}", "Ptr{Float64}"},
        "f32*": {"Ptr{Cfloat}", "Ref{Cfloat}"}, "ptr": {"Ptr{Cvoid}"}, "ptr*": {"Ptr{Ptr{Cvoid}}", "Ref{Ptr{Cvoid}}"},
    }
    shim = open(os.path.join(ROOT, "smearfem.jl_b200", "julia", "SmearFEMB200.jl")).read()
    hdr = _header_param_types()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(smfem_[a-z0-9_]+),\s*LIB\),\s*(\w+),\s*\(", shim):
        name = m.group(1)
        i = m.end() - 1
        depth, j = 0, i
        while True:
            depth += shim[j] == "("
            depth -= shim[j] == ")"
            if depth == 0:
                break
            j += 1
        got = _split_top_level(shim[i + 1:j])
        want = hdr[name]
        assert len(got) == len(want), name
        for k, (g, w) in enumerate(zip(got, want)):
            ok = g in julia[w] or (w.endswith("*") and g == "Ptr{Cvoid}")  # a raw buffer may be passed as an untyped pointer
            assert ok, f"{name}: argument {k + 1} is {g} in the shim, the header says {w}"
        # the values: ccall is a special form, so exactly one value per type must be spelled out (no splatting)
        depth, e = 1, j + 1
        while depth:
            depth += shim[e] in "([{"
            depth -= shim[e] in ")]}"
            e += 1
        values = _split_top_level(shim[j + 1:e - 1].replace("[", "(").replace("]", ")"))
        assert len(values) == len(want), f"{name}: {len(values)} values for {len(want)} parameters"
        assert not any(v.endswith("...") for v in values), f"{name}: splatted ccall argument"
        seen += 1
    assert seen >= 25


def test_every_entry_point_rejects_null_handles_without_crashing():
    """Error behaviour at the boundary (the reference throws Julia exceptions; no exception, abort or segfault may cross the C ABI):
    every exported function called with NULL pointers and zero scalars returns a status (free / destroy of NULL are no-ops that
    return SMFEM_OK, everything else SMFEM_ERR_INVALID with a message).  Runs in a child process so that a crash is a test
    failure, not the end of the test session.  No GPU is touched."""
    import subprocess
    import sys

    code = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
from smearfem_b200 import _lib
L = _lib.lib()
noop_ok = {"smfem_destroy", "smfem_mesh_free", "smfem_matrix_free", "smfem_multi_destroy", "smfem_multi_matrix_free", "smfem_multi_mesh_free"}
for name, types in _lib.SIGNATURES.items():
    args = [0 if t in (C.c_int, C.c_int64) else 0.0 if t is C.c_double else None for t in types]
    rc = getattr(L, name)(*args)
    want = _lib.OK if name in noop_ok else _lib.ERR_INVALID
    assert rc == want, (name, rc, L.smfem_last_error())
    assert name in noop_ok or L.smfem_last_error(), name
print("ok", len(_lib.SIGNATURES))
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().startswith("ok")
