"""ctypes loader for libsmearfem_b200.so (the C ABI declared in include/smearfem_b200.h).

There is NO CPU fallback: if the shared library is missing the import fails loudly, and if there is
no CUDA device `smfem_init` fails with SMFEM_ERR_CUDA (surfaced as SmearFEMError)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMEARFEM_B200_LIB") or os.path.join(_HERE, "libsmearfem_b200.so")  # same override as the Julia shim

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_SINGULAR = 0, 1, 2, 3, 4
Q1, Q2 = 1, 2
IPC_HANDLE_BYTES = 64


class SmearFEMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[smfem status {code}] {msg}")
        self.code = code


_lib = None

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> argtypes; every function returns int except the two noted below.  This table IS the
# binding a Julia maintainer would write with `ccall`.
SIGNATURES = {
    "smfem_gaussian_quadrature": [C.c_double, C.c_double, C.c_int, _f64p, _f64p],
    "smfem_basis_function": [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _f64p, _f64p, C.POINTER(C.c_int)],
    "smfem_init": [C.c_int, C.c_int, C.c_int, C.POINTER(_vp)],
    "smfem_destroy": [_vp],
    "smfem_stream": [_vp, C.POINTER(_vp)],
    "smfem_timer_start": [_vp],
    "smfem_timer_stop": [_vp, C.POINTER(C.c_float)],
    "smfem_sync": [_vp],
    "smfem_launch_count": [_vp, _i64p],
    "smfem_flush_l2": [_vp],
    "smfem_meshgrid": [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int,
                       C.POINTER(_vp)],
    "smfem_mesh_from_host": [_vp, _f64p, _i64p, _i64p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64,
                             C.POINTER(_vp)],
    "smfem_inflate_sphere": [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double],
    "smfem_inflate_sphere_host": [_vp, _f64p, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double],
    "smfem_mesh_set_nodelist": [_vp, _vp, _f64p],
    "smfem_mesh_info": [_vp, _i64p, _i64p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _i64p, _i64p],
    "smfem_mesh_export": [_vp, _vp, _f64p, _i64p, _i64p, _i64p, _i64p],
    "smfem_mesh_free": [_vp],
    "smfem_assemble": [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(_vp)],
    "smfem_assemble_system": [_vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int64, C.c_int64, C.c_int,
                              C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(_vp), C.POINTER(_vp)],
    "smfem_assembly_kernel_ms": [_vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)],
    "smfem_transfer_bytes": [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    "smfem_pattern_build": [_vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)],
    "smfem_assemble_values": [_vp, _vp, _vp, C.c_double, C.c_double],
    "smfem_pattern_rebuild": [_vp, _vp, _vp],
    "smfem_reassemble": [_vp, _vp, _vp, C.c_double, C.c_double],
    "smfem_matrix_info": [_vp, _i64p, _i64p, _i64p, _i64p, _i64p, _i64p],
    "smfem_matrix_export_csc": [_vp, _vp, C.c_int, _i64p, _i64p, _f64p],
    "smfem_matrix_diag": [_vp, _vp, _f64p],
    "smfem_mesh_colors": [_vp, _vp, C.POINTER(C.c_int), _i64p],
    "smfem_pcg_use_multigrid": [_vp, _vp, _vp, C.c_int],
    "smfem_pcg_use_matrix_free": [_vp, _vp, _vp, C.c_int],
    "smfem_matfree_operator": [_vp, _vp, C.c_double, C.c_double, C.POINTER(_vp)],
    "smfem_pcg_apply_preconditioner": [_vp, _vp, _f64p, _f64p],
    "smfem_project_nodes": [_vp, _vp, _vp, _i64p, C.c_int64, _f64p, _f64p, _f64p],
    "smfem_extract_borders": [_vp, _vp, _vp, _i64p, C.c_int64, _f64p, C.c_int, C.c_int64, _f64p, C.c_int64, _i64p, _f64p],
    "smfem_matrix_free": [_vp],
    "smfem_matrix_clone": [_vp, _vp, C.POINTER(_vp)],
    "smfem_surface_mass": [_vp, _vp, _vp, _i64p, _i64p, C.c_int64, C.c_double, C.c_int],
    "smfem_set_dirichlet_zplanes": [_vp, _vp, _vp, C.c_double],
    "smfem_set_dirichlet": [_vp, _vp, _i64p, _f64p, C.c_int64],
    "smfem_pcg_solve": [_vp, _vp, C.c_double, C.c_int, _f64p, _f64p, C.POINTER(C.c_int), _f64p],
    "smfem_pcg_set_warm_start": [_vp, C.c_double],
    "smfem_spmv_host": [_vp, _vp, _f64p, _f64p],
    "smfem_bench_spmv": [_vp, _vp, C.c_int, C.c_int, C.POINTER(C.c_float)],
    "smfem_set_spmv_variant": [_vp, C.c_int],
    "smfem_pcg_stats": [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)],
    "smfem_pcg_wait_stats": [_vp, _f64p, _f64p, _f64p],
    "smfem_comm_export": [_vp, _vp, _vp],
    "smfem_comm_connect": [_vp, _vp, _vp],
    "smfem_comm_prepare": [_vp, _vp],
    "smfem_comm_connect_local": [_vp, _vp, C.POINTER(_vp), C.c_int],
    # one process, several GPUs
    "smfem_init_multi": [C.c_int, C.POINTER(C.c_int), C.POINTER(_vp)],
    "smfem_multi_destroy": [_vp],
    "smfem_multi_size": [_vp, C.POINTER(C.c_int)],
    "smfem_multi_sync": [_vp],
    "smfem_multi_rank_handles": [_vp, _vp, _vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)],
    "smfem_multi_meshgrid": [_vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int,
                             C.POINTER(_vp)],
    "smfem_multi_inflate_sphere": [_vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double],
    "smfem_multi_assemble": [_vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(_vp)],
    "smfem_multi_assemble_system": [_vp, _f64p, _i64p, _i64p, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    C.c_double, C.c_double, C.POINTER(_vp), C.POINTER(_vp)],
    "smfem_multi_reassemble": [_vp, _vp, _vp, C.c_double, C.c_double],
    "smfem_multi_matrix_info": [_vp, _vp, _i64p, _i64p, _i64p],
    "smfem_multi_surface_mass": [_vp, _vp, _vp, C.c_double],
    "smfem_multi_set_dirichlet_zplanes": [_vp, _vp, _vp, C.c_double],
    "smfem_multi_pcg_use_multigrid": [_vp, _vp, _vp, C.c_int],
    "smfem_multi_pcg_set_warm_start": [_vp, _vp, C.c_double],
    "smfem_multi_pcg_solve": [_vp, _vp, C.c_double, C.c_int, _f64p, _f64p, C.POINTER(C.c_int), _f64p],
    "smfem_multi_pcg_stats": [_vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_int)],
    "smfem_multi_matrix_export_csc": [_vp, _vp, C.c_int, _i64p, _i64p, _f64p],
    "smfem_multi_matrix_free": [_vp, _vp],
    "smfem_multi_mesh_free": [_vp, _vp],
}
NON_STATUS = {"smfem_abi_version": (C.c_int, []), "smfem_last_error": (C.c_char_p, [])}


def lib():
    """Load the CUDA library (once).  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  smearfem_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        for name, (res, args) in NON_STATUS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise SmearFEMError(status, lib().smfem_last_error().decode("utf-8", "replace"))


def call(name, *args):
    check(getattr(lib(), name)(*args))
