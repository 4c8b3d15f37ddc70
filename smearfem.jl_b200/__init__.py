"""smearfem_b200 -- host-side mirror of smearFEM.jl's call surface over the B200 C ABI.

The reference is a Julia package; Julia is not available in this image, so the host side above the
C ABI (include/smearfem_b200.h) is written in Python with the SAME function names, positional
arguments, defaults and array conventions as the reference:

    gaussian_quadrature, basis_function, assemble_system        src/fem.jl:21, :48, :135
    inflate_sphere                                               src/PostProcess.jl:30
    meshgrid, setboundaryCond, apply_boundary_conditions         examples/vector3D.jl:10, :133, :175
    solve (the idiom of examples/vector3D.jl:315-322)

Array conventions are Julia's: NodeList (ndim, nNodes) float64; IEN (nEl, nLocal) int64 1-based;
ID (nNodes, nDof) int64 1-based; sparse results expose SparseMatrixCSC parts (colptr, rowval,
nzval), 1-based.  (smearfem.jl_b200/julia/SmearFEMB200.jl is the equivalent `ccall` shim.)

All numerical work is done by hand-written sm_100a CUDA kernels inside libsmearfem_b200.so; there
is no CPU fallback and nothing here imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import SmearFEMError, call

__all__ = [
    "gaussian_quadrature", "basis_function", "assemble_system", "inflate_sphere", "meshgrid", "setboundaryCond",
    "apply_boundary_conditions", "solve", "load_steps", "greet_fem", "Context", "Mesh", "SparseMatrixB200", "SmearFEMError", "context",
]

_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def _pf(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


def _pi(a):
    return a.ctypes.data_as(_i64p) if a is not None else None


def _fclass(FunctionClass):
    if FunctionClass == "Q1":
        return _lib.Q1
    if FunctionClass == "Q2":
        return _lib.Q2
    raise SmearFEMError(_lib.ERR_INVALID, f"unknown FunctionClass {FunctionClass!r} (reference: UndefVarError)")


def greet_fem():
    """src/fem.jl:4-6"""
    print("Hello, I am the FEM module")


# ------------------------------------------------------------------------------------------------
# handles
# ------------------------------------------------------------------------------------------------
class Context:
    """One GPU / one process.  rank, nranks give the position in the z-slab partition."""

    def __init__(self, device=None, rank=None, nranks=None):
        if rank is None:
            rank = int(os.environ.get("RANK", "0")) if nranks is None else 0
        if nranks is None:
            nranks = int(os.environ.get("WORLD_SIZE", "1"))
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device, self.rank, self.nranks = device, rank, nranks
        h = C.c_void_p()
        call("smfem_init", device, rank, nranks, C.byref(h))
        self.handle = h

    def timer_start(self):
        call("smfem_timer_start", self.handle)

    def timer_stop(self):
        ms = C.c_float()
        call("smfem_timer_stop", self.handle, C.byref(ms))
        return float(ms.value)

    def sync(self):
        call("smfem_sync", self.handle)

    def flush_l2(self):
        call("smfem_flush_l2", self.handle)

    @property
    def launches(self):
        n = C.c_int64()
        call("smfem_launch_count", self.handle, C.byref(n))
        return int(n.value)

    def stream(self):
        s = C.c_void_p()
        call("smfem_stream", self.handle, C.byref(s))
        return s.value

    def close(self):
        if self.handle:
            _lib.lib().smfem_destroy(self.handle)
            self.handle = None


_default_ctx = None


def context():
    """The process-wide default context (created on first use; device = LOCAL_RANK)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class Mesh:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    @classmethod
    def meshgrid(cls, ctx, x0, x1, y0, y1, z0, z1, ne, ndim=3):
        """Device-side meshgrid (examples/vector3D.jl:10-130); only this rank's slab is materialised."""
        h = C.c_void_p()
        call("smfem_meshgrid", ctx.handle, float(x0), float(x1), float(y0), float(y1), float(z0), float(z1), int(ne), int(ndim),
             C.byref(h))
        return cls(ctx, h)

    @classmethod
    def from_host(cls, ctx, NodeList, IEN, ID, ndim, nDof, ne):
        NodeList = np.asfortranarray(NodeList, dtype=np.float64)
        IEN = np.asfortranarray(IEN, dtype=np.int64)
        if NodeList.ndim != 2 or NodeList.shape[0] != ndim:
            raise SmearFEMError(_lib.ERR_INVALID, "NodeList must be ndim x nNodes (reference: DimensionMismatch)")
        IDa = None if ID is None else np.asfortranarray(ID, dtype=np.int64)
        if IDa is not None and (IDa.ndim != 2 or IDa.shape[0] != NodeList.shape[1]):
            raise SmearFEMError(_lib.ERR_INVALID, "ID must be nNodes x nDof")
        nd = nDof if IDa is None else IDa.shape[1]
        h = C.c_void_p()
        call("smfem_mesh_from_host", ctx.handle, _pf(NodeList), _pi(IEN), _pi(IDa), NodeList.shape[1], IEN.shape[0], IEN.shape[1],
             int(ndim), int(nd), int(ne), C.byref(h))
        return cls(ctx, h)

    def info(self):
        nN, nE, n0, nO = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        nl, nd, st = C.c_int(), C.c_int(), C.c_int()
        call("smfem_mesh_info", self.handle, C.byref(nN), C.byref(nE), C.byref(nl), C.byref(nd), C.byref(st), C.byref(n0), C.byref(nO))
        return dict(nNodes=nN.value, nEl=nE.value, nLocal=nl.value, ndim=nd.value, structured=bool(st.value),
                    node0_owned=n0.value, nNodes_owned=nO.value)

    def colors(self):
        """Element colouring of a general mesh (built on first use): (ncolors, elements per colour); ncolors = -1 when the
        atomic scatter is used instead."""
        nc = C.c_int()
        sizes = np.zeros(64, dtype=np.int64)
        call("smfem_mesh_colors", self.ctx.handle, self.handle, C.byref(nc), _pi(sizes))
        return nc.value, sizes[: max(nc.value, 0)].copy()

    def inflate_sphere(self, x0, x1, y0, y1):
        call("smfem_inflate_sphere", self.ctx.handle, self.handle, float(x0), float(x1), float(y0), float(y1))
        return self

    def set_nodelist(self, NodeList_global):
        a = np.asfortranarray(NodeList_global, dtype=np.float64)
        call("smfem_mesh_set_nodelist", self.ctx.handle, self.handle, _pf(a))
        return self

    def nodelist(self):
        """(ndim, nNodes_owned) coordinates of this rank's owned nodes."""
        i = self.info()
        out = np.zeros((i["ndim"], i["nNodes_owned"]), order="F")
        call("smfem_mesh_export", self.ctx.handle, self.handle, _pf(out), None, None, None, None)
        return out

    def connectivity(self):
        """(IEN, ID, IEN_top, IEN_btm) in Julia layout (structured meshes; small ne)."""
        i = self.info()
        ne = round(i["nEl"] ** (1 / 3))
        IEN = np.zeros((i["nEl"], 8), dtype=np.int64, order="F")
        ID = np.zeros((i["nNodes"], 3), dtype=np.int64, order="F")
        top = np.zeros((ne * ne, 4), dtype=np.int64, order="F")
        btm = np.zeros((ne * ne, 4), dtype=np.int64, order="F")
        call("smfem_mesh_export", self.ctx.handle, self.handle, None, _pi(IEN), _pi(ID), _pi(top), _pi(btm))
        return IEN, ID, top, btm

    def free(self):
        if self.handle:
            _lib.lib().smfem_mesh_free(self.handle)
            self.handle = None


class SparseMatrixB200:
    """Device-resident K (this rank's row slab).  Stands in for SparseMatrixCSC{Float64,Int64}."""

    def __init__(self, ctx, handle, mesh):
        self.ctx, self.handle, self.mesh = ctx, handle, mesh

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def assemble(cls, ctx, mesh, ne, ndim, FunctionClass, nDof, Young, nu):
        h = C.c_void_p()
        call("smfem_assemble", ctx.handle, mesh.handle, int(ne), int(ndim), _fclass(FunctionClass), int(nDof), float(Young),
             float(nu), C.byref(h))
        return cls(ctx, h, mesh)

    @classmethod
    def matrix_free(cls, ctx, mesh, Young, nu):
        """The hex-lattice operator of assemble_system WITHOUT the assembled matrix (smfem_matfree_operator): solves, SpMVs and
        the multigrid preconditioner work on it; anything that needs CSR arrays raises."""
        h = C.c_void_p()
        call("smfem_matfree_operator", ctx.handle, mesh.handle, float(Young), float(nu), C.byref(h))
        return cls(ctx, h, mesh)

    @classmethod
    def pattern(cls, ctx, mesh, ndim, nDof):
        h = C.c_void_p()
        call("smfem_pattern_build", ctx.handle, mesh.handle, int(ndim), int(nDof), C.byref(h))
        return cls(ctx, h, mesh)

    def assemble_values(self, Young, nu):
        call("smfem_assemble_values", self.ctx.handle, self.mesh.handle, self.handle, float(Young), float(nu))
        return self

    def reassemble(self, Young, nu):
        """pattern + values into the existing buffers (fused colind/value kernel); asynchronous."""
        call("smfem_reassemble", self.ctx.handle, self.mesh.handle, self.handle, float(Young), float(nu))
        return self

    def pattern_rebuild(self):
        call("smfem_pattern_rebuild", self.ctx.handle, self.mesh.handle, self.handle)
        return self

    # -- queries ------------------------------------------------------------------------------
    def info(self):
        v = [C.c_int64() for _ in range(6)]
        call("smfem_matrix_info", self.handle, *[C.byref(x) for x in v])
        m, n, nnz, row0, nrl, nnzl = [x.value for x in v]
        return dict(m=m, n=n, nnz=nnz, row0=row0, nrows_local=nrl, nnz_local=nnzl)

    @property
    def shape(self):
        i = self.info()
        return (i["m"], i["n"])

    @property
    def nnz(self):
        return self.info()["nnz"]

    def to_csc(self, which=0):
        """(colptr, rowval, nzval): this rank's column slab as SparseMatrixCSC parts, 1-based."""
        i = self.info()
        colptr = np.zeros(i["nrows_local"] + 1, dtype=np.int64)
        rowval = np.zeros(i["nnz_local"], dtype=np.int64)
        nzval = np.zeros(i["nnz_local"], dtype=np.float64)
        call("smfem_matrix_export_csc", self.ctx.handle, self.handle, int(which), _pi(colptr), _pi(rowval), _pf(nzval))
        return colptr, rowval, nzval

    def to_scipy(self, which=0):
        import scipy.sparse as sp

        colptr, rowval, nzval = self.to_csc(which)
        i = self.info()
        return sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(i["m"], i["nrows_local"]))

    def diag(self):
        d = np.zeros(self.info()["nrows_local"])
        call("smfem_matrix_diag", self.ctx.handle, self.handle, _pf(d))
        return d

    # -- K_bar = K + beta*b  (examples/vector3D.jl:308), evaluated in place on the device ----------
    def add_surface_mass(self, beta, IEN_top=None, IEN_btm=None, keep_b=False, mesh=None):
        """K += beta * b in place.  mesh: the mesh whose coordinates b is integrated over (default: K's own);
        apply_boundary_conditions passes the NodeList IT was given (examples/vector3D.jl:306), which need not be K's."""
        t = None if IEN_top is None else np.asfortranarray(IEN_top, dtype=np.int64)
        b = None if IEN_btm is None else np.asfortranarray(IEN_btm, dtype=np.int64)
        nf = 0 if t is None else t.shape[0]
        m = self.mesh if mesh is None else mesh
        if m is not self.mesh and m.info()["structured"] != self.mesh.info()["structured"]:
            raise SmearFEMError(_lib.ERR_INVALID, "apply_boundary_conditions: IEN / ID describe a different kind of mesh than K's")
        call("smfem_surface_mass", self.ctx.handle, self.handle, m.handle, _pi(t), _pi(b), nf, float(beta), int(keep_b))
        return self

    def clone(self):
        """Device-side copy (pattern, values, diagonal)."""
        h = C.c_void_p()
        call("smfem_matrix_clone", self.ctx.handle, self.handle, C.byref(h))
        return SparseMatrixB200(self.ctx, h, self.mesh)

    def __add__(self, other):
        """K + beta*b (examples/vector3D.jl:308): a NEW device matrix, K itself stays K as in the reference; use
        add_surface_mass for the in-place form that saves the copy."""
        if isinstance(other, SurfaceMatrix):
            return other._add_into(self.clone())
        return NotImplemented

    __radd__ = __add__

    # -- Dirichlet + solve ----------------------------------------------------------------------
    def set_dirichlet_zplanes(self, d):
        call("smfem_set_dirichlet_zplanes", self.ctx.handle, self.handle, self.mesh.handle, float(d))
        return self

    def set_dirichlet(self, dofs, values):
        dofs = np.ascontiguousarray(dofs, dtype=np.int64)
        values = np.ascontiguousarray(values, dtype=np.float64)
        call("smfem_set_dirichlet", self.ctx.handle, self.handle, _pi(dofs), _pf(values), dofs.shape[0])
        return self

    def pcg_solve(self, rtol=1e-12, maxit=20000, rhs_extra=None, want_q=True, warm_scale=0.0):
        """Jacobi-PCG on the Dirichlet-masked operator.  warm_scale != 0: start from warm_scale * previous solution."""
        if warm_scale:
            call("smfem_pcg_set_warm_start", self.handle, float(warm_scale))
        n = self.info()["nrows_local"]
        q = np.zeros(n) if want_q else None
        ex = None if rhs_extra is None else np.ascontiguousarray(rhs_extra, dtype=np.float64)
        it, rel = C.c_int(), C.c_double()
        call("smfem_pcg_solve", self.ctx.handle, self.handle, float(rtol), int(maxit), _pf(ex), _pf(q), C.byref(it), C.byref(rel))
        return q, int(it.value), float(rel.value)

    def project_nodes(self, node_ids, CameraMatrix):
        """examples/vector3D.jl:325-329 on the device: (NodeList_new[:, ids], back_project(NodeList_new[:, ids], CameraMatrix))
        with the displacement of the last pcg_solve; node_ids are 1-based."""
        ids = np.ascontiguousarray(node_ids, dtype=np.int64).ravel()
        cam = np.asfortranarray(CameraMatrix, dtype=np.float64)
        p3 = np.zeros((3, ids.size), order="F")
        p2 = np.zeros((2, ids.size), order="F")
        call("smfem_project_nodes", self.ctx.handle, self.mesh.handle, self.handle, _pi(ids), ids.size, _pf(cam), _pf(p3), _pf(p2))
        return p3, p2

    def extract_borders(self, CameraMatrix, BorderNodesList, state, ne=None):
        """extract_borders(NodeList_new, CameraMatrix, BorderNodesList, state, ne) of src/PostProcess.jl:60-117 with
        NodeList_new = NodeList + motion of the last solve (examples/vector3D.jl:325-329): (BorderPoints, SideNodes2D)."""
        if state not in ("init", "update"):
            raise SmearFEMError(_lib.ERR_INVALID, "extract_borders: state must be 'init' or 'update' (reference: UndefVarError)")
        if state == "init" and ne is None:
            raise SmearFEMError(_lib.ERR_INVALID, "Number of elements must be provided")
        ids = np.ascontiguousarray(BorderNodesList[0], dtype=np.int64).ravel()
        cam = np.asfortranarray(CameraMatrix, dtype=np.float64)
        n = ids.size
        cap = n if state == "update" else 2 * (int(ne) + 1) + 2 * (n // (int(ne) + 1)) + 2
        border = np.zeros((2, max(cap, 1)), order="F")
        side = np.zeros((2, n), order="F")
        nb = C.c_int64()
        call("smfem_extract_borders", self.ctx.handle, self.mesh.handle, self.handle, _pi(ids), n, _pf(cam), 0 if state == "init" else 1,
             -1 if ne is None else int(ne), _pf(border), cap, C.byref(nb), _pf(side))
        return border[:, : nb.value].copy(), side

    def use_multigrid(self, enable=True):
        """Opt-in: later pcg_solve calls use CG preconditioned by a geometric multigrid V-cycle (hex lattice; with several
        ranks call distributed.connect(K) before the first solve)."""
        call("smfem_pcg_use_multigrid", self.ctx.handle, self.handle, self.mesh.handle if self.mesh is not None else None, int(bool(enable)))
        return self

    def use_matrix_free(self, enable=True):
        """Opt-in: the solve's operator K + beta*b is applied matrix-free from the mesh coordinates (hex lattice, nDof 3)."""
        call("smfem_pcg_use_matrix_free", self.ctx.handle, self.handle, self.mesh.handle if self.mesh is not None else None, int(bool(enable)))
        return self

    def apply_preconditioner(self, r):
        """z = M^-1 r for the multigrid V-cycle (after use_multigrid(True)); this rank's rows."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.zeros_like(r)
        call("smfem_pcg_apply_preconditioner", self.ctx.handle, self.handle, _pf(r), _pf(z))
        return z

    def pcg_stats(self):
        ms, ms2, it = C.c_float(), C.c_float(), C.c_int()
        call("smfem_pcg_stats", self.handle, C.byref(ms), C.byref(ms2), C.byref(it))
        return dict(ms_total=float(ms.value), iters=int(it.value))

    def pcg_wait_stats(self):
        """us the first CTA spent waiting on halo flags / the r'z all-reduce / the p'Ap all-reduce during the last Jacobi-PCG solve."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        call("smfem_pcg_wait_stats", self.handle, C.byref(a), C.byref(b), C.byref(c))
        return dict(halo_us=a.value, allreduce_rz_us=b.value, allreduce_pap_us=c.value)

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        call("smfem_spmv_host", self.ctx.handle, self.handle, _pf(x), _pf(y))
        return y

    def bench_spmv(self, reps=20, variant=0):
        ms = C.c_float()
        call("smfem_bench_spmv", self.ctx.handle, self.handle, int(variant), int(reps), C.byref(ms))
        return float(ms.value)

    def set_spmv_variant(self, variant):
        call("smfem_set_spmv_variant", self.handle, int(variant))
        return self

    def free(self):
        if self.handle:
            _lib.lib().smfem_matrix_free(self.handle)
            self.handle = None


# ------------------------------------------------------------------------------------------------
# one process, several GPUs (the reference host is one Julia process: examples/vector3D.jl:266-345)
# ------------------------------------------------------------------------------------------------
class MultiContext:
    """n GPUs driven from this process: rank r = z-slab r on device devices[r]; every method runs on all ranks concurrently
    (one worker thread per GPU inside the library, smfem_init_multi)."""

    def __init__(self, n_gpus, devices=None):
        dv = None if devices is None else (C.c_int * n_gpus)(*devices)
        h = C.c_void_p()
        call("smfem_init_multi", int(n_gpus), dv, C.byref(h))
        self.handle, self.n = h, int(n_gpus)

    def meshgrid(self, x0, x1, y0, y1, z0, z1, ne, ndim=3):
        h = C.c_void_p()
        call("smfem_multi_meshgrid", self.handle, float(x0), float(x1), float(y0), float(y1), float(z0), float(z1), int(ne), int(ndim), C.byref(h))
        return MultiMesh(self, h)

    def assemble_system(self, ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3):
        """assemble_system(...) of src/fem.jl:135 with the reference's host arrays; returns the n-GPU matrix."""
        NodeList = np.asfortranarray(NodeList, dtype=np.float64)
        IEN = np.asfortranarray(IEN, dtype=np.int64)
        IDa = None if ID is None else np.asfortranarray(ID, dtype=np.int64)
        mh, kh = C.c_void_p(), C.c_void_p()
        call("smfem_multi_assemble_system", self.handle, _pf(NodeList), _pi(IEN), _pi(IDa), NodeList.shape[1], IEN.shape[0], IEN.shape[1],
             int(ne), int(ndim), _fclass(FunctionClass), int(nDof), float(Young), float(nu), C.byref(mh), C.byref(kh))
        return MultiMatrix(self, kh, MultiMesh(self, mh))

    def sync(self):
        call("smfem_multi_sync", self.handle)

    def close(self):
        if self.handle:
            _lib.lib().smfem_multi_destroy(self.handle)
            self.handle = None


class MultiMesh:
    def __init__(self, mctx, handle):
        self.mctx, self.handle = mctx, handle

    def inflate_sphere(self, x0, x1, y0, y1):
        call("smfem_multi_inflate_sphere", self.mctx.handle, self.handle, float(x0), float(x1), float(y0), float(y1))
        return self

    def free(self):
        if self.handle:
            _lib.lib().smfem_multi_mesh_free(self.mctx.handle, self.handle)
            self.handle = None


class MultiMatrix:
    """K distributed over the GPUs of a MultiContext (row slabs); host vectors / CSC exports are global."""

    def __init__(self, mctx, handle, mesh):
        self.mctx, self.handle, self.mesh = mctx, handle, mesh

    @classmethod
    def assemble(cls, mctx, mesh, ne, ndim, FunctionClass, nDof, Young, nu):
        h = C.c_void_p()
        call("smfem_multi_assemble", mctx.handle, mesh.handle, int(ne), int(ndim), _fclass(FunctionClass), int(nDof), float(Young), float(nu), C.byref(h))
        return cls(mctx, h, mesh)

    def reassemble(self, Young, nu):
        call("smfem_multi_reassemble", self.mctx.handle, self.mesh.handle, self.handle, float(Young), float(nu))
        return self

    def info(self):
        m, n, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        call("smfem_multi_matrix_info", self.mctx.handle, self.handle, C.byref(m), C.byref(n), C.byref(nnz))
        return dict(m=m.value, n=n.value, nnz=nnz.value)

    def add_surface_mass(self, beta, mesh=None):
        call("smfem_multi_surface_mass", self.mctx.handle, self.handle, (mesh or self.mesh).handle, float(beta))
        return self

    def set_dirichlet_zplanes(self, d):
        call("smfem_multi_set_dirichlet_zplanes", self.mctx.handle, self.handle, self.mesh.handle, float(d))
        return self

    def use_multigrid(self, enable=True):
        call("smfem_multi_pcg_use_multigrid", self.mctx.handle, self.handle, self.mesh.handle, int(bool(enable)))
        return self

    def pcg_solve(self, rtol=1e-12, maxit=20000, rhs_extra=None, warm_scale=0.0):
        if warm_scale:
            call("smfem_multi_pcg_set_warm_start", self.mctx.handle, self.handle, float(warm_scale))
        q = np.zeros(self.info()["m"])
        ex = None if rhs_extra is None else np.ascontiguousarray(rhs_extra, dtype=np.float64)
        it, rel = C.c_int(), C.c_double()
        call("smfem_multi_pcg_solve", self.mctx.handle, self.handle, float(rtol), int(maxit), _pf(ex), _pf(q), C.byref(it), C.byref(rel))
        return q, int(it.value), float(rel.value)

    def pcg_stats(self):
        ms, it = C.c_float(), C.c_int()
        call("smfem_multi_pcg_stats", self.mctx.handle, self.handle, C.byref(ms), C.byref(it))
        return dict(ms_total=float(ms.value), iters=int(it.value))

    def to_csc(self, which=0):
        i = self.info()
        colptr = np.zeros(i["m"] + 1, dtype=np.int64)
        rowval = np.zeros(i["nnz"], dtype=np.int64)
        nzval = np.zeros(i["nnz"], dtype=np.float64)
        call("smfem_multi_matrix_export_csc", self.mctx.handle, self.handle, int(which), _pi(colptr), _pi(rowval), _pf(nzval))
        return colptr, rowval, nzval

    def rank_matrix(self, rank):
        """The per-rank objects behind this matrix (borrowed handles: do not free)."""
        ch, mh, kh = C.c_void_p(), C.c_void_p(), C.c_void_p()
        call("smfem_multi_rank_handles", self.mctx.handle, self.mesh.handle, self.handle, int(rank), C.byref(ch), C.byref(mh), C.byref(kh))
        ctx = Context.__new__(Context)
        ctx.device, ctx.rank, ctx.nranks, ctx.handle = None, rank, self.mctx.n, ch
        return SparseMatrixB200(ctx, kh, Mesh(ctx, mh))

    def free(self):
        if self.handle:
            _lib.lib().smfem_multi_matrix_free(self.mctx.handle, self.handle)
            self.handle = None


def load_steps(K_bar, deltas, rtol=1e-12, maxit=20000):
    """The load-stepping loop of examples/vector3D.jl:310-338 on a device-resident K̄ (structured mesh):
    for every prescribed top displacement d, set the Dirichlet data, solve, and yield (d, q, iterations).
    K̄ is assembled once; since q is linear in d every solve after the first is warm-started with q*d/d_prev."""
    prev = None
    for d in deltas:
        K_bar.set_dirichlet_zplanes(float(d))
        warm = (float(d) / prev) if prev else 0.0
        q, it, rel = K_bar.pcg_solve(rtol=rtol, maxit=maxit, warm_scale=warm)
        prev = float(d)
        yield float(d), q, it


class SurfaceMatrix:
    """What apply_boundary_conditions returns: b = ∫ N'N over the top and bottom faces, kept lazy so
    that `K + β*b` (examples/vector3D.jl:308) becomes one in-place device kernel on K's pattern."""

    def __init__(self, ctx, mesh, ne, IEN_top, IEN_btm, ID, beta=1.0):
        self.ctx, self.mesh, self.ne, self.beta = ctx, mesh, ne, beta
        self.IEN_top, self.IEN_btm, self.ID = IEN_top, IEN_btm, ID

    def __rmul__(self, beta):
        return SurfaceMatrix(self.ctx, self.mesh, self.ne, self.IEN_top, self.IEN_btm, self.ID, self.beta * float(beta))

    __mul__ = __rmul__

    def _add_into(self, K):
        # b is integrated over the NodeList apply_boundary_conditions received (self.mesh), not over K's mesh
        return K.add_surface_mass(self.beta, self.IEN_top, self.IEN_btm, mesh=self.mesh)

    def to_csc(self):
        """b as SparseMatrixCSC parts with ITS OWN pattern (pairs of dofs of nodes sharing a face)."""
        K = SparseMatrixB200.pattern(self.ctx, self.mesh, 3, 3).assemble_values(0.0, 0.25)
        K.add_surface_mass(0.0, self.IEN_top, self.IEN_btm, keep_b=True)  # K.mesh is self.mesh here
        colptr, rowval, nzval = K.to_csc(which=1)
        K.free()
        # structural entries of b: (dof of node a, dof of node b) with a, b in a common face
        ID = np.asarray(self.ID)
        ndof = int(ID.max())
        dof2node = np.zeros(ndof + 1, dtype=np.int64)
        dof2node[ID.ravel(order="F")] = np.tile(np.arange(1, ID.shape[0] + 1), ID.shape[1])
        faces = np.vstack([np.asarray(self.IEN_btm), np.asarray(self.IEN_top)])
        nN = ID.shape[0]
        a = np.repeat(faces, 4, axis=1).ravel()
        bb = np.tile(faces, (1, 4)).ravel()
        pair_keys = np.unique((a - 1) * nN + (bb - 1))
        cols = np.repeat(np.arange(1, colptr.shape[0]), np.diff(colptr))
        keys = (dof2node[rowval] - 1) * nN + (dof2node[cols] - 1)
        keep = np.isin(keys, pair_keys)
        newptr = np.zeros(ndof + 1, dtype=np.int64)
        np.add.at(newptr, cols[keep], 1)
        newptr = np.concatenate([[0], np.cumsum(newptr[1:])]) + 1
        # Julia's sparse(E,J,V) sizes b by the largest index present (examples/vector3D.jl:262)
        return newptr[: int(cols[keep].max()) + 1], rowval[keep], nzval[keep] * self.beta


class Constraint:
    """The constraint matrix C of setboundaryCond (identity with the constrained columns removed),
    stored as the 1-based ids of the kept columns."""

    def __init__(self, ndof, free):
        self.ndof, self.free = ndof, free

    @property
    def shape(self):
        return (self.ndof, self.free.shape[0])

    def to_scipy(self):
        import scipy.sparse as sp

        n = self.free.shape[0]
        return sp.csc_matrix((np.ones(n), (self.free - 1, np.arange(n))), shape=(self.ndof, n))


# ------------------------------------------------------------------------------------------------
# the reference's call surface
# ------------------------------------------------------------------------------------------------
def gaussian_quadrature(a, b, nGaussPoints=2):
    """src/fem.jl:21-31 -> (xi, w).  nGaussPoints outside {2,3} raises (reference: UndefVarError)."""
    n = int(nGaussPoints)
    xi = np.zeros(max(n, 1))
    w = np.zeros(max(n, 1))
    call("smfem_gaussian_quadrature", float(a), float(b), n, _pf(xi), _pf(w))
    return xi, w


def basis_function(xi, eta=None, zeta=None, FunctionClass="Q1"):
    """src/fem.jl:48-114 -> (N, dN) with dN (nnodes, ndim); the 1-D case keeps the 1x2 row quirk."""
    ndim = 1 if eta is None else (2 if zeta is None else 3)
    N = np.zeros(9)
    dN = np.zeros(27)
    nn = C.c_int()
    call("smfem_basis_function", ndim, _fclass(FunctionClass), float(xi), 0.0 if eta is None else float(eta),
         0.0 if zeta is None else float(zeta), _pf(N), _pf(dN), C.byref(nn))
    n = nn.value
    if ndim == 1:
        return N[:n].copy(), dN[:2].reshape(1, 2).copy()
    return N[:n].copy(), dN[: n * ndim].reshape((n, ndim), order="F").copy()


def meshgrid(x0, x1, y0, y1, z0, z1, ne, ndim):
    """examples/vector3D.jl:10-130 -> (NodeList, IEN, ID, IEN_top, IEN_btm, [Border, Bottom, Top]).
    3-D: coordinates generated by the device kernel; integer tables by the library's exporter."""
    ne = int(ne)
    if ndim == 3:
        ctx = context()
        if ctx.nranks != 1:
            raise SmearFEMError(_lib.ERR_UNSUPPORTED, "host-array meshgrid is single-process; use Mesh.meshgrid per rank")
        m = Mesh.meshgrid(ctx, x0, x1, y0, y1, z0, z1, ne, 3)
        NodeList = m.nodelist()
        IEN, ID, top, btm = m.connectivity()
        m.free()
        n1 = ne + 1
        k, j, i = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
        i, j, k = i.ravel(), j.ravel(), k.ravel()
        ids = np.arange(1, n1**3 + 1)
        side = (i == 0) | (i == ne) | (j == 0) | (j == ne)
        borders = [list(ids[side]), list(ids[~side & (k == 0)]), list(ids[~side & (k == ne) & (k != 0)])]
        return NodeList, IEN, ID, top, btm, borders
    if ndim == 2:  # examples/vector3D.jl:23-58: integer bookkeeping + two ranges, host side
        n1 = ne + 1
        x = x0 + (x1 - x0) * (np.arange(n1) / ne)
        y = y0 + (y1 - y0) * (np.arange(n1) / ne)
        x[0], x[-1], y[0], y[-1] = x0, x1, y0, y1
        jj, ii = np.meshgrid(np.arange(n1), np.arange(n1), indexing="ij")
        NodeList = np.asfortranarray(np.vstack([x[ii.ravel()], y[jj.ravel()]]))
        m = np.arange(1, n1 * n1 + 1)
        ID = np.asfortranarray(np.column_stack([2 * (m - 1) + 1, 2 * (m - 1) + 2]).astype(np.int64))
        ej, ei = np.meshgrid(np.arange(1, ne + 1), np.arange(1, ne + 1), indexing="ij")
        ej, ei = ej.ravel(), ei.ravel()
        IEN = np.asfortranarray(np.column_stack([(ej - 1) * n1 + ei, (ej - 1) * n1 + ei + 1, ej * n1 + ei + 1, ej * n1 + ei]).astype(np.int64))
        btm = np.zeros((ne, 2), dtype=np.int64)
        top = np.zeros((ne, 2), dtype=np.int64)
        btm[:, 0], btm[:, 1] = IEN[ej == 1, 0], IEN[ej == 1, 1]
        if ne > 1:
            top[:, 0], top[:, 1] = IEN[ej == ne, 3], IEN[ej == ne, 2]
        border = list(m[(ii.ravel() == 0) | (ii.ravel() == ne)])
        return NodeList, IEN, ID, top, btm, [border, [], []]
    raise SmearFEMError(_lib.ERR_UNSUPPORTED, "meshgrid: ndim must be 2 or 3")


def inflate_sphere(NodeList, x0, x1, y0, y1):
    """src/PostProcess.jl:30-44.  IN PLACE (and returned), like the reference."""
    if not (isinstance(NodeList, np.ndarray) and NodeList.dtype == np.float64 and NodeList.flags.f_contiguous):
        raise SmearFEMError(_lib.ERR_INVALID, "inflate_sphere mutates its argument: pass a float64 column-major (ndim,nNodes) array")
    call("smfem_inflate_sphere_host", context().handle, _pf(NodeList), NodeList.shape[0], NodeList.shape[1], float(x0), float(x1),
         float(y0), float(y1))
    return NodeList


def assemble_system(ne, NodeList, IEN, ndim, FunctionClass="Q1", nDof=1, ID=None, Young=1, nu=0.3):
    """src/fem.jl:135-256 -> K (device-resident SparseMatrixB200; `.to_csc()` gives Julia's CSC parts)."""
    ctx = context()
    if nDof > 1 and ID is None:
        raise SmearFEMError(_lib.ERR_INVALID, "ID is required when nDof > 1 (reference: MethodError size(nothing, 2))")
    NodeList = np.asfortranarray(NodeList, dtype=np.float64)
    IEN = np.asfortranarray(IEN, dtype=np.int64)
    if NodeList.ndim != 2 or NodeList.shape[0] != ndim:
        raise SmearFEMError(_lib.ERR_INVALID, "NodeList must be ndim x nNodes (reference: DimensionMismatch)")
    IDa = None if (ID is None or nDof == 1) else np.asfortranarray(ID, dtype=np.int64)
    if IDa is not None and (IDa.ndim != 2 or IDa.shape[0] != NodeList.shape[1]):
        raise SmearFEMError(_lib.ERR_INVALID, "ID must be nNodes x nDof")
    nd = nDof if IDa is None else IDa.shape[1]
    if nd != nDof:  # size(ID,2) != nDof: let the two-step path report it like before
        mesh = Mesh.from_host(ctx, NodeList, IEN, IDa, ndim, nDof, ne)
        return SparseMatrixB200.assemble(ctx, mesh, ne, ndim, FunctionClass, nDof, Young, nu)
    mh, kh = C.c_void_p(), C.c_void_p()
    call("smfem_assemble_system", ctx.handle, _pf(NodeList), _pi(IEN), _pi(IDa), NodeList.shape[1], IEN.shape[0], IEN.shape[1],
         int(ne), int(ndim), _fclass(FunctionClass), int(nDof), float(Young), float(nu), C.byref(mh), C.byref(kh))
    return SparseMatrixB200(ctx, kh, Mesh(ctx, mh))


def apply_boundary_conditions(ne, NodeList, IEN, IEN_top, IEN_btm, ndim, FunctionClass, ID, nDof=3):
    """examples/vector3D.jl:175-264 -> b (lazy SurfaceMatrix; `K + β*b` runs on the device)."""
    if ndim != 3:
        raise SmearFEMError(_lib.ERR_UNSUPPORTED, "apply_boundary_conditions: the reference's 2-D branch is not executable")
    _fclass(FunctionClass)
    ctx = context()
    mesh = Mesh.from_host(ctx, NodeList, IEN, ID, ndim, nDof, ne)
    return SurfaceMatrix(ctx, mesh, ne, np.asarray(IEN_top), np.asarray(IEN_btm), np.asarray(ID))


def setboundaryCond(NodeList, ne, ndim, FunctionClass, d, nDof=1):
    """examples/vector3D.jl:133-173 -> (q_d (ndof x 1), C).  Host data preparation, as upstream; the
    conditions themselves are applied inside the SpMV / CG kernels by `solve`."""
    if FunctionClass != "Q1":
        raise SmearFEMError(_lib.ERR_INVALID, "UndefVarError: q_d not defined (reference defines it for Q1 only)")
    ndof = nDof * (ne + 1) ** ndim
    q_d = np.zeros((ndof, 1))
    z = np.asarray(NodeList)[2]
    nodes = np.arange(1, z.shape[0] + 1)
    btm = z == 0
    top = (z == 1) & ~btm
    q_d[3 * nodes[top] - 1, 0] = -d
    rCol = np.concatenate([3 * nodes[btm], 3 * nodes[top]])
    free = np.setdiff1d(np.arange(1, ndim * (ne + 1) ** ndim + 1), rCol)
    return q_d, Constraint(ndim * (ne + 1) ** ndim, free)


def solve(K_bar, q_d, C_, rtol=1e-12, maxit=20000, return_info=False, multigrid=None):
    """examples/vector3D.jl:315-322: K_free = C'K̄C; q_f = K_free⁻¹ C'(-K̄ q_d); q = q_d + C q_f,
    with the dense inverse replaced by Jacobi-PCG on the Dirichlet-masked device operator.
    multigrid=True/False switches K̄ to / from the multigrid-preconditioned CG first (hex lattice, one GPU); None leaves it."""
    q_d = np.asarray(q_d, dtype=np.float64).reshape(-1)
    fixed = np.setdiff1d(np.arange(1, C_.ndof + 1), C_.free)
    K_bar.set_dirichlet(fixed, q_d[fixed - 1])
    if multigrid is not None:
        K_bar.use_multigrid(bool(multigrid))
    q, it, rel = K_bar.pcg_solve(rtol=rtol, maxit=maxit)
    if return_info:
        return q, dict(iters=it, relres=rel)
    return q
