"""e2e step (host arrays -> smfem_assemble_system -> diagonal on the host) for several host-thread counts of the lattice check."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smearfem_b200 as sf
from smearfem_b200 import _lib
from oracle import fem_oracle as o

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = sf.context()
NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne, 3)
o.inflate_sphere(NL, 0, 1, 0, 1)
pin = lambda a, dt: torch.from_numpy(np.asfortranarray(a, dtype=dt).T.copy()).pin_memory()   # .T of F-order = C-order view of the same bytes
NL_h, IEN_h, ID_h = pin(NL, np.float64), pin(IEN, np.int64), pin(ID, np.int64)
nN, nEl = NL.shape[1], IEN.shape[0]
diag_h = torch.empty(3 * nN, dtype=torch.float64, pin_memory=True)
_f, _i = C.POINTER(C.c_double), C.POINTER(C.c_int64)


def step():
    mh, kh = C.c_void_p(), C.c_void_p()
    _lib.call("smfem_assemble_system", ctx.handle, C.cast(NL_h.data_ptr(), _f), C.cast(IEN_h.data_ptr(), _i), C.cast(ID_h.data_ptr(), _i),
              nN, nEl, 8, ne, 3, _lib.Q1, 3, 40.0, 0.4, C.byref(mh), C.byref(kh))
    _lib.call("smfem_matrix_diag", ctx.handle, kh, C.cast(diag_h.data_ptr(), _f))
    _lib.lib().smfem_matrix_free(kh)
    _lib.lib().smfem_mesh_free(mh)


def moved():
    a, b = C.c_int64(), C.c_int64()
    _lib.call("smfem_transfer_bytes", ctx.handle, C.byref(a), C.byref(b))
    return a.value


print("host cores:", os.cpu_count())
for nt in ("0", "1", "2", "4", "8", None):
    if nt is None:
        os.environ.pop("SMFEM_HOST_THREADS", None)
    else:
        os.environ["SMFEM_HOST_THREADS"] = nt
    step(); step()
    ts = []
    m0 = moved()
    for _ in range(9):
        t0 = time.perf_counter(); step(); ctx.sync(); ts.append(time.perf_counter() - t0)
    mb = (moved() - m0) / 9 / 1e6
    ts.sort()
    print(f"host threads {nt or 'default':>7s}: median {ts[4]*1e3:.2f} ms  min {ts[0]*1e3:.2f}  max {ts[-1]*1e3:.2f}  -> {ne**3/ts[4]/1e6:.0f} M el/s;  "
          f"H2D {mb:.1f} MB/step;  trace {float(diag_h.sum()):.6f}")
