"""One application of the V-cycle: N ranks vs the one-GPU twin, for a random and a smooth input, with a depth limit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import smearfem_b200 as sf
from smearfem_b200 import distributed as sd

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 81
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", local))


def say(*a):
    print(f"[r{rank}]", *a, file=sys.stderr, flush=True)


ctx = sf.Context(device=local, rank=rank, nranks=ws)
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
sd.connect(K)
K.set_dirichlet_zplanes(0.001)
info = K.info()
r0, nr = info["row0"], info["nrows_local"]
ctx1 = sf.Context(device=local, rank=0, nranks=1)
mesh1 = sf.Mesh.meshgrid(ctx1, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K1 = sf.SparseMatrixB200.assemble(ctx1, mesh1, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
K1.set_dirichlet_zplanes(0.001)
K.use_multigrid(True)
K1.use_multigrid(True)
n = K1.shape[0]
n1 = ne + 1
X, Y, Z = mesh1.nodelist()
inputs = {"random": np.random.default_rng(1).standard_normal(n),
          "smooth": np.column_stack([np.sin(2 * X + 1) * (1 + Z), np.cos(3 * Y) + Z, 1 + X * Y + Z * Z]).ravel(),
          "const_x": np.column_stack([np.ones_like(X), 0 * X, 0 * X]).ravel(),
          "const_z": np.column_stack([0 * X, 0 * X, np.ones_like(X)]).ravel()}
for depth in (2,):
    os.environ["SMFEM_GMG_DEBUG_DEPTH"] = str(depth)
    for name, r in list(inputs.items())[3:]:
        os.environ["SMFEM_GMG_TRACE"] = "1"
        z = K.apply_preconditioner(r[r0:r0 + nr])
        z1 = K1.apply_preconditioner(r)[r0:r0 + nr]
        d = np.abs(z - z1).reshape(-1, n1 * n1, 3).max(axis=(1, 2))
        say(f"depth {depth} {name:8s}: max|z - z1| {np.abs(z - z1).max():.2e} (|z1| {np.abs(z1).max():.2e}); planes with error > 1e-9 |z|: {np.nonzero(d > 1e-9 * np.abs(z1).max())[0][:12].tolist()}")
dist.barrier()
dist.destroy_process_group()
