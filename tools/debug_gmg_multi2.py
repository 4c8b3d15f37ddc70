"""Multi-GPU multigrid vs the one-GPU twin after 1, 2, 3 iterations (debugging aid)."""
import os
import sys
import time

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
rhs = None
if len(sys.argv) > 2 and sys.argv[2] == "mms":
    X, Y, Z = mesh1.nodelist()
    us = np.column_stack([0.01 * np.sin(np.pi * X) * np.cos(2 * Y) * Z, 0.01 * np.cos(X) * np.sin(np.pi * Y) * (1 + Z),
                          -0.001 * Z + 0.02 * np.sin(np.pi * Z) * (1 + X * Y)]).ravel()
    rhs = K1.spmv(us)
for maxit in (1, 2, 3, 5, 10, 80):
    q, it, rel = K.pcg_solve(rtol=1e-13, maxit=maxit, rhs_extra=None if rhs is None else rhs[r0:r0 + nr])
    q1, it1, rel1 = K1.pcg_solve(rtol=1e-13, maxit=maxit, rhs_extra=rhs)
    d = q - q1[r0:r0 + nr]
    n1 = ne + 1
    per_plane = np.abs(d).reshape(-1, 3 * n1 * n1).max(axis=1)
    worst = int(per_plane.argmax())
    if maxit == 1:
        prof = np.abs(d).reshape(-1, n1 * n1, 3).max(axis=1)
        say("per-plane max|dq| (x,y,z) at maxit 1: " + " ".join(f"{k}:{a:.0e}/{b:.0e}/{c:.0e}" for k, (a, b, c) in enumerate(prof) if k % 4 == 0 or k > per_plane.shape[0] - 4 or k < 3))
    say(f"maxit {maxit}: multi {it} its relres {rel:.2e} | one GPU {it1} its relres {rel1:.2e} | max|dq| {np.abs(d).max():.2e} (|q| {np.abs(q1).max():.2e}) worst local plane {worst} of {per_plane.shape[0]}")
dist.barrier()
dist.destroy_process_group()
