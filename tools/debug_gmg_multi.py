"""Stage-by-stage run of the multi-GPU multigrid-PCG (debugging aid): torchrun --nproc-per-node N tools/debug_gmg_multi.py ne"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import smearfem_b200 as sf
from smearfem_b200 import distributed as sd

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rank, ws, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", local))
t0 = time.time()


def say(*a):
    print(f"[r{rank} {time.time() - t0:6.2f}s]", *a, file=sys.stderr, flush=True)


ctx = sf.Context(device=local, rank=rank, nranks=ws)
mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
K = sf.SparseMatrixB200.assemble(ctx, mesh, ne, 3, "Q1", 3, 40, 0.4).add_surface_mass(100.0)
sd.connect(K)
K.set_dirichlet_zplanes(0.001)
say("assembled, connected")
q, it, rel = K.pcg_solve(rtol=1e-10, maxit=20000)
say(f"jacobi: {it} its relres {rel:.2e} {K.pcg_stats()['ms_total']:.1f} ms")
K.use_multigrid(True)
say("multigrid enabled")
for rep in range(3):
    qm, itm, relm = K.pcg_solve(rtol=1e-10, maxit=100)
    say(f"gmg solve {rep}: {itm} its relres {relm:.2e} {K.pcg_stats()['ms_total']:.1f} ms, |q - q_jacobi|/|q| = {np.linalg.norm(qm - q) / np.linalg.norm(q):.2e}")
dist.barrier()
dist.destroy_process_group()
