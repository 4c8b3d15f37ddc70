"""Multi-GPU plumbing: one process per GPU (torchrun), z-slab partition of node planes.

torch.distributed is used ONLY to exchange the 64-byte CUDA IPC handles of the ranks' peer windows
and for host-level barriers; afterwards the CG kernels write halo planes and dot-product partials
straight into peer memory over NVLink (see csrc/solver.cu).  Works with the gloo backend for the
handle exchange, so the host logic is testable on CPU with world_size 2."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def slab_range(n1, rank, nranks):
    """Owned node planes [k0, k1) of `rank` (same formula as csrc/smfem_internal.cuh:slab_range)."""
    return (n1 * rank) // nranks, (n1 * (rank + 1)) // nranks


def slab_rows(ne, nDof, rank, nranks):
    n1 = ne + 1
    k0, k1 = slab_range(n1, rank, nranks)
    return k0 * n1 * n1 * nDof, (k1 - k0) * n1 * n1 * nDof


def gather_handles(my_handle: bytes, group=None):
    """all_gather of the per-rank IPC handles -> bytes of length world_size*64, rank-ordered."""
    import torch
    import torch.distributed as dist

    assert len(my_handle) == _lib.IPC_HANDLE_BYTES
    ws = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(my_handle), dtype=torch.uint8).to(dev)
    outs = [torch.empty(_lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(ws)]
    dist.all_gather(outs, mine, group=group)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in outs)


def connect(K, group=None):
    """Create K's peer window, exchange handles, map the peers.  Collective over `group`."""
    import torch.distributed as dist

    ctx = K.ctx
    buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
    _lib.call("smfem_comm_export", ctx.handle, K.handle, buf)
    if ctx.nranks == 1:
        return K
    allh = gather_handles(buf.raw, group)
    hb = C.create_string_buffer(allh, len(allh))
    _lib.call("smfem_comm_connect", ctx.handle, K.handle, hb)
    ctx.sync()
    dist.barrier(group)
    return K


def barrier(ctx, group=None):
    """Device + host barrier between solves (the mailbox protocol needs one)."""
    import torch.distributed as dist

    ctx.sync()
    if ctx.nranks > 1:
        dist.barrier(group)


def gather_vector(local: np.ndarray, group=None):
    """Concatenate the ranks' row slabs on every rank (host side; tests / small problems)."""
    import torch
    import torch.distributed as dist

    ws = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros(mx, dtype=torch.float64, device=dev)
    pad[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(dev)
    outs = [torch.zeros(mx, dtype=torch.float64, device=dev) for _ in range(ws)]
    dist.all_gather(outs, pad, group=group)
    return np.concatenate([o[:s].cpu().numpy() for o, s in zip(outs, sizes)])
