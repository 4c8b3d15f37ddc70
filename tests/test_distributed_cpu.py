"""World-size-2 gloo tests of the multi-GPU host logic (slab partition, handle exchange, slab
gather); the device side of the same path is covered by tests/test_gpu_multi.py on the GPU box."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_all_planes():
    from smearfem_b200.distributed import slab_range, slab_rows

    for n1 in (3, 21, 101, 201):
        for nr in (1, 2, 4, 8):
            if n1 < nr:
                continue
            rs = [slab_range(n1, r, nr) for r in range(nr)]
            assert rs[0][0] == 0 and rs[-1][1] == n1
            assert all(rs[i][1] == rs[i + 1][0] for i in range(nr - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
            rows = [slab_rows(n1 - 1, 3, r, nr) for r in range(nr)]
            assert sum(n for _, n in rows) == 3 * n1**3
            assert all(rows[i][0] + rows[i][1] == rows[i + 1][0] for i in range(nr - 1))


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    import torch.distributed as dist

    from smearfem_b200.distributed import gather_handles, gather_vector, slab_rows

    dist.init_process_group("gloo", rank=rank, world_size=ws)
    h = bytes([rank * 16 + i % 16 for i in range(64)])
    allh = gather_handles(h)
    ok = len(allh) == 64 * ws and all(allh[64 * r:64 * (r + 1)] == bytes([r * 16 + i % 16 for i in range(64)]) for r in range(ws))
    row0, n = slab_rows(4, 3, rank, ws)
    full = gather_vector(np.arange(row0, row0 + n, dtype=np.float64))
    ok = ok and np.array_equal(full, np.arange(3 * 125, dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_handle_exchange_and_slab_gather_gloo_ws2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
