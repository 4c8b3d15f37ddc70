"""BASELINE config 5 on N GPUs (torchrun, one rank per GPU, z-slabs): assembly-only sweep, pattern / values / fused timed
separately; time = max over ranks of the CUDA-event time between two barriers.  Rank 0 prints one JSON line per size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import smearfem_b200 as sf

sizes = [int(a) for a in sys.argv[1:]] or [200, 300, 400]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sf.Context(device=local, rank=rank, nranks=world)


def maxr(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for ne in sizes:
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)
    K.assemble_values(40.0, 0.4)
    reps = 5

    def t(f):
        f()
        ctx.sync()
        if world > 1:
            dist.barrier()
        ctx.timer_start()
        for _ in range(reps):
            f()
        return maxr(ctx.timer_stop() / reps)

    tp = t(lambda: K.pattern_rebuild())
    tv = t(lambda: K.assemble_values(40.0, 0.4))
    tf = t(lambda: K.reassemble(40.0, 0.4))
    i = K.info()
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, ne=ne, elements=ne**3, nnz=i["nnz"], nnz_local_rank0=i["nnz_local"], pattern_ms=tp, values_ms=tv,
                              fused_ms=tf, el_per_s_values=ne**3 / tv * 1e3, el_per_s_fused=ne**3 / tf * 1e3)), flush=True)
    K.free()
    mesh.free()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
