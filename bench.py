#!/usr/bin/env python
"""bench.py -- hex elements assembled/s and CG SpMV GB/s vs HBM peak (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (CPU restatement of the reference algorithm)

A step = one pass of the assembly hot path (device pattern build + element values) over a synthetic
inflated hex mesh resident in HBM.  Workload (weak scaling, cubic meshes as the reference's meshgrid
requires): ne = round(100 * N^(1/3)) -> 100^3 on 1 GPU (BASELINE config 3), 200^3 on 8 GPUs (config 4),
z-slab partitioned.  The same run then measures the CSR SpMV (GB/s vs the measured HBM copy peak)
and a Jacobi-PCG solve of the example problem, and the end-to-end path through the reference-facing
call with HOST mesh arrays.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hex elements assembled/s"
UNIT = "elements/s"


def ne_for(n_gpus):
    return int(round(100 * n_gpus ** (1.0 / 3.0)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def asm_bytes(ne, n_planes_owned=None):
    """Algorithmic bytes of BASELINE.md: values-only and pattern+values, for a slab of node planes."""
    n1 = ne + 1
    s1 = 3 * n1 - 2
    if n_planes_owned is None:
        n_planes_owned = n1
    frac = n_planes_owned / n1
    nnz = 9 * s1**3 * frac
    nN, nEl = n1**3 * frac, ne**3 * frac
    values = 8 * nnz + 24 * nN + 64 * nEl + 24 * nN
    total = values + 4 * nnz + 8 * (3 * nN + 1)
    # what the tile kernel itself has to move (the lattice path never touches IEN / ID; rowptr comes from its own small kernel):
    # values + diagonal written, coordinates read; the fused launch also writes the column indices
    k_values = 8 * nnz + 8 * 3 * nN + 24 * nN
    k_fused = k_values + 4 * nnz
    return values, total, k_values, k_fused


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(ne, world):
    """`config` of the JSON line: identical for the GPU arm and the reference arm (same workload, same partition)."""
    n1 = ne + 1
    return {"workload": f"hex{ne}: 3-D hex elasticity {ne}^3 elements, inflated unit cube (examples/vector3D.jl), E=40 nu=0.4",
            "ne": ne, "elements": ne**3, "ndof": 3 * n1**3, "nnz": 9 * (3 * n1 - 2) ** 3, "partition": f"z-slabs x{world}",
            "step": "device pattern build + element values, every entry of rowptr/colind/val rewritten each step",
            "l2": "no flush needed: each step writes K (>= 2.9 GB per GPU) >> 126 MB L2"}


def cpu_sample_layers(ne, target_elements=80000):
    return max(1, min(ne, int(round(target_elements / float(ne * ne)))))


def host_mesh_slab(ne, layers):
    """The first `layers` element layers of the inflated ne^3 lattice in the reference's host layout (meshgrid numbering:
    examples/vector3D.jl:60-127; inflate_sphere: src/PostProcess.jl:30-44): NodeList 3 x nN, IEN nEl x 8, ID nN x 3."""
    from oracle import fem_oracle as o

    n1 = ne + 1
    ax = o._julia_range(0.0, 1.0, n1)
    kk, jj, ii = np.meshgrid(np.arange(layers + 1), np.arange(n1), np.arange(n1), indexing="ij")
    NL = np.asfortranarray(np.stack([ax[ii.ravel()], ax[jj.ravel()], ax[kk.ravel()]]))
    o.inflate_sphere(NL, 0, 1, 0, 1)
    e = np.arange(ne * ne * layers, dtype=np.int64)
    base = (e // (ne * ne)) * n1 * n1 + ((e // ne) % ne) * n1 + e % ne + 1
    IEN = np.asfortranarray(np.stack([base + off for off in (0, 1, n1 + 1, n1, n1 * n1, n1 * n1 + 1, n1 * n1 + n1 + 1, n1 * n1 + n1)], axis=1))
    m = np.arange(NL.shape[1], dtype=np.int64)
    ID = np.asfortranarray(3 * m[:, None] + np.arange(1, 4, dtype=np.int64)[None, :])
    return NL, IEN, ID


def cpu_slab_rate(ne, layers, threads, mesh=None):
    """elements/s of the C restatement of the reference algorithm (element loop -> COO triplets -> sparse(), src/fem.jl:135-256)
    on a bounded sample of the ne^3 workload: its first `layers` element layers."""
    from oracle import c_oracle

    NL, IEN, ID = mesh if mesh is not None else host_mesh_slab(ne, layers)
    nEl = ne * ne * layers
    t = time.perf_counter()
    K = c_oracle.assemble_system(ne, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=threads, nEl=nEl)
    dt = time.perf_counter() - t
    s1 = 3 * (ne + 1) - 2
    assert K.nnz == 9 * s1 * s1 * (3 * (layers + 1) - 2), (K.nnz, ne, layers)   # the slab's own lattice pattern
    return nEl / dt, dt


def cpu_sample_text(ne, layers):
    return (f"the first {layers} element layers ({ne * ne * layers} elements) of the {ne}^3 inflated hex mesh per step: "
            "element loop -> 576 COO triplets per element -> sparse(), C restatement of src/fem.jl:135-256 (Julia is not installed: no oracle/_ref)")


def cpu_pcg_baseline(ne_cpu, iters=30):
    """CPU solve baseline beside the GPU one - Jacobi-PCG iterations with SciPy's CSR SpMV (one thread) on the
    example problem at ne_cpu (the reference's own dense inverse, examples/vector3D.jl:318, is O(n^3) and not timeable here).
    K comes from the oracle's C port: nothing on this leg touches the GPU library."""
    import scipy.sparse as sp
    from oracle import c_oracle, fem_oracle as o

    NL, IEN, ID, top, btm, _ = o.meshgrid(0, 1, 0, 1, 0, 1, ne_cpu, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    K = c_oracle.assemble_system(ne_cpu, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=c_oracle.max_threads())
    A = sp.csc_matrix((K.nzval, K.rowval - 1, K.colptr - 1), shape=(K.m, K.n)).tocsr()  # K is symmetric: CSC of K = CSR of K'
    n1 = ne_cpu + 1
    kz = np.arange(n1**3) // (n1 * n1)
    fixed = np.zeros(3 * n1**3, bool)
    fixed[3 * np.where((kz == 0) | (kz == ne_cpu))[0] + 2] = True
    qd = np.zeros(3 * n1**3)
    qd[3 * np.where(kz == ne_cpu)[0] + 2] = -0.001
    dinv = np.where(fixed, 0.0, 1.0 / A.diagonal())
    b = np.where(fixed, 0.0, -(A @ qd))
    x = np.zeros_like(b)
    r = b.copy()
    z = dinv * r
    p = z.copy()
    rz = r @ z
    t0 = time.perf_counter()
    for _ in range(iters):
        Ap = np.where(fixed, 0.0, A @ p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    dt = (time.perf_counter() - t0) / iters
    bytes_spmv = 12 * A.nnz + 24 * A.shape[0]
    return {"ms_per_iter": dt * 1e3, "spmv_GB/s": bytes_spmv / dt / 1e9, "cores": 1, "kind": "port", "iters_timed": iters,
            "sample": f"{ne_cpu}^3, Jacobi-PCG iterations with SciPy CSR SpMV (K without the surface term: same pattern and cost), one thread"}


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    threads = c_oracle.max_threads()   # every core this process may use; NOT OMP_NUM_THREADS (torchrun exports 1)
    ne = args.ne or ne_for(args.gpus)
    layers = cpu_sample_layers(ne)
    mesh = host_mesh_slab(ne, layers)
    for _ in range(max(args.warmup, 1)):
        cpu_slab_rate(ne, 1, threads, mesh=(mesh[0], mesh[1][: ne * ne], mesh[2]))
    t_tot, n_el = 0.0, 0
    for _ in range(args.steps):
        r, dt = cpu_slab_rate(ne, layers, threads, mesh=mesh)
        t_tot += dt
        n_el += ne * ne * layers
    val = n_el / t_tot
    one = cpu_slab_rate(ne, 1, 1, mesh=(mesh[0], mesh[1][: ne * ne], mesh[2]))[0]
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(ne, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample_text(ne, layers),
                         "value_1thread": one, "sample_1thread": f"one element layer ({ne * ne} elements), one thread (the reference's loop is serial, src/fem.jl:179)",
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    # The contract is ONE JSON line on stdout: libraries (NCCL's version banner, torchrun notices) also write to
    # fd 1, so park the real stdout and point fd 1 at stderr until the line is printed.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ne", type=int, default=0, help="override the mesh size (debug)")
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, emit)

    import torch

    import smearfem_b200 as sf
    from smearfem_b200 import distributed as sd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    affinity = None
    if world > 1:
        # several ranks on one host: run this rank (and allocate its pinned buffers: first touch) on the CPUs next to its GPU
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            affinity = len(os.sched_getaffinity(0))
        except Exception as exc:
            affinity = f"not set: {str(exc)[:80]}"
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from smearfem_b200 import _lib

    ne = args.ne or ne_for(world)
    n1 = ne + 1
    k0, k1 = sd.slab_range(n1, rank, world)
    ctx = sf.Context(device=local, rank=rank, nranks=world)
    mesh = sf.Mesh.meshgrid(ctx, 0, 1, 0, 1, 0, 1, ne, 3).inflate_sphere(0, 1, 0, 1)
    ctx.sync()
    hbm_peak, peak_src = peaks()
    traffic = {}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get(f"hex{ne}_n{world}", {})
    except Exception:
        pass

    # ------------------------------------------------------------------ assembly steps (the metric)
    K = sf.SparseMatrixB200.pattern(ctx, mesh, 3, 3)  # allocates K once; steps reuse the buffers

    def step():
        K.reassemble(40.0, 0.4)        # rowptr kernel + fused tile kernel (colind + values + diagonal, one pass over K)

    for _ in range(args.warmup):
        step()
    ctx.sync()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = ctx.launches
    ctx.timer_start()                  # CUDA events on the library's stream
    for _ in range(args.steps):
        step()
    ms_steps = ctx.timer_stop()
    barrier()
    launches = ctx.launches - l0
    _avg, _used = C.c_float(), C.c_int()
    _lib.call("smfem_assembly_kernel_ms", ctx.handle, min(args.steps, 64), C.byref(_avg), C.byref(_used))
    fused_kernel_ms = float(_avg.value)   # average over the timed steps' own launches of k_values_tile (CUDA events inside the library)
    # the dominant kernel alone, same inputs, same events
    ctx.timer_start()
    for _ in range(args.steps):
        K.assemble_values(40.0, 0.4)
    ms_values = ctx.timer_stop()
    barrier()
    t_step = max_over_ranks(ms_steps / args.steps * 1e-3)
    value = ne**3 / t_step
    b_values, b_total, bk_values, bk_fused = asm_bytes(ne, k1 - k0)
    # the step's dominant kernel: the fused launches of k_values_tile inside the timed steps, one CUDA event pair per launch
    t_val_kernel = max_over_ranks(ms_values / args.steps * 1e-3)
    t_fused_kernel = max_over_ranks(fused_kernel_ms * 1e-3)
    roof_val = bk_fused / t_fused_kernel / 1e9

    # ------------------------------------------------------------------ SpMV + PCG (second half of the metric)
    spmv = pcg = None
    if not args.no_solve:
        K.add_surface_mass(100.0)
        sd.connect(K)
        info = K.info()
        b_spmv = 12 * info["nnz_local"] + 20 * info["nrows_local"] + 4 * info["nrows_local"]  # CSR-algorithmic, int64 rowptr
        # what the row-triple kernel really has to read: 8 B per stored value; colind only for the rows that are not interior
        # lattice rows (closed-form columns there), once per row triple; rowptr + y + x once per row (x gathers hit L1/L2)
        n_int_planes = sum(1 for k in range(k0, k1) if 0 < k < n1 - 1)
        nnz_interior_rows = 81 * 3 * (n1 - 2) ** 2 * n_int_planes
        b_spmv_real = 8 * info["nnz_local"] + (4.0 / 3.0) * (info["nnz_local"] - nnz_interior_rows) + 24 * info["nrows_local"]
        names = {4: "row-triple (default)", 2: "csr-stream", 1: "warp-per-row"}
        tried = {}
        for variant in (4, 2, 1):
            barrier()
            ms = max_over_ranks(K.bench_spmv(reps=30, variant=variant))
            tried[names[variant]] = {"ms": ms, "GB/s_csr_algorithmic": b_spmv / (ms * 1e-3) / 1e9,
                                     "frac_csr_algorithmic": b_spmv / (ms * 1e-3) / 1e9 / hbm_peak}
        ms4 = tried[names[4]]["ms"]
        gbs4 = b_spmv_real / (ms4 * 1e-3) / 1e9
        # the same operator applied matrix-free from the coordinates (72 B per node of traffic, no CSR arrays)
        matfree = None
        try:
            K.use_matrix_free(True).use_matrix_free(False)   # names the mesh; variant 5 of bench_spmv switches the operator
            barrier()
            ms5 = max_over_ranks(K.bench_spmv(reps=30, variant=5))
            matfree = {"ms": ms5, "bytes_per_application_per_gpu": 72 * info["nrows_local"] // 3,
                       "note": "y = (K + beta b) x per element from coordinates + x (8 colour passes, no atomics, bit-reproducible); fp64-bound"}
        except Exception as exc:
            matfree = {"error": str(exc)[:200]}
        # roofline of the SpMV the solver uses (always the library default, variant 4; not chosen by timing)
        spmv = {"bound": "hbm", "kernel": "k_spmv_group3", "achieved": gbs4, "peak": hbm_peak, "unit": "GB/s", "frac": gbs4 / hbm_peak,
                "traffic": traffic.get("k_spmv_group3"), "traffic_source": "static: ncu --set full capture of the same kernel at this size, profiles/traffic.json (not re-measured in this run)",
                "GB/s_per_gpu": gbs4, "ms": ms4, "variant": 4,
                "bytes_per_spmv_per_gpu": b_spmv_real, "nnz_per_gpu": info["nnz_local"], "variants": tried, "matrix_free": matfree,
                "csr_algorithmic": {"bytes": b_spmv, "GB/s": b_spmv / (ms4 * 1e-3) / 1e9, "frac": b_spmv / (ms4 * 1e-3) / 1e9 / hbm_peak,
                                    "note": "BASELINE.md's 12 B/nnz + 24 B/row; exceeds 1.0 only because the kernel does not read colind for interior rows"},
                "note": "achieved = bytes the kernel must really read / time: 8 B/nnz values, colind (4 B/nnz) once per row triple and only for "
                        "non-interior rows, 24 B/row for rowptr + x + y; halo push and flag waits included for N > 1"}
        K.set_spmv_variant(4)
        K.set_dirichlet_zplanes(0.001)
        sd.barrier(ctx)
        _, it, relres = K.pcg_solve(rtol=1e-10, maxit=6000, want_q=False)
        st = K.pcg_stats()
        ms_tot = max_over_ranks(st["ms_total"])
        pcg = {"iters": it, "relres": relres, "ms_total": ms_tot, "ms_per_iter": ms_tot / max(it, 1),
               "spmv_GB/s_in_solve_per_gpu": b_spmv / (ms_tot / max(it, 1) * 1e-3) / 1e9, "rtol": 1e-10}
        ws = K.pcg_wait_stats()   # where an iteration waits for the other ranks (this rank's first CTA; max over ranks)
        pcg["wait_us_per_iter"] = {k.replace("_us", ""): max_over_ranks(v) / max(it, 1) for k, v in ws.items()}
        pcg["wait_us_per_iter"]["note"] = ("time the first CTA spins on the neighbours' halo flags (boundary-plane SpMV launches) and on the two "
                                           "mailbox all-reduces per iteration; %globaltimer, summed over the solve / iterations, max over ranks")
        sd.barrier(ctx)
        try:   # the same Jacobi-PCG with the matrix-free operator
            K.use_matrix_free(True)
            _, itf, relf = K.pcg_solve(rtol=1e-10, maxit=6000, want_q=False)
            msf = max_over_ranks(K.pcg_stats()["ms_total"])
            pcg["matrix_free"] = {"iters": itf, "relres": relf, "ms_total": msf, "ms_per_iter": msf / max(itf, 1)}
        except Exception as exc:
            pcg["matrix_free"] = {"error": str(exc)[:200]}
        K.use_matrix_free(False)
        sd.barrier(ctx)
        # the same solve with the geometric-multigrid V-cycle as preconditioner (distributed fine
        # levels + replicated coarse hierarchy on several GPUs)
        try:
            K.use_multigrid(True)
            K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)  # builds the hierarchy
            ms_first = max_over_ranks(K.pcg_stats()["ms_total"])
            _, itg, relg = K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
            pcg["multigrid"] = {"iters": itg, "relres": relg, "ms_total": max_over_ranks(K.pcg_stats()["ms_total"]),
                                "ms_first_solve_with_hierarchy_build": ms_first,
                                "note": "CG + V-cycle (re-assembled coarse levels, Chebyshev(2) smoothing); Jacobi-PCG above is the north-star path"}
            # ... and with the matrix-free operator for the fine-level products of the cycle and of the outer CG
            K.use_matrix_free(True)
            K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
            _, itf2, relf2 = K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
            pcg["multigrid_matrix_free"] = {"iters": itf2, "relres": relf2, "ms_total": max_over_ranks(K.pcg_stats()["ms_total"])}
            K.use_matrix_free(False)
            K.use_multigrid(False)
        except Exception as exc:  # never let the optional measurement take the bench line down
            pcg["multigrid"] = {"error": str(exc)[:200]}
            K.use_matrix_free(False)
            K.use_multigrid(False)
        # solution check at the bench size: a smooth manufactured field u* through the example's boundary conditions
        # (u_z = 0 on z = 0, u_z = -d on z = 1); rhs = K_bar u* with the library's own (multi-rank) SpMV; ||u - u*|| / ||u*||
        try:
            X, Y, Z = mesh.nodelist()
            us = np.column_stack([0.01 * np.sin(np.pi * X) * np.cos(2 * Y) * Z, 0.01 * np.cos(X) * np.sin(np.pi * Y) * (1 + Z),
                                  -0.001 * Z + 0.02 * np.sin(np.pi * Z) * (1 + X * Y)]).ravel()
            sd.barrier(ctx)
            rhs = K.spmv(us)
            sd.barrier(ctx)

            def rel_err(u):
                num, den = float(np.sum((u - us) ** 2)), float(np.sum(us**2))
                if dist is not None:
                    t = torch.tensor([num, den], dtype=torch.float64, device="cuda")
                    dist.all_reduce(t)
                    num, den = float(t[0].item()), float(t[1].item())
                return (num / den) ** 0.5

            man = {"field": "u* = smooth trigonometric field satisfying the example's Dirichlet data; rhs = K_bar u*", "rtol": 1e-13}
            uj, itj, relj = K.pcg_solve(rtol=1e-13, maxit=20000, rhs_extra=rhs)
            man["jacobi_pcg"] = {"rel_u": rel_err(uj), "iters": itj, "relres_true": relj, "ms_total": max_over_ranks(K.pcg_stats()["ms_total"])}
            sd.barrier(ctx)
            K.use_multigrid(True)
            ug, itm, relm = K.pcg_solve(rtol=1e-13, maxit=500, rhs_extra=rhs)
            man["multigrid_pcg"] = {"rel_u": rel_err(ug), "iters": itm, "relres_true": relm, "ms_total": max_over_ranks(K.pcg_stats()["ms_total"])}
            K.use_multigrid(False)
            pcg["manufactured_solution"] = man
            del us, rhs, uj, ug, X, Y, Z
        except Exception as exc:
            pcg["manufactured_solution"] = {"error": str(exc)[:300]}
        sd.barrier(ctx)
    clocks = sampler.stop()

    # ------------------------------------------------------------------ end to end: HOST mesh arrays -> K on device -> diag to host
    K.free()
    K = None

    nN, nEl = n1**3, ne**3
    NL_h = torch.empty((nN, 3), dtype=torch.float64, pin_memory=True)
    IEN_h = torch.empty((8, nEl), dtype=torch.int64, pin_memory=True)   # Julia column-major nEl x 8
    ID_h = torch.empty((3, nN), dtype=torch.int64, pin_memory=True)     # Julia column-major nNodes x 3
    # fill from the device mesh (rank-local slab is enough for coordinates on 1 GPU; build globally on host)
    ar = np.arange(n1, dtype=np.float64) / ne
    kk, jj, ii = np.meshgrid(np.arange(n1), np.arange(n1), np.arange(n1), indexing="ij")
    x, y, z = ar[ii.ravel()] - 0.5, ar[jj.ravel()] - 0.5, ar[kk.ravel()]
    scale = np.maximum(np.abs(x), np.abs(y))
    r = np.sqrt(x * x + y * y)
    r[scale == 0] = 1.0
    NLn = NL_h.numpy()
    NLn[:, 0], NLn[:, 1], NLn[:, 2] = scale * x / r, scale * y / r, z
    del x, y, z, scale, r, kk, jj, ii
    e = np.arange(nEl, dtype=np.int64)
    ei, ej, ek = e % ne, (e // ne) % ne, e // (ne * ne)
    base = ek * n1 * n1 + ej * n1 + ei + 1
    IENn = IEN_h.numpy()
    for a, off in enumerate([0, 1, n1 + 1, n1, n1 * n1, n1 * n1 + 1, n1 * n1 + n1 + 1, n1 * n1 + n1]):
        IENn[a] = base + off
    m = np.arange(nN, dtype=np.int64)
    IDn = ID_h.numpy()
    for l in range(3):
        IDn[l] = 3 * m + l + 1
    del e, ei, ej, ek, base, m
    nrows_local = 3 * (k1 - k0) * n1 * n1
    diag_h = torch.empty(nrows_local, dtype=torch.float64, pin_memory=True)
    _f = C.POINTER(C.c_double)
    _i = C.POINTER(C.c_int64)

    def e2e_step():
        mh, kh = C.c_void_p(), C.c_void_p()
        _lib.call("smfem_assemble_system", ctx.handle, C.cast(NL_h.data_ptr(), _f), C.cast(IEN_h.data_ptr(), _i),
                  C.cast(ID_h.data_ptr(), _i), nN, nEl, 8, ne, 3, _lib.Q1, 3, 40.0, 0.4, C.byref(mh), C.byref(kh))
        _lib.call("smfem_matrix_diag", ctx.handle, kh, C.cast(diag_h.data_ptr(), _f))
        _lib.lib().smfem_matrix_free(kh)
        _lib.lib().smfem_mesh_free(mh)

    e2e_steps = 7
    e2e_step()

    def moved():
        a, b = C.c_int64(), C.c_int64()
        _lib.call("smfem_transfer_bytes", ctx.handle, C.byref(a), C.byref(b))
        return a.value, b.value

    times = []
    m0 = moved()
    for _ in range(e2e_steps):
        barrier()
        t0 = time.perf_counter()
        e2e_step()          # blocking: returns after the diagonal is in host memory
        ctx.sync()
        times.append(time.perf_counter() - t0)
    m1 = moved()
    barrier()
    t_e2e = max_over_ranks(float(np.median(times)))   # median of 7 steps (PCIe transfers on shared hosts are noisy), max over ranks
    host_bytes = int(NL_h.numel() * 8 + IEN_h.numel() * 8 + ID_h.numel() * 8)
    e2e = {"value": ne**3 / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int((m1[0] - m0[0]) // e2e_steps),
           "d2h_bytes_per_step": int((m1[1] - m0[1]) // e2e_steps), "host_input_bytes_per_step": host_bytes,
           "ms_per_step": t_e2e * 1e3, "ms_per_step_all": [round(t * 1e3, 3) for t in times],
           "timing": "median of 7 steps, each bracketed by a barrier + stream sync, max over ranks",
           "call": "smfem_assemble_system(pinned NodeList, IEN, ID) -> smfem_matrix_diag (host)",
           "note": "NodeList always crosses PCIe; IEN/ID (Int64 connectivity, redundant on a lattice) are verified chunk by chunk, "
                   "partly on the device behind a copy, partly by host threads in place; h2d_bytes_per_step counts what was copied",
           "trace_check": float(diag_h.sum())}

    # e2e leg 2: the whole reference pipeline through the C ABI with HOST arrays in and the SOLUTION on the host out
    # (examples/vector3D.jl:302-322: assemble_system -> K + beta*b -> setboundaryCond -> solve), multigrid-PCG with the matrix-free
    # fine-level operator (the fastest configuration) to rtol 1e-10
    pipe = None
    if not args.no_solve:
        q_h = torch.empty(nrows_local, dtype=torch.float64, pin_memory=True)

        def pipeline():
            mh, kh = C.c_void_p(), C.c_void_p()
            t0 = time.perf_counter()
            _lib.call("smfem_assemble_system", ctx.handle, C.cast(NL_h.data_ptr(), _f), C.cast(IEN_h.data_ptr(), _i),
                      C.cast(ID_h.data_ptr(), _i), nN, nEl, 8, ne, 3, _lib.Q1, 3, 40.0, 0.4, C.byref(mh), C.byref(kh))
            Kp = sf.SparseMatrixB200(ctx, kh, sf.Mesh(ctx, mh))
            Kp.add_surface_mass(100.0)
            sd.connect(Kp)
            Kp.set_dirichlet_zplanes(0.001)
            Kp.use_multigrid(True)
            Kp.use_matrix_free(True)
            it, rel = C.c_int(), C.c_double()
            _lib.call("smfem_pcg_solve", ctx.handle, kh, 1e-10, 500, None, C.cast(q_h.data_ptr(), _f), C.byref(it), C.byref(rel))
            ctx.sync()
            dt = time.perf_counter() - t0
            solve_ms = Kp.pcg_stats()["ms_total"]
            Kp.free()
            Kp.mesh.free()
            return dt, int(it.value), float(rel.value), solve_ms

        try:
            pipeline()
            barrier()
            dt, itp, relp, solve_ms = pipeline()
            barrier()
            dt = max_over_ranks(dt)
            pipe = {"ms_total": dt * 1e3, "elements_per_s": ne**3 / dt, "pcg_iters": itp, "relres": relp, "solve_ms": max_over_ranks(solve_ms),
                    "h2d_bytes": host_bytes, "d2h_bytes": int(q_h.numel() * 8 * world), "u_checksum": float(q_h.sum()),
                    "call": "smfem_assemble_system(host NodeList, IEN, ID) -> smfem_surface_mass -> smfem_set_dirichlet_zplanes -> "
                            "smfem_pcg_use_multigrid + smfem_pcg_use_matrix_free -> smfem_pcg_solve(q on host); includes the multigrid hierarchy build"}
        except Exception as exc:
            pipe = {"error": str(exc)[:300]}
        del q_h
    e2e["pipeline_to_solution"] = pipe
    # e2e leg 3: assemble_system RETURNS K -- assembly from host arrays + export of the CSC arrays (what sparse(E,J,V) returned,
    # src/fem.jl:253) to host memory, at 50^3 (0.37 GB of CSC per call), one GPU
    if world == 1:
        try:
            NL5, IEN5, ID5, *_ = sf.meshgrid(0, 1, 0, 1, 0, 1, 50, 3)   # the product's own meshgrid / inflate_sphere (host arrays out)
            NL5 = sf.inflate_sphere(NL5, 0, 1, 0, 1)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                K5 = sf.assemble_system(50, NL5, IEN5, 3, "Q1", 3, ID5, 40, 0.4)
                cp5, rv5, nz5 = K5.to_csc()
                ts.append(time.perf_counter() - t0)
                K5.free()
            e2e["assemble_and_export_csc_50"] = {"ms": min(ts) * 1e3, "elements_per_s": 50**3 / min(ts), "csc_bytes_to_host": int(cp5.nbytes + rv5.nbytes + nz5.nbytes),
                                                 "call": "assemble_system(50, NodeList, IEN, 3, 'Q1', 3, ID, 40, 0.4) -> SparseMatrixCSC parts on the host (pageable arrays, Python mirror)"}
            del NL5, IEN5, ID5, cp5, rv5, nz5
        except Exception as exc:
            e2e["assemble_and_export_csc_50"] = {"error": str(exc)[:300]}

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1:
        from oracle import c_oracle

        th = c_oracle.max_threads()
        layers = cpu_sample_layers(ne, 160000)
        cmesh = host_mesh_slab(ne, layers)
        one_layer = (cmesh[0], cmesh[1][: ne * ne], cmesh[2])
        cpu_slab_rate(ne, 1, th, mesh=one_layer)   # warm-up (page faults of the first allocation)
        rate, dt = cpu_slab_rate(ne, layers, th, mesh=cmesh)
        rate1, dt1 = cpu_slab_rate(ne, 1, 1, mesh=one_layer)
        cpu = {"value": rate, "unit": UNIT, "cores": th, "kind": "port", "seconds": dt, "sample": cpu_sample_text(ne, layers),
               "value_1thread": rate1, "seconds_1thread": dt1,
               "sample_1thread": f"one element layer ({ne * ne} elements) of the same mesh, one thread (the reference's element loop is serial, src/fem.jl:179)"}
        del cmesh, one_layer
        try:
            cpu["solve"] = cpu_pcg_baseline(40)
        except Exception as exc:  # an optional figure must not take the line down
            cpu["solve"] = {"error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(ne, world),
            "roofline": {"bound": "hbm", "kernel": "k_values_tile2 (layer-march kernel), fused launch of the step (column indices + values + diagonal)",
                         "achieved": roof_val, "peak": hbm_peak, "unit": "GB/s", "frac": roof_val / hbm_peak,
                         "traffic": traffic.get("k_values_tile2_fused_colind", traffic.get("k_values_tile_fused_colind")),
                         "traffic_source": "static: ncu --set full capture of the same launch at this size, profiles/traffic.json (not re-measured in this run)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bk_fused, "ms_per_launch": t_fused_kernel * 1e3,
                         "launches_timed": int(_used.value),
                         "bytes": "12 B/nnz (value + column index) + 24 B/node coordinates read + 24 B/node diagonal written; "
                                  "IEN/ID/rowptr of SURVEY 8(d)'s 3.07 KB/element are not touched by this kernel and not counted",
                         "survey_bytes_per_step": b_total, "survey_GB/s_per_step": b_total / t_step / 1e9,
                         # the fp64 pipe is the co-equal ceiling of this kernel.  Algorithmic flops per element with the
                         # post-quadrature material: G = sum_gp g_a g_b' 8 x 64 x 9 FMA = 9216, gradients
                         # (J, adj, det, dN J^-1) ~350 x 8 gp = 2800, material ~570  -> 12.6 kflop; peak = DFMA rate measured by
                         # tools/microbench/peaks.cu / dmma.cu on this pool (34-37 TFLOP/s; DMMA runs at the same rate)
                         "fp64": {"flop_per_element": 12600, "achieved": 12600.0 * (ne**3 / world) / t_fused_kernel / 1e12, "peak": 36.5,
                                  "unit": "TFLOP/s", "frac": 12600.0 * (ne**3 / world) / t_fused_kernel / 1e12 / 36.5,
                                  "peak_source": "measured DFMA microbenchmark (tools/microbench), not in MEASURED_PEAKS.json"},
                         "values_only": {"ms_per_launch": t_val_kernel * 1e3, "algorithmic_bytes_per_launch": bk_values,
                                         "achieved": bk_values / t_val_kernel / 1e9, "frac": bk_values / t_val_kernel / 1e9 / hbm_peak,
                                         "traffic": traffic.get("k_values_tile"), "elements_per_s": ne**3 / t_val_kernel}},
            "spmv": spmv, "pcg": pcg, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "host": {"cpus_visible": len(os.sched_getaffinity(0)), "affinity_after_numa_pin": affinity},
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
