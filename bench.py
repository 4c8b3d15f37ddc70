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


def cpu_port_rate(ne_cpu, threads):
    """elements/s of the C restatement of the reference algorithm (element loop -> COO -> sparse())."""
    from oracle import c_oracle, fem_oracle as o

    NL, IEN, ID, *_ = o.meshgrid(0, 1, 0, 1, 0, 1, ne_cpu, 3)
    o.inflate_sphere(NL, 0, 1, 0, 1)
    t = time.perf_counter()
    K = c_oracle.assemble_system(ne_cpu, NL, IEN, 3, "Q1", 3, ID, 40, 0.4, nthreads=threads)
    dt = time.perf_counter() - t
    assert K.nnz == 9 * (3 * (ne_cpu + 1) - 2) ** 3
    return ne_cpu**3 / dt, dt


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

    threads = c_oracle.max_threads()
    ne_cpu = 40
    for _ in range(args.warmup):
        cpu_port_rate(16, threads)
    t_tot, n_el = 0.0, 0
    for _ in range(args.steps):
        r, dt = cpu_port_rate(ne_cpu, threads)
        t_tot += dt
        n_el += ne_cpu**3
    val = n_el / t_tot
    sample = f"{ne_cpu}^3 inflated hex elements per step (same element type/material as the GPU workload, bounded sample)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"hex{ne_for(args.gpus)} (3-D hex elasticity, inflated unit cube, E=40, nu=0.4); CPU arm times a {ne_cpu}^3 sample"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of src/fem.jl:135-256 + sparse(); Julia itself is not installed (no oracle/_ref)"},
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
        b_spmv = 12 * info["nnz_local"] + 20 * info["nrows_local"] + 4 * info["nrows_local"]  # int64 rowptr
        names = {4: "row-triple (default)", 2: "csr-stream", 1: "warp-per-row"}
        tried = {}
        for variant in (4, 2, 1):
            barrier()
            ms = max_over_ranks(K.bench_spmv(reps=30, variant=variant))
            tried[names[variant]] = {"ms": ms, "GB/s": b_spmv / (ms * 1e-3) / 1e9}
        ms4 = tried[names[4]]["ms"]
        gbs4 = tried[names[4]]["GB/s"]
        # roofline of the SpMV the solver uses (always the library default, variant 4; not chosen by timing)
        spmv = {"bound": "hbm", "kernel": "k_spmv_group3", "achieved": gbs4, "peak": hbm_peak, "unit": "GB/s", "frac": gbs4 / hbm_peak,
                "traffic": traffic.get("k_spmv_group3"), "GB/s_per_gpu": gbs4, "ms": ms4, "variant": 4,
                "bytes_per_spmv_per_gpu": b_spmv, "nnz_per_gpu": info["nnz_local"], "variants": tried,
                "note": "achieved = CSR-algorithmic bytes (12 B/nnz + 24 B/row) / time; the kernel reads colind once per row triple "
                        "and interior lattice rows need no colind at all (8.3 B/nnz of real traffic, see `traffic`), and a read-only stream runs above the copy peak (DESIGN.md 4)"}
        K.set_spmv_variant(4)
        K.set_dirichlet_zplanes(0.001)
        sd.barrier(ctx)
        _, it, relres = K.pcg_solve(rtol=1e-10, maxit=6000, want_q=False)
        st = K.pcg_stats()
        ms_tot = max_over_ranks(st["ms_total"])
        pcg = {"iters": it, "relres": relres, "ms_total": ms_tot, "ms_per_iter": ms_tot / max(it, 1),
               "spmv_GB/s_in_solve_per_gpu": b_spmv / (ms_tot / max(it, 1) * 1e-3) / 1e9, "rtol": 1e-10}
        sd.barrier(ctx)
        if world == 1:
            # SURVEY 8(f) row 3 (opt-in, one GPU): the same solve with the geometric-multigrid V-cycle as preconditioner
            try:
                K.use_multigrid(True)
                K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)  # builds the hierarchy
                _, itg, relg = K.pcg_solve(rtol=1e-10, maxit=500, want_q=False)
                pcg["multigrid"] = {"iters": itg, "relres": relg, "ms_total": K.pcg_stats()["ms_total"],
                                    "note": "CG + V-cycle (re-assembled coarse levels, Chebyshev(2) smoothing); Jacobi-PCG above is the north-star path"}
                K.use_multigrid(False)
            except Exception as exc:  # never let the optional measurement take the bench line down
                pcg["multigrid"] = {"error": str(exc)[:200]}
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

    # ------------------------------------------------------------------ CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1:
        from oracle import c_oracle

        th = c_oracle.max_threads()
        ne_cpu = 40
        rate, dt = cpu_port_rate(ne_cpu, th)
        cpu = {"value": rate, "unit": UNIT, "cores": th, "kind": "port", "seconds": dt,
               "sample": f"{ne_cpu}^3 inflated hex elements, C restatement of src/fem.jl:135-256 + sparse() (Julia not installed)"}
        try:
            cpu["solve"] = cpu_pcg_baseline(ne_cpu)
        except Exception as exc:  # an optional figure must not take the line down
            cpu["solve"] = {"error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"hex{ne}: 3-D hex elasticity {ne}^3 elements, inflated unit cube (examples/vector3D.jl), E=40 nu=0.4",
                       "ne": ne, "elements": ne**3, "ndof": 3 * n1**3, "nnz": 9 * (3 * n1 - 2) ** 3, "partition": f"z-slabs x{world}",
                       "step": "device pattern build + element values, every entry of rowptr/colind/val rewritten each step",
                       "l2": "no flush needed: each step writes K (>= 2.9 GB per GPU) >> 126 MB L2"},
            "roofline": {"bound": "hbm", "kernel": "k_values_tile, fused launch of the step (column indices + values + diagonal)",
                         "achieved": roof_val, "peak": hbm_peak, "unit": "GB/s", "frac": roof_val / hbm_peak,
                         "traffic": traffic.get("k_values_tile_fused_colind"), "peak_source": peak_src,
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
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
