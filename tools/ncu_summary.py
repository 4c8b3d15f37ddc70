"""Readable summary of an `ncu --set full` report: per kernel launch the throughput triage (DRAM / L2 / L1TEX / fp64 pipe / issue
slots), occupancy, instruction and shared-memory counts, DRAM bytes, and the warp-state breakdown (stalled warps per issued
instruction, largest first).  Input: the CSV of `ncu -i <report> --page raw --csv`.  No GPU needed.
usage: ncu_summary.py raw.csv [title]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, k, default=float("nan")):
    try:
        return float(r[col[k]].replace(",", ""))
    except (KeyError, ValueError):
        return default


def u(k):
    return units[col[k]] if k in col else ""


print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("| # | kernel | grid x block | regs | ms | DRAM GB (rd+wr) | DRAM % | L2 % | L1TEX % | fp64 pipe % | issue slots % | warps/SM active | warp instr | shared wavefronts (bank conflicts) | local-memory instr |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in data:
    name = r[col["Kernel Name"]].replace("<unnamed>::", "").replace("void ", "")
    rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
    scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
    rd *= scale.get(u("dram__bytes_read.sum"), 1.0)
    wr *= scale.get(u("dram__bytes_write.sum"), 1.0)
    ms = f(r, "gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u("gpu__time_duration.sum"), 1.0)
    wav = f(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    conf = f(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    loc = f(r, "smsp__sass_inst_executed_op_local_ld.sum", 0.0) + f(r, "smsp__sass_inst_executed_op_local_st.sum", 0.0)
    print(f"| {r[col['ID']]} | `{name[:70]}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | {f(r, 'launch__registers_per_thread'):.0f} | {ms:.3f} | "
          f"{rd + wr:.3f} ({rd:.2f}+{wr:.2f}) | {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{f(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {f(r, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{f(r, 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed'):.0f} | {f(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{f(r, 'sm__warps_active.avg.per_cycle_active'):.1f} | {f(r, 'smsp__inst_executed.sum'):.3g} | {wav:.3g} ({conf:.3g}) | {loc:.3g} |")
print("\nWarp states (average number of warps per scheduler in that state per issued instruction; the sum is the scheduler's resident")
print("warps / issue rate - a state that is large compared with 1 is where the warps sit instead of issuing):\n")
for r in data:
    st = sorted(((f(r, k, 0.0), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stall_keys), reverse=True)
    print(f"* #{r[col['ID']]}: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:7] if v > 0.005))
