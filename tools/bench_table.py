"""Markdown table of DESIGN.md section 6 from bench lines: python tools/bench_table.py profiles/r2_final_bench_n{1,2,4,8}.json"""
import json
import sys

print("| N GPUs | mesh | elements/s (assembly step) | ms/step | fused kernel ms (GB/s, frac of HBM peak) | values-only ms (frac) | fp64 frac | SpMV ms (real-bytes frac; CSR-algorithmic frac) | matrix-free ms | Jacobi-PCG iters × ms/iter (matrix-free) | multigrid-PCG | ‖u−u*‖/‖u*‖ Jacobi / multigrid | e2e elements/s (ms/step) | pipeline to solution ms |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for f in sys.argv[1:]:
    d = json.load(open(f))
    r, s, p, e = d["roofline"], d["spmv"], d["pcg"], d["e2e"]
    mf = s.get("matrix_free") or {}
    pm = p.get("matrix_free") or {}
    g = p.get("multigrid") or {}
    m = p.get("manufactured_solution") or {}
    pipe = e.get("pipeline_to_solution") or {}
    print(f"| {d['n_gpus']} | {d['config']['ne']}³ | {d['value'] / 1e6:.0f} M | {d['ms_per_step']:.3f} | {r['ms_per_launch']:.3f} ({r['achieved']:.0f}, {r['frac']:.3f}) | "
          f"{r['values_only']['ms_per_launch']:.3f} ({r['values_only']['frac']:.3f}) | {r['fp64']['frac']:.3f} | "
          f"{s['ms']:.3f} ({s['frac']:.2f}; {s['csr_algorithmic']['frac']:.2f}) | {mf.get('ms', float('nan')):.3f} | "
          f"{p['iters']} × {p['ms_per_iter']:.3f} ({pm.get('ms_per_iter', float('nan')):.3f}) | {g.get('iters')} it, {g.get('ms_total', float('nan')):.0f} ms ({(p.get('multigrid_matrix_free') or {}).get('ms_total', float('nan')):.0f} ms matrix-free) | "
          f"{m.get('jacobi_pcg', {}).get('rel_u', float('nan')):.1e} / {m.get('multigrid_pcg', {}).get('rel_u', float('nan')):.1e} | "
          f"{e['value'] / 1e6:.0f} M ({e['ms_per_step']:.2f}) | {pipe.get('ms_total', float('nan')):.0f} |")
