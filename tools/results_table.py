"""Markdown tables of the bench records under profiles/ (README.md).
usage: python tools/results_table.py profiles/r2_bench_*.json"""
import json
import sys
from collections import defaultdict

recs = defaultdict(dict)
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception:
        continue
    if d.get("impl") == "reference":
        continue
    recs[d["config"]["name"]][d["n_gpus"]] = d

order = ["northstar", "c2", "c3", "c4", "c5"]
print("| config | GPUs | evals/s (device-timed) | ms/step | e2e evals/s | e2e ms | dominant kernel(s) ms | roofline (frac of measured peak) | collective |")
print("|---|---|---|---|---|---|---|---|---|")
for name in order:
    base = recs[name].get(1)
    for n in sorted(recs[name]):
        d = recs[name][n]
        r = d["roofline"]
        eff = f" ({d['value'] / (n * base['value']):.3f} of linear)" if base and n > 1 and base["config"]["n_rows_total"] == d["config"]["n_rows_total"] else ""
        effe = f" ({d['e2e']['value'] / (n * base['e2e']['value']):.3f})" if base and n > 1 and base["config"]["n_rows_total"] == d["config"]["n_rows_total"] else ""
        coll = d.get("collective")
        cs = "—"
        if coll:
            cs = f"{coll['bytes'] / 1024:.0f} KB: " + (f"peer {coll['peer_us']:.1f} µs, " if "peer_us" in coll else "") + f"NCCL {coll['nccl_us']:.1f} µs"
        rows = d["config"]["n_rows_total"]
        print(f"| {name} ({rows:.3g} rows) | {n} | {d['value']:.2f}{eff} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.2f}{effe} | {d['e2e']['ms_per_step']:.3f} | "
              f"{r['kernel_ms']:.3f} | {r['achieved']:.4g} {r['unit']} = {r['frac']:.3f} ({r['bound']}) | {cs} |")
print()
print("| config | GPUs | digest: objective | max abs gradient | trace(H) | parity_check (max rel obj / grad / Hessian, ranks) |")
print("|---|---|---|---|---|---|")
for name in order:
    for n in sorted(recs[name]):
        d = recs[name][n]
        g, pc = d["digest"], d["parity_check"]
        print(f"| {name} | {n} | {g['obj']:.15g} | {g['grad_inf']:.15g} | {g['hess_trace']:.15g} | "
              f"{'ok' if pc['ok'] else 'FAIL'}: {pc['max_rel']['obj']:.1e} / {pc['max_rel']['grad']:.1e} / {pc['max_rel']['hess']:.1e}, {pc['ranks']} |")
