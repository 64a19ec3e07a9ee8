#!/usr/bin/env python
"""Per-plan timing of the config-5 (first five layers) patterns: gather plans vs per-row elimination plans."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network
from continuousbayesiannetwork_b200.ve import PlanTooLarge, RowPlan
from tools.bench_kernels import timeit
dev = "cuda:0"
spec = synth.layered_dag()
tables, infer = install_cpts(spec, dev)
rows = 1 << 18
full = sample_network(spec, seed=1244, first=0, n=rows, device=dev, tables=tables)
rng = np.random.default_rng(1242)
tot = {"gather": 0.0, "rows": 0.0}
for p in range(64):
    kk = int(rng.integers(5, 51))
    vs = [int(v) for v in rng.choice(250, size=kk + 1, replace=False)]
    try:
        plan = infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
    except PlanTooLarge:
        continue
    ev = full[vs[1:]].contiguous()
    o = torch.empty((rows, plan.card_t), dtype=torch.float32, device=dev)
    us = timeit(lambda: plan.run_codes(ev, rows, out=o), 5, 2)
    kind = "rows" if isinstance(plan, RowPlan) else "gather"
    tot[kind] += us
    if kind == "rows":
        st = plan.stats
        sch = infer.compiler.last_row_schedule
        temps = sum((x["out_size"] + 3) // 4 * 4 for x in sch["steps"])
        terms = sum(x["out_size"] * x["sum_card"] * len(x["in_id"]) for x in sch["steps"])
        kern = "thread" if temps <= int(os.environ.get("CBN_ROWT_MAX_TEMPS", 96)) and terms <= int(os.environ.get("CBN_ROWT_MAX_TERMS", 480)) else "warp"
        outs = [x["out_size"] for x in sch["steps"]]
        print(f"pattern {p:2d} rows   k={kk:2d} hidden/row={st.per_row_hidden:3d} steps={len(sch['steps']):3d} temps={temps:5d} madds/row={st.per_row_madds:6d} {kern:6s} {us:9.1f} us  {rows/us:8.2f} M rows/s  {st.per_row_madds*rows/us/1e3:7.1f} Gmadd/s  out_sizes={outs}")
    else:
        print(f"pattern {p:2d} gather k={kk:2d} cells={[c for _, c in plan.stats.final_tables]} {us:9.1f} us  {rows/us:8.2f} M rows/s")
print(tot)
