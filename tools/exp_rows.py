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
        print(f"pattern {p:2d} rows   k={kk:2d} hidden/row={st.per_row_hidden:3d} steps={len(plan.handle and [0]) and st.n_steps:3d} madds/row={st.contraction_madds:.3g} {us:9.1f} us  {rows/us:8.2f} M rows/s")
    else:
        print(f"pattern {p:2d} gather k={kk:2d} cells={[c for _, c in plan.stats.final_tables]} {us:9.1f} us  {rows/us:8.2f} M rows/s")
print(tot)
