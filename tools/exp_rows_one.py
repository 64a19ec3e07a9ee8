#!/usr/bin/env python
"""One per-row plan of the layered-DAG patterns (default: pattern 28, the largest schedule) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network
which = int(sys.argv[1]) if len(sys.argv) > 1 else 28
dev = "cuda:0"
spec = synth.layered_dag()
tables, infer = install_cpts(spec, dev)
rows = 1 << 18
full = sample_network(spec, seed=1244, first=0, n=rows, device=dev, tables=tables)
rng = np.random.default_rng(1242)
for p in range(64):
    kk = int(rng.integers(5, 51))
    vs = [int(v) for v in rng.choice(250, size=kk + 1, replace=False)]
    if p == which:
        plan = infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
        ev = full[vs[1:]].contiguous()
        o = torch.empty((rows, plan.card_t), dtype=torch.float32, device=dev)
        for _ in range(3):
            plan.run_codes(ev, rows, out=o)
        torch.cuda.synchronize()
        print("pattern", p, type(plan).__name__, plan.stats.per_row_madds)
