#!/usr/bin/env python
"""The 200-node DAG query leg alone: 8 patterns x 1M rows, every pattern rotating through a ring of 6 batches; per-launch time
of the gather kernel on one stream (pattern-major, eager back-to-back launches and one CUDA graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network

dev = "cuda:0"
spec = synth.random_ktree_dag()
t, infer = install_cpts(spec, dev)
rng = np.random.default_rng(1240)
rows, n_ring = 1 << 20, 6
fulls = [sample_network(spec, seed=1241 + r, first=0, n=rows, device=dev, tables=t) for r in range(n_ring)]
pats = []
for _ in range(8):
    vs = [int(v) for v in rng.choice(spec.n, size=11, replace=False)]
    plan = infer.plan(spec.names[vs[0]], [spec.names[v] for v in vs[1:]])
    plan.set_static_evidence(True)
    pats.append((plan, [(f[vs[1:]].contiguous(), torch.empty((rows, plan.card_t), dtype=torch.float32, device=dev)) for f in fulls]))
del fulls
alg = sum(rows * p.algorithmic_bytes_per_row() for p, _ in pats)

def one_pass_pattern_major(reps):
    for plan, ring in pats:
        for r in range(reps):
            ev, o = ring[r % n_ring]
            plan.run_codes(ev, rows, out=o)

def timed(fn, iters):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / iters

reps = 6
s = timed(lambda: one_pass_pattern_major(reps), 5) / reps
print(f"eager, one stream, pattern-major: {s / 8 * 1e6:6.2f} us per launch, {alg / s / 1e9:7.1f} GB/s algorithmic ({alg / s / 1e9 / 6521.1:.3f} of peak)")
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    one_pass_pattern_major(reps)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=st):
        one_pass_pattern_major(reps)
s = timed(g.replay, 5) / reps
print(f"graph, one stream, pattern-major: {s / 8 * 1e6:6.2f} us per launch, {alg / s / 1e9:7.1f} GB/s algorithmic ({alg / s / 1e9 / 6521.1:.3f} of peak)")
