#!/usr/bin/env python
"""The reference-facing API on the 1000-node layered DAG: BayesianNetwork(dag, DataFrame) fit + infer timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import networkx as nx, numpy as np, pandas as pd, torch
from continuousbayesiannetwork_b200 import BayesianNetwork, synth
spec = synth.layered_dag()
n = 50_000
codes = synth.sample_forward_numpy(spec, 3, 0, n)
df = pd.DataFrame({nm: codes[i].astype(np.float32) for i, nm in enumerate(spec.names)})
dag = nx.DiGraph(); dag.add_nodes_from(spec.names)
dag.add_edges_from([(spec.names[p], spec.names[i]) for i in range(spec.n) for p in spec.parents[i]])
t0 = time.perf_counter()
bn = BayesianNetwork(dag, df, {"estimator_name": "brute_force"}, {"inference_obj": "exact"}, device="cuda:0")
torch.cuda.synchronize()
print(f"BayesianNetwork(1000 nodes, {n} rows): {time.perf_counter() - t0:.2f} s")
t0 = time.perf_counter()
bn.update_knowledge(df)
torch.cuda.synchronize()
print(f"update_knowledge (re-fit): {time.perf_counter() - t0:.2f} s")
rng = np.random.default_rng(1242)
vs = [int(v) for v in rng.choice(250, size=20, replace=False)]
ev = {spec.names[v]: torch.tensor(df[spec.names[v]].to_numpy()[:4096, None]) for v in vs[1:]}
t0 = time.perf_counter()
pdf, dom = bn.infer(spec.names[vs[0]], ev)
torch.cuda.synchronize()
print(f"first infer (compile + run 4096 rows): {time.perf_counter() - t0:.3f} s; rows sum {float(pdf.sum(1).mean()):.6f}")
t0 = time.perf_counter()
pdf, dom = bn.infer(spec.names[vs[0]], ev)
torch.cuda.synchronize()
print(f"second infer (cached plan): {time.perf_counter() - t0:.4f} s")
