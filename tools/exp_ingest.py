#!/usr/bin/env python
"""Ingestion throughput: float32 columns -> domains (cbn_domain_f32) -> uint8 codes (cbn_encode_f32) -> counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
from continuousbayesiannetwork_b200.tables import DiscreteTables
from tools.bench_kernels import timeit
DEV = "cuda:0"; PEAK = 6521.1
for name, spec, n in (("asia", synth.asia(), 1 << 26), ("alarm", synth.alarm(), 1 << 24)):
    t0 = tables_from_spec(spec, DEV)
    codes = sample_network(spec, 3, 0, n, DEV, tables=t0)
    cols = {nm: (codes[i, :n].to(torch.float32) * 0.5 - 1.0) for i, nm in enumerate(spec.names)}
    del codes
    t = DiscreteTables(spec.names, spec.parents_by_name(), device=DEV)
    col0 = cols[spec.names[0]]
    us = timeit(lambda: t.discover_domain(col0), 10, 2)
    print(f"{name:6s} domain  n={n} {us:9.1f} us/column  {4*n/us/1e3:8.1f} GB/s  {4*n/us/1e3/PEAK:5.3f} of peak")
    t.set_domains([t.discover_domain(cols[nm]) for nm in spec.names])
    out = t.new_code_matrix(n)
    us = timeit(lambda: t.encode(col0, 0, out[0]), 10, 2)
    print(f"{name:6s} encode  n={n} {us:9.1f} us/column  {5*n/us/1e3:8.1f} GB/s  {5*n/us/1e3/PEAK:5.3f} of peak")
    us = timeit(lambda: t.fit_columns(cols), 3, 1)
    b = n * spec.n * (4 + 4 + 1 + 1)   # domain read + encode read + code write + count read
    print(f"{name:6s} fit_columns (domains + encode + count + CPTs) n={n} x {spec.n} cols: {us:9.1f} us  {n/us/1e3:7.2f} G samples/s  {b/us/1e3:8.1f} GB/s")
