#!/usr/bin/env python
"""Float32 ingestion (domains -> codes -> counts) on the Asia shape, part by part."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec

dev = "cuda:0"
spec = synth.asia()
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
t0 = tables_from_spec(spec, dev)
t0.set_cond_tables(spec.cpts)
c = sample_network(spec, seed=99, first=0, n=n, device=dev, tables=t0)
cols = {nm: (c[i, :n].to(torch.float32) * 0.5 - 1.0) for i, nm in enumerate(spec.names)}
del c
ing = tables_from_spec(spec, dev)

def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / iters

vals = n * spec.n
flat = [cols[nm].reshape(-1) for nm in ing.names]
s = timed(lambda: ing.discover_domains(flat))
print(f"discover_domains: {s * 1e3:7.3f} ms  {vals * 4 / s / 1e9:7.1f} GB/s (4 B per value)")
ing.fit_columns(cols)
s = timed(lambda: ing.encode_columns(cols, strict=False))
print(f"encode_columns:   {s * 1e3:7.3f} ms  {vals * 5 / s / 1e9:7.1f} GB/s (4 + 1 B per value)")
codes = ing.encode_columns(cols, strict=False)
s = timed(lambda: ing.count(codes, n))
print(f"count:            {s * 1e3:7.3f} ms  {vals / s / 1e9:7.1f} GB/s (1 B per value)")
s = timed(lambda: ing.fit_columns(cols))
print(f"fit_columns:      {s * 1e3:7.3f} ms  {vals * 10 / s / 1e9:7.1f} GB/s (10 B per value)")
