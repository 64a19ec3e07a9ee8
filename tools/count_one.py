#!/usr/bin/env python
"""One network's counting pass, a few launches (ncu-friendly).  usage: python tools/count_one.py ktree200|alarm|asia [log2 n] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from continuousbayesiannetwork_b200 import synth  # noqa: E402
from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ktree200"
spec = {"ktree200": synth.random_ktree_dag, "alarm": synth.alarm, "asia": synth.asia}[name]()
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
t = tables_from_spec(spec, "cuda:0")
codes = sample_network(spec, 9, 0, n, "cuda:0", tables=t)
for _ in range(2):
    t.count(codes, n)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters):
    t.count(codes, n)
b.record()
torch.cuda.synchronize()
us = a.elapsed_time(b) / iters * 1e3
print(f"count {name} n={n} upd/sample={t.count_updates_per_sample()} groups={t.count_groups()} {us:.1f} us  {n * spec.n / us / 1e3:.1f} GB/s  {n * spec.n / us / 1e3 / 6521.1:.3f} of peak")
