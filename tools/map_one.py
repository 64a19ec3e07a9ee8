#!/usr/bin/env python
"""The fused MAP kernel (posterior + argmax + domain value) on the Alarm workload, alone, for ncu / timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network

rows = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = "cuda:0"
spec = synth.alarm()
tables, infer = install_cpts(spec, dev)
ids = [spec.names.index(e) for e in synth.ALARM_EVIDENCE]
full = sample_network(spec, seed=1236, first=0, n=rows, device=dev, tables=tables)
ev = full[ids].contiguous()
del full
if os.environ.get("MAP_UNIFORM"):      # evidence drawn uniformly instead of from the network: no hot configurations
    g = torch.Generator(device=dev).manual_seed(5)
    for j, i in enumerate(ids):
        ev[j] = torch.randint(0, spec.cards[i], (ev.shape[1],), generator=g, device=dev, dtype=torch.uint8)
if os.environ.get("MAP_SORTED"):       # rows sorted by configuration: neighbouring rows share table sectors
    key = torch.zeros(ev.shape[1], dtype=torch.int64, device=dev)
    for j, i in enumerate(ids):
        key = key * spec.cards[i] + ev[j].long()
    ev = ev[:, torch.argsort(key)].contiguous()
plan = infer.plan(synth.ALARM_TARGETS[0], synth.ALARM_EVIDENCE)
plan.set_static_evidence(True)
out = torch.empty(rows, dtype=torch.float32, device=dev)
for _ in range(3):
    plan.run_codes_map(ev, rows, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.run_codes_map(ev, rows, out=out)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
nb = rows * (len(plan.stats.relevant_evidence) + 4)
print(f"MAP {rows} rows  {us:8.1f} us  {rows / us / 1e3:7.2f} G rows/s  {nb / us / 1e3:7.1f} GB/s algorithmic  table cells {plan.stats.final_tables}")
