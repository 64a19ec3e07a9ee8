#!/usr/bin/env python
"""Where does the Alarm gather spend its time?  Same plan, evidence distributions that change only the table locality."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from continuousbayesiannetwork_b200 import synth  # noqa: E402
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network  # noqa: E402
from tools.bench_kernels import timeit  # noqa: E402

DEV = "cuda:0"
PEAK = 6521.1
spec = synth.alarm()
rows = 1 << 24
t, inf = install_cpts(spec, DEV)
ids = [spec.names.index(e) for e in synth.ALARM_EVIDENCE]
joint = sample_network(spec, 5, 0, rows, DEV, tables=t)[ids].contiguous()
zero = torch.zeros_like(joint)
g = torch.Generator(device=DEV); g.manual_seed(1)
uni = torch.stack([torch.randint(0, spec.cards[i], (rows,), device=DEV, generator=g, dtype=torch.uint8) for i in ids])
srt = joint.clone()
key = torch.zeros(rows, dtype=torch.int64, device=DEV)
for k in range(len(ids)):
    key = key * 4 + joint[k].long()
perm = torch.argsort(key)
srt = joint[:, perm].contiguous()
fused = inf.fused_plan(synth.ALARM_TARGETS, synth.ALARM_EVIDENCE)
outs = [torch.empty((rows, 2), dtype=torch.float32, device=DEV) for _ in synth.ALARM_TARGETS]
b = rows * fused.algorithmic_bytes_per_row()
for name, ev in (("joint", joint), ("zero", zero), ("uniform", uni), ("sorted", srt)):
    us = timeit(lambda: fused.run_codes(ev, rows, outs=outs), 30)
    print(f"alarm fused x4 evidence={name:8s} {us:9.2f} us  {b / us / 1e3:8.1f} GB/s  {b / us / 1e3 / PEAK:6.3f} of peak")
single = inf.plan(synth.ALARM_TARGETS[0], synth.ALARM_EVIDENCE)
b1 = rows * single.algorithmic_bytes_per_row()
for name, ev in (("joint", joint), ("zero", zero)):
    us = timeit(lambda: single.run_codes(ev, rows, out=outs[0]), 30)
    print(f"alarm single   evidence={name:8s} {us:9.2f} us  {b1 / us / 1e3:8.1f} GB/s  {b1 / us / 1e3 / PEAK:6.3f} of peak")
# pure streaming reference of the same shape: 12 B in + 32 B out per row
a = torch.empty(rows * 12, dtype=torch.uint8, device=DEV); o = torch.empty(rows * 8, dtype=torch.float32, device=DEV)
us = timeit(lambda: (a.add_(1), o.zero_()), 30)
print(f"torch add_(12 B/row r+w) + zero_(32 B/row)  {us:9.2f} us  ({(rows * 56) / us / 1e3:8.1f} GB/s incl. the read-modify-write)")
