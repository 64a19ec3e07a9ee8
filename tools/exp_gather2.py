#!/usr/bin/env python
"""Alarm gather, joint evidence only (one process per variant: knobs are read from the environment)."""
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
fused = inf.fused_plan(synth.ALARM_TARGETS, synth.ALARM_EVIDENCE)
outs = [torch.empty((rows, 2), dtype=torch.float32, device=DEV) for _ in synth.ALARM_TARGETS]
b = rows * fused.algorithmic_bytes_per_row()
us = timeit(lambda: fused.run_codes(joint, rows, outs=outs), 30)
single = inf.plan(synth.ALARM_TARGETS[0], synth.ALARM_EVIDENCE)
us1 = timeit(lambda: single.run_codes(joint, rows, out=outs[0]), 30)
tag = " ".join(f"{k[4:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("CBN_G"))
print(f"{tag:50s} fused {us:8.2f} us {b / us / 1e3 / PEAK:6.3f}   single {us1:8.2f} us {rows * 20 / us1 / 1e3 / PEAK:6.3f}")
