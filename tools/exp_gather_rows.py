#!/usr/bin/env python
"""Direct vs tile-staged gather as a function of the batch size (Alarm, 4 fused targets): picks the switch point."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network
from tools.bench_kernels import timeit
DEV = "cuda:0"; PEAK = 6521.1
for name, spec, evn, targets in (("alarm", synth.alarm(), synth.ALARM_EVIDENCE, synth.ALARM_TARGETS), ("asia", synth.asia(), ["asia", "smoke", "xray", "dysp"], ["lung", "tub", "bronc"])):
    t, inf = install_cpts(spec, DEV)
    ids = [spec.names.index(e) for e in evn]
    big = 1 << 24
    ev = sample_network(spec, 5, 0, big, DEV, tables=t)[ids].contiguous()
    fused = inf.fused_plan(targets, evn)
    outs = [torch.empty((big, 2), dtype=torch.float32, device=DEV) for _ in targets]
    for lg in (18, 19, 20, 21, 22, 23, 24):
        rows = 1 << lg
        k = [0]
        nslots = big // rows
        def f():
            s0 = (k[0] % nslots) * rows; k[0] += 1
            fused.run_codes(ev[:, s0:], rows, outs=[o[s0:] for o in outs])
        us = timeit(f, 40, 5)
        b = rows * fused.algorithmic_bytes_per_row()
        print(f"{name:6s} tiles={os.environ.get('CBN_GATHER_TILES','auto'):4s} rows=2^{lg} {us:9.2f} us  {b/us/1e3/PEAK:6.3f} of peak")
