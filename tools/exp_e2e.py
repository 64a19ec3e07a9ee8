#!/usr/bin/env python
"""e2e host-buffer call: staged (CBN_HOST_DIRECT=0) vs direct (1) on the Asia and Alarm workloads."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network
dev = "cuda:0"
for name, spec, evn, targets, rows in (("asia", synth.asia(), ["asia", "smoke", "xray", "dysp"], ["lung", "tub", "bronc"], 1 << 20),
                                       ("alarm", synth.alarm(), synth.ALARM_EVIDENCE, synth.ALARM_TARGETS, 1 << 22)):
    t, inf = install_cpts(spec, dev)
    ids = [spec.names.index(e) for e in evn]
    ev = sample_network(spec, 5, 0, rows, dev, tables=t)[ids].contiguous()
    fused = inf.fused_plan(targets, evn)
    want = [o.cpu() for o in fused.run_codes(ev, rows)]
    hev = ev.cpu().pin_memory()
    houts = [torch.empty((rows, 2), dtype=torch.float32).pin_memory() for _ in targets]
    for _ in range(3):
        fused.run_codes_host(hev, rows, houts)
    assert all(torch.equal(a, b) for a, b in zip(want, houts))
    t0 = time.perf_counter(); k = 30
    for _ in range(k):
        fused.run_codes_host(hev, rows, houts)
    el = (time.perf_counter() - t0) / k
    byts = rows * (len(ids) + 8 * len(targets))
    print(f"{name:6s} direct={os.environ.get('CBN_HOST_DIRECT','1')} {el*1e6:9.1f} us/call  {rows*len(targets)/el/1e9:7.3f} G queries/s  {byts/el/1e9:6.1f} GB/s over PCIe (both directions)")
