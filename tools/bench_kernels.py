#!/usr/bin/env python
"""Quick per-kernel timings (CUDA events) used while tuning; not the contract bench.
usage: python tools/bench_kernels.py [gather] [count] [--ncu-friendly]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from continuousbayesiannetwork_b200 import synth  # noqa: E402
from continuousbayesiannetwork_b200.engine import install_cpts, sample_network, tables_from_spec  # noqa: E402

DEV = "cuda:0"
PEAK = 6521.1
short = "--ncu-friendly" in sys.argv


def timeit(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def flush_l2(buf=[None]):
    if buf[0] is None:
        buf[0] = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    buf[0].add_(1)


def gather():
    for name, spec, evn, targets, rows in (("asia", synth.asia(), ["asia", "smoke", "xray", "dysp"], ["lung", "tub", "bronc"], 1 << 20),
                                           ("alarm", synth.alarm(), synth.ALARM_EVIDENCE, synth.ALARM_TARGETS, 1 << 24)):
        t, inf = install_cpts(spec, DEV)
        ids = [spec.names.index(e) for e in evn]
        ring = 8 if name == "asia" else 1
        evs = [sample_network(spec, 5, r * rows, rows, DEV, tables=t)[ids].contiguous() for r in range(ring)]
        fused = inf.fused_plan(targets, evn)
        outs = [[torch.empty((rows, 2), dtype=torch.float32, device=DEV) for _ in targets] for _ in range(ring)]
        single = inf.plan(targets[0], evn)
        k = [0]

        def f_fused():
            r = k[0] % ring; k[0] += 1
            fused.run_codes(evs[r], rows, outs=outs[r])

        def f_single():
            r = k[0] % ring; k[0] += 1
            single.run_codes(evs[r], rows, out=outs[r][0])

        it = 3 if short else 100
        us = timeit(f_fused, it)
        b = rows * fused.algorithmic_bytes_per_row()
        print(f"gather {name:6s} fused x{len(targets)}  {us:9.2f} us/launch  {b / us / 1e3:8.1f} GB/s  {b / us / 1e3 / PEAK:6.3f} of peak   ({rows * len(targets) / us / 1e3:.1f} G queries/s)")
        us = timeit(f_single, it)
        b = rows * single.algorithmic_bytes_per_row()
        print(f"gather {name:6s} single     {us:9.2f} us/launch  {b / us / 1e3:8.1f} GB/s  {b / us / 1e3 / PEAK:6.3f} of peak")
        if not short:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for r in range(ring):
                    fused.run_codes(evs[r], rows, outs=outs[r])
            us = timeit(g.replay, 30) / ring
            b = rows * fused.algorithmic_bytes_per_row()
            print(f"gather {name:6s} fused graph {us:8.2f} us/launch  {b / us / 1e3:8.1f} GB/s  {b / us / 1e3 / PEAK:6.3f} of peak")


def count():
    for name, spec, n in (("asia", synth.asia(), 1 << 26), ("alarm", synth.alarm(), 1 << 25), ("ktree200", synth.random_ktree_dag(), 1 << 24)):
        t = tables_from_spec(spec, DEV)
        codes = sample_network(spec, 9, 0, n, DEV, tables=t)
        torch.cuda.synchronize()

        def f():
            t.count(codes, n)

        us = timeit(f, 3 if short else 10, 2)
        b = n * spec.n
        print(f"count  {name:9s} n={n:>10d} vars={spec.n:4d} groups={t.count_groups():3d} upd/sample={t.count_updates_per_sample():3d}  {us:10.1f} us  {n / us / 1e3:8.2f} G samples/s  {b / us / 1e3:8.1f} GB/s  {b / us / 1e3 / PEAK:6.3f} of peak")


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["gather", "count"]
    if "gather" in which:
        gather()
    if "count" in which:
        count()
