#!/usr/bin/env python
"""Host<->device copy bandwidth of the box with pinned memory: one process per GPU (torchrun) or a single GPU.
Every rank copies H2D and D2H on its own GPU at the same time as the others (and both directions at once in the duplex
leg), so the aggregate is the ceiling the e2e legs of bench.py can reach on this host.
usage: python tools/probe_pcie.py            |  python -m torch.distributed.run --nproc-per-node 8 tools/probe_pcie.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
d = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=d)
mb = 256
n = mb << 20
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
g_in, g_out = torch.empty(n, dtype=torch.uint8, device=d), torch.empty(n, dtype=torch.uint8, device=d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(kind, iters=10):
    def once():
        if kind in ("h2d", "duplex"):
            with torch.cuda.stream(s1):
                g_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "duplex"):
            with torch.cuda.stream(s2):
                h_out.copy_(g_out, non_blocking=True)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([el], dtype=torch.float64, device=d)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    per_dir = n * iters / el / 1e9
    return per_dir * (2 if kind == "duplex" else 1)


res = {k: run(k) for k in ("h2d", "d2h", "duplex")}
if rank == 0:
    print(json.dumps({"n_gpus": world, "copy_MB": mb, "per_gpu_GBs": res, "aggregate_GBs": {k: v * world for k, v in res.items()},
                      "visible_cpus": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
