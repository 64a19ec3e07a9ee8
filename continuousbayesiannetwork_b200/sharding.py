"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPUs, gloo in the CPU tests).

The path shards two ways and needs exactly one collective:

* queries: rows are independent -> contiguous row ranges per rank, plan and CPTs replicated,
  no data-path collective;
* CPT fit: samples sharded by contiguous ranges, every rank builds full private int64 tables,
  ONE all-reduce (sum, int64) of the concatenated tables; integer addition is associative, so
  the result is bit-identical for any number of ranks.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n: int, rank: int, world_size: int, align: int = 16) -> Tuple[int, int]:
    """Contiguous [start, end) of rank ``rank``; starts are multiples of ``align`` so every shard of a
    code matrix keeps the 16-byte alignment the kernels' vector loads need."""
    per = (n + world_size - 1) // world_size
    per = (per + align - 1) // align * align
    s = min(n, rank * per)
    e = min(n, s + per) if rank < world_size - 1 else n
    return s, max(s, e)


_library_comms = {}     # device index -> cbn_comm handle (the library's own NCCL communicator)


class _StdoutToStderr:
    """NCCL may print its version banner on stdout when it is first initialised in a process; a caller that prints
    machine-readable output there (bench.py's JSON line) must not see it."""

    def __enter__(self):
        import os
        import sys

        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        import os

        os.dup2(self.saved, 1)
        os.close(self.saved)


def library_comm(device):
    """The library's communicator for ``device`` (``cbn_comm_create``), bootstrapped through the torch.distributed group:
    only the 128-byte NCCL id travels through torch; the collective itself is the C ABI's ``cbn_counts_allreduce``.
    Returns None when NCCL cannot be bound (``CBN_LIBRARY_COMM=0`` also disables it): callers then use torch.distributed."""
    import ctypes as C
    import os

    from . import _native as N

    if os.environ.get("CBN_LIBRARY_COMM", "1") == "0" or not (dist.is_available() and dist.is_initialized()):
        return None
    dev = torch.device(device)
    if dev.type != "cuda":
        return None
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _library_comms:
        return _library_comms[idx]
    rank, world_size = dist.get_rank(), dist.get_world_size()
    ident = torch.zeros(N.COMM_ID_BYTES + 1, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * N.COMM_ID_BYTES)()
        with _StdoutToStderr():
            ok = N.lib().cbn_comm_unique_id(buf) == N.OK
        ident[: N.COMM_ID_BYTES] = torch.tensor(list(buf), dtype=torch.uint8)
        ident[N.COMM_ID_BYTES] = 1 if ok else 0
    on_dev = dist.get_backend() == "nccl"
    t = ident.to(dev) if on_dev else ident
    dist.broadcast(t, 0)
    ident = t.cpu()
    handle = None
    if int(ident[N.COMM_ID_BYTES]) == 1:
        ctx = N.context_for(dev)
        raw = (C.c_uint8 * N.COMM_ID_BYTES)(*ident[: N.COMM_ID_BYTES].tolist())
        h = C.c_void_p()
        with _StdoutToStderr():
            rc = N.lib().cbn_comm_create(ctx.handle, raw, world_size, rank, C.byref(h))
        N.check(rc, ctx.handle)
        handle = (ctx, h)
    _library_comms[idx] = handle
    return handle


def allreduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """In-place sum of the int64 count tables over all ranks (a no-op in a single process).  Device tables go through
    the library's own collective (``cbn_counts_allreduce`` = ncclAllReduce(int64, sum) on the current stream); host
    tensors (the gloo CPU tests) and processes without NCCL through torch.distributed."""
    assert counts.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        comm = library_comm(counts.device) if counts.is_cuda else None
        if comm is not None and counts.is_contiguous():
            from . import _native as N

            ctx, h = comm
            N.check(N.lib().cbn_counts_allreduce(ctx.handle, h, counts.data_ptr(), counts.numel(), N.stream_ptr(counts.device)), ctx.handle)
        else:
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def global_rows(local_rows: int, device=None) -> int:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(local_rows)
    t = torch.tensor([int(local_rows)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def fit_sharded(tables, codes: torch.Tensor, n_local: int):
    """Count the local shard, all-reduce the tables, normalise with the GLOBAL sample count.

    The first call on fresh tables reduces them in place.  After that the tables already hold the global counts, so a
    further call (incremental / chunked ingest) counts its shard into a zeroed delta buffer, reduces the DELTA and adds
    it: reducing the accumulated buffer again would multiply the earlier counts by the world size."""
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    if multi and tables.is_reduced():
        delta = tables.count_delta(codes, n_local)
        allreduce_counts(delta)
        tables.add_delta(delta)
    else:
        tables.count(codes, n_local)
        if multi:
            # tables and sample count travel together: one collective, no host synchronisation
            allreduce_counts(tables.allreduce_buffer())
            tables.mark_reduced()
    tables.finalize()
    return tables


def count_local(tables, codes: torch.Tensor, n_local: int):
    """Chunked ingest, phase 1: accumulate a local chunk WITHOUT communicating (call ``reduce_and_finalize`` once at the
    end).  Must not be mixed with ``fit_sharded`` on tables that are already reduced."""
    if tables.is_reduced():
        raise RuntimeError("tables already hold global counts: use fit_sharded (delta reduction) for further chunks")
    tables.count(codes, n_local)


def reduce_and_finalize(tables):
    """Chunked ingest, phase 2: ONE all-reduce of everything counted locally so far, then the CPTs."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if tables.is_reduced():
            raise RuntimeError("tables were already reduced")
        allreduce_counts(tables.allreduce_buffer())
        tables.mark_reduced()
    tables.finalize()
    return tables
