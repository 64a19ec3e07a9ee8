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


def allreduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """In-place sum of the int64 count tables over all ranks (a no-op in a single process)."""
    assert counts.dtype == torch.int64
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
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
