import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from continuousbayesiannetwork_b200 import sharding, synth  # noqa: E402
from oracle import cbn_oracle as O  # noqa: E402


def main():
    out_dir = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    spec = synth.alarm()
    n = 30011
    s, e = sharding.shard_range(n, rank, world)
    codes = synth.sample_forward_numpy(spec, 3, s, e - s)
    fams = [spec.parents[i] + [i] for i in range(spec.n)]
    local = np.concatenate([O.dense_counts(codes, f, spec.cards).reshape(-1) for f in fams])
    # tables and the local sample count travel in ONE buffer (DiscreteTables.allreduce_buffer): a single collective
    t = torch.from_numpy(np.concatenate([local, [e - s, 0]]).astype(np.int64))
    sharding.allreduce_counts(t)
    assert int(t[-2]) == n and int(t[-1]) == 0
    t = t[:-2]
    full_codes = synth.sample_forward_numpy(spec, 3, 0, n)
    want = np.concatenate([O.dense_counts(full_codes, f, spec.cards).reshape(-1) for f in fams])
    assert np.array_equal(t.numpy(), want)
    assert sharding.global_rows(e - s) == n

    # fit_sharded called TWICE on the same tables (incremental ingest): the second call must reduce only its own delta
    # (reducing the accumulated buffer again would count the first chunk world_size times).  The control flow of
    # sharding.fit_sharded is exercised with a CPU stand-in for DiscreteTables whose counting is the oracle's.
    class CpuTables:
        def __init__(self):
            self.total = int(sum(np.prod([spec.cards[v] for v in f]) for f in fams))
            self.buf = torch.zeros(self.total + 2, dtype=torch.int64)
            self.reduced = False
            self.finalized = 0

        def _dense(self, c):
            return torch.from_numpy(np.concatenate([O.dense_counts(c, f, spec.cards).reshape(-1) for f in fams]).astype(np.int64))

        def count(self, c, m):
            self.buf[: self.total] += self._dense(c)
            self.buf[self.total] += m

        def count_delta(self, c, m):
            d = torch.zeros_like(self.buf)
            d[: self.total] = self._dense(c)
            d[self.total] = m
            return d

        def add_delta(self, d):
            self.buf += d

        def allreduce_buffer(self):
            return self.buf

        def mark_reduced(self):
            self.reduced = True

        def is_reduced(self):
            return self.reduced

        def finalize(self):
            self.finalized += 1

    tb = CpuTables()
    n2 = 20011
    s2, e2 = sharding.shard_range(n2, rank, world)
    codes2 = synth.sample_forward_numpy(spec, 5, s2, e2 - s2)
    sharding.fit_sharded(tb, codes, e - s)
    sharding.fit_sharded(tb, codes2, e2 - s2)
    full2 = synth.sample_forward_numpy(spec, 5, 0, n2)
    want2 = want + np.concatenate([O.dense_counts(full2, f, spec.cards).reshape(-1) for f in fams])
    assert np.array_equal(tb.buf[: tb.total].numpy(), want2), "second sharded call double-counted the first"
    assert int(tb.buf[tb.total]) == n + n2 and tb.finalized == 2
    # chunked ingest without communication until the end
    tc = CpuTables()
    sharding.count_local(tc, codes, e - s)
    sharding.count_local(tc, codes2, e2 - s2)
    sharding.reduce_and_finalize(tc)
    assert np.array_equal(tc.buf[: tc.total].numpy(), want2) and int(tc.buf[tc.total]) == n + n2
    open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
