"""torchrun worker of tests/test_multigpu_gpu.py: one process per GPU, NCCL.  Sharded CPT fit (first call reduces in
place, second call reduces its delta) against a single-GPU recount of the concatenated shards and against the CPU oracle;
row-sharded queries against the unsharded run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from continuousbayesiannetwork_b200 import sharding, synth  # noqa: E402
from continuousbayesiannetwork_b200.engine import bind_inference, sample_network, tables_from_spec  # noqa: E402
from oracle import cbn_oracle as O  # noqa: E402


def main():
    out_dir = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    # the library's own collective (cbn_counts_allreduce) against torch.distributed's on the same buffer
    from continuousbayesiannetwork_b200 import _native as N

    comm = sharding.library_comm(dev)
    assert comm is not None, "cbn_comm could not be created (NCCL not loadable?)"
    assert N.lib().cbn_comm_size(comm[1]) == world
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    a = torch.randint(0, 1 << 40, (100_003,), dtype=torch.int64, device=dev, generator=g)
    b = a.clone()
    sharding.allreduce_counts(a)                       # library path
    dist.all_reduce(b, op=dist.ReduceOp.SUM)           # torch path
    assert torch.equal(a, b), "cbn_counts_allreduce differs from torch.distributed.all_reduce"
    for spec, n1, n2 in ((synth.alarm(), 3_000_017, 1_000_003), (synth.random_ktree_dag(), 400_009, 65_537)):
        t = tables_from_spec(spec, dev)
        s1, e1 = sharding.shard_range(n1, rank, world)
        s2, e2 = sharding.shard_range(n2, rank, world)
        c1 = sample_network(spec, seed=71, first=s1, n=e1 - s1, device=dev, tables=t)
        c2 = sample_network(spec, seed=72, first=s2, n=e2 - s2, device=dev, tables=t)
        sharding.fit_sharded(t, c1, e1 - s1)          # count_tiles_kernel + ncclAllReduce(int64) in place
        assert t.is_reduced() and t.n_total == n1
        sharding.fit_sharded(t, c2, e2 - s2)          # second call: delta buffer, reduce the delta, add
        assert t.n_total == n1 + n2
        # every rank recounts the CONCATENATED shards alone (no collective) and must hold identical tables
        alone = tables_from_spec(spec, dev)
        f1 = sample_network(spec, seed=71, first=0, n=n1, device=dev, tables=alone)
        f2 = sample_network(spec, seed=72, first=0, n=n2, device=dev, tables=alone)
        alone.count(f1, n1)
        alone.count(f2, n2)
        alone.finalize()
        assert torch.equal(t.counts, alone.counts), f"{spec.n}-node: all-reduced tables differ from the single-GPU recount"
        assert torch.equal(t.cond, alone.cond) and torch.equal(t.joint, alone.joint)
        # and rank 0 checks a few families against the CPU oracle on the same samples
        if rank == 0:
            full = np.concatenate([f1[:, :n1].cpu().numpy(), f2[:, :n2].cpu().numpy()], axis=1)
            for i in list(range(0, spec.n, max(1, spec.n // 12))):
                want = O.dense_counts(full, spec.parents[i] + [i], spec.cards)
                got = t.table_view(t.counts, spec.names[i]).cpu().numpy()
                assert np.array_equal(got, want), spec.names[i]
        del c1, c2, f1, f2
    # queries: contiguous row ranges per rank, plan replicated, no collective on the data path
    spec = synth.alarm()
    t = tables_from_spec(spec, dev)
    t.set_cond_tables(spec.cpts)
    infer = bind_inference(t)
    n = 1_000_003
    ids = [spec.names.index(e) for e in synth.ALARM_EVIDENCE]
    s, e = sharding.shard_range(n, rank, world)
    ev = sample_network(spec, seed=73, first=s, n=e - s, device=dev, tables=t)[ids].contiguous()
    fused = infer.fused_plan(synth.ALARM_TARGETS, synth.ALARM_EVIDENCE)
    mine = fused.run_codes(ev, e - s)
    ev_full = sample_network(spec, seed=73, first=0, n=n, device=dev, tables=t)[ids].contiguous()
    full = fused.run_codes(ev_full, n)
    for a, b in zip(mine, full):
        assert torch.equal(a, b[s:e]), "row-sharded posteriors differ from the unsharded run"
    assert sharding.global_rows(e - s, device=dev) == n
    torch.cuda.synchronize()
    open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
