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
    open(os.path.join(out_dir, f"ok_{rank}"), "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
