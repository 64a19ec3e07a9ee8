import sys, time, torch
sys.path.insert(0, "/root/repo")
from continuousbayesiannetwork_b200 import synth, sharding
from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
spec = synth.asia()
t = tables_from_spec(spec, "cuda:0")
for n in (10_000_000, 10_000_000 - 1664, 1 << 23, 2048*100+5):
    codes = sample_network(spec, 1, 0, n, "cuda:0", tables=t)
    torch.cuda.synchronize()
    def tm(fn, it=10):
        fn(); torch.cuda.synchronize()
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(it): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b)/it*1e3
    print(n, "count %.1f us" % tm(lambda: t.count(codes, n)), "finalize %.1f us" % tm(lambda: t.finalize()), "zero %.1f us" % tm(lambda: t.counts.zero_()),
          "fit_sharded %.1f us" % tm(lambda: (t.counts.zero_(), setattr(t, 'n_total', 0), sharding.fit_sharded(t, codes, n))))
