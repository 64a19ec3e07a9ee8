import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
from continuousbayesiannetwork_b200.tables import DiscreteTables
DEV = "cuda:0"
spec = synth.alarm(); n = 1 << 24
t0 = tables_from_spec(spec, DEV)
codes = sample_network(spec, 3, 0, n, DEV, tables=t0)
cols = {nm: (codes[i, :n].to(torch.float32) * 0.5 - 1.0) for i, nm in enumerate(spec.names)}
t = DiscreteTables(spec.names, spec.parents_by_name(), device=DEV)
t.fit_columns(cols); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t.fit_columns(cols); torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
