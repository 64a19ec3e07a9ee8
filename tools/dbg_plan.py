import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CBN_COUNT_DEBUG"] = "1"
from continuousbayesiannetwork_b200 import synth
from continuousbayesiannetwork_b200.engine import tables_from_spec
for spec in (synth.alarm(), synth.random_ktree_dag()):
    t = tables_from_spec(spec, "cuda:0")
    print(t.count_groups(), t.count_updates_per_sample())
