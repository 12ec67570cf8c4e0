"""Reference schedule (ONE batch of <= 32 queries per optimizer step) through the per-head CUDA graphs: a profiling target.
usage: python scripts/profile_b1.py [steps] [plus]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rnnlogic_b200 import KnowledgeGraph

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
plus = "plus" in sys.argv
shape, N, R, train, valid, test, rules = bench.build_workload("fb15k237")
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
dev = torch.device("cuda:0")
kw = dict(type="lstm", num_layers=3, hidden_dim=16, entity_feature="bias", aggregator="sum") if plus else dict(entity_feature="bias")
out = bench.side_config("b1", kg, rules, batches, kw, 1, steps, 1, 0, dev, plus=plus)
print(out)
