"""Run a few end-to-end fused train steps of ONE configuration (profiling target for ncu).
usage: python scripts/profile_config.py <fb15k237|wn18rr> <predictor|plus_lstm_sum|plus_emb_pna|plus_rotate> [batches] [steps] [--typed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rnnlogic_b200 import KnowledgeGraph

shape_name, kind = sys.argv[1], sys.argv[2]
per = int(sys.argv[3]) if len(sys.argv) > 3 else 64
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
typed = "--typed" in sys.argv
shape, N, R, train, valid, test, rules = bench.build_workload(shape_name, typed=typed)
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
dev = torch.device("cuda:0")
kw = {"predictor": dict(entity_feature="bias"),
      "plus_lstm_sum": dict(type="lstm", num_layers=3, hidden_dim=16, entity_feature="bias", aggregator="sum"),
      "plus_emb_pna": dict(type="emb", hidden_dim=16, entity_feature="bias", aggregator="pna"),
      "plus_rotate": dict(type="lstm", num_layers=3, hidden_dim=16, entity_feature="RotatE", aggregator="sum")}[kind]
if kind == "plus_rotate":
    kw["embedding_path"] = bench.rotate_dir(N, R)
out = bench.side_config(kind, kg, rules, batches, kw, per, steps, 1, 0, dev, plus=kind != "predictor")
print(out)
