"""Multi-GPU check (run under torchrun with 2 ranks, or alone for the 1-GPU reference):
TrainerPredictor at world size 2 (one batch per rank per step, flat gradient all-reduce, DDP-mean) must give
the same parameters and MRR as ONE process that takes the same two batches per step.
    python scripts/check_ddp_equivalence.py single /tmp/out1.pt
    torchrun --nproc-per-node 2 ... scripts/check_ddp_equivalence.py ddp /tmp/out2.pt ; then `compare`"""
import os, sys, random
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _golden as G
from rnnlogic_b200.data import KnowledgeGraph, TrainDataset, ValidDataset, TestDataset
from rnnlogic_b200.predictors import Predictor, PredictorPlus
from rnnlogic_b200.trainer import TrainerPredictor
from rnnlogic_b200.utils import set_seed
from rnnlogic_b200 import comm

mode, out = sys.argv[1], sys.argv[2]
if mode == "compare":
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    for k in a["sd"]:
        err = (a["sd"][k] - b["sd"][k]).abs().max().item()
        scale = a["sd"][k].abs().max().item()
        print("%-40s max|diff| %.3e (scale %.3e)" % (k, err, scale))
        assert err <= 1e-5 * max(scale, 1.0), k
    print("mrr", a["mrr"], b["mrr"])
    assert abs(a["mrr"] - b["mrr"]) <= 1e-6 * max(abs(a["mrr"]), 1e-12) + 1e-9
    print("DDP equivalence OK")
    sys.exit(0)

fx = G.load("umls")
set_seed(1)
kg = KnowledgeGraph(entity_size=int(fx["N"]), relation_size=int(fx["R"]), train=fx["train"], valid=fx["valid"], test=fx["test"])
tr, va, te = TrainDataset(kg, 32), ValidDataset(kg, 32), TestDataset(kg, 32)
torch.manual_seed(0)
model = PredictorPlus(kg, type="emb", hidden_dim=16, entity_feature="bias", aggregator="sum")
model.set_rules(G.rules_of(fx))
opt = torch.optim.Adam(model.parameters(), lr=0.005)
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
solver = TrainerPredictor(model, tr, va, te, opt, gpus=list(range(world)))
if mode == "single":
    solver.slots_per_step = 2
    solver.train(batch_per_epoch=24, smoothing=0.2, print_every=1000)
else:
    solver.train(batch_per_epoch=12, smoothing=0.2, print_every=1000)
mrr = solver.evaluate("valid", expectation=True)
if rank == 0:
    torch.save({"sd": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "mrr": float(mrr)}, out)
comm.synchronize()
