"""Diagnostic: how many entities / (entity, query) cells of a bench step carry a rule-grounded score?"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from rnnlogic_b200 import KnowledgeGraph
from rnnlogic_b200.predictors import Predictor
shape, N, R, train, valid, test, rules = bench.build_workload()
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
m = Predictor(kg, "bias"); m.set_rules([[h] + list(b) for h, b in rules]); m = m.cuda()
sk = m._driver(torch.device("cuda:0"))
sl = sk.gr.make_slots_host(batches[:64], with_etr=True)
sk.gr.ground(sl)
Z, nz = sk.predictor_scores(sl, m.rule_weights.detach(), m.bias.detach(), False)
nz = nz.cpu().numpy().view(np.uint32)
ent = (nz != 0).sum(1)
cells = np.bitwise_count(nz).sum(1)
nq = np.array([len(b[0]) if hasattr(b[0], '__len__') else 32 for b in batches[:64]])
print("slots %d N %d | active entities/slot: mean %.0f (%.1f%% of N) min %d max %d | active cells/slot mean %.0f (%.2f%% of N*32)" % (
    sl.S, N, ent.mean(), 100 * ent.mean() / N, ent.min(), ent.max(), cells.mean(), 100 * cells.mean() / (N * 32)))
print("items/slot mean %.0f" % (sl.item_cnt.cpu().numpy()[:sl.S].mean() if hasattr(sl, 'item_cnt') else -1))
