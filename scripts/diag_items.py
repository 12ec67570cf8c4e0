"""Diagnostic: distribution of the item-list work per (slot, entity word) warp of k_predictor_scores."""
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
sk.predictor_scores(sl, m.rule_weights.detach(), m.bias.detach(), False)
torch.cuda.synchronize()
W = kg.rank_words
NB = W * 32 + 32
n_items = 4 * max(1, sl.item_cap)
off = sl.scratch[2 * n_items:2 * n_items + sl.S * NB].cpu().numpy().reshape(sl.S, NB)
per_ent = np.diff(off[:, :W * 32 + 1], axis=1)                      # [S, W*32]
per_word = per_ent.reshape(sl.S, W, 32).sum(2)
groups = ((per_ent + 3) // 4).reshape(sl.S, W, 32).sum(2)          # 4-row groups a warp issues serially
print("items/slot: mean %.0f max %d" % (per_ent.sum(1).mean(), per_ent.sum(1).max()))
print("items/entity: max %d; entities with items %.1f%%" % (per_ent.max(), 100 * (per_ent > 0).mean()))
print("items/word: mean %.1f p50 %d p90 %d p99 %d max %d" % (per_word.mean(), *np.percentile(per_word, [50, 90, 99]).astype(int), per_word.max()))
print("serial groups/warp: mean %.1f p90 %d p99 %d max %d" % (groups.mean(), *np.percentile(groups, [90, 99]).astype(int), groups.max()))
blk = groups.reshape(sl.S, -1)[:, :W // 8 * 8].reshape(sl.S, W // 8, 8)
print("block max groups: mean %.1f max %d; sum over blocks of max / 888 resident = %.0f group-times" % (blk.max(2).mean(), blk.max(2).max(), blk.max(2).sum() / 888))
