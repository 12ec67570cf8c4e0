"""Diagnostic: how sparse are the frontiers of the bench workload (per trie depth)?"""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from rnnlogic_b200 import KnowledgeGraph
from rnnlogic_b200.predictors import Predictor
shape, N, R, train, valid, test, rules = bench.build_workload(typed='--typed' in sys.argv)
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
m = Predictor(kg, "bias"); m.set_rules([[h] + list(b) for h, b in rules]); m = m.cuda()
sk = m._driver(torch.device("cuda:0")); cr = m.compiled
sl = sk.gr.make_slots_host(batches[:64], with_etr=True)
sl.use_workspace = False
sk.gr.ground(sl)
state = sl.state.cpu().numpy()
n_mask, n_cnt = sl.mask_words + 1, sl.nz_total + 1
masks = state[:n_mask].view(np.uint32); cnt = state[n_mask:n_mask + n_cnt]
pop = np.bitwise_count(masks)
tot_nodes = np.zeros(4); nz_nodes = np.zeros(4); rows = np.zeros(4); vrows = np.zeros(4); chunks = np.zeros(4); nzchunks = np.zeros(4); maxrows=np.zeros(4)
moff = 0; noff = 0
for s, q in enumerate(sl.heads):
    n0, n1 = cr.head_node_ptr[q], cr.head_node_ptr[q + 1]
    dep = cr.node_depth[n0:n1]; nr = kg.rel_rows[cr.node_rel_host[n0:n1]]
    c = cnt[noff:noff + (n1 - n0)]
    nch = (nr + 31) // 32
    cs = np.concatenate([[0], np.cumsum(nch)])
    pm = pop[moff:moff + cs[-1]]
    for d in (1, 2, 3):
        sel = dep == d
        tot_nodes[d] += sel.sum(); nz_nodes[d] += (c[sel] > 0).sum(); rows[d] += nr[sel].sum(); vrows[d] += c[sel].sum()
        chunks[d] += nch[sel].sum()
        for i in np.flatnonzero(sel):
            nzchunks[d] += (pm[cs[i]:cs[i + 1]] > 0).sum()
        maxrows[d]=max(maxrows[d], c[sel].max() if sel.any() else 0)
    moff += cs[-1]; noff += n1 - n0
for d in (1, 2, 3):
    print("  depth %d: valid rows per non-empty node %.2f, per non-empty chunk %.2f" % (d, vrows[d] / max(1, nz_nodes[d]), vrows[d] / max(1, nzchunks[d])))
    print("depth %d: nodes %d nonzero %d | rows %d valid %d (%.3f%%) max/node %d | chunks %d nonzero %d (%.2f%%)" % (
        d, tot_nodes[d], nz_nodes[d], rows[d], vrows[d], 100 * vrows[d] / max(1, rows[d]), maxrows[d], chunks[d], nzchunks[d], 100 * nzchunks[d] / max(1, chunks[d])))
