"""Time items sort + k_predictor_scores (with the fused softmax partials) for several step sizes."""
import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from rnnlogic_b200 import KnowledgeGraph
from rnnlogic_b200.predictors import Predictor
shape, N, R, train, valid, test, rules = bench.build_workload()
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
m = Predictor(kg, "bias"); m.set_rules([[h] + list(b) for h, b in rules]); m = m.cuda()
sk = m._driver(torch.device("cuda:0"))
w, b = m.rule_weights.detach(), m.bias.detach()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for nb in [int(a) for a in sys.argv[1:]] or [64]:
    sl = sk.gr.make_slots_host(batches[:nb], with_etr=True)
    sk.gr.ground(sl)
    partial = torch.empty(sl.S * sk.nblk * 64, dtype=torch.float32, device="cuda")
    us = t(lambda: sk.predictor_scores(sl, w, b, False, partial))
    print("batches %4d: arena %.2f GB, items %d: sort+scores+partial %.1f us = %.2f us/batch" % (
        nb, sl.arena.numel() * 4 / 1e9, sl.item_cap, us, us / nb))
