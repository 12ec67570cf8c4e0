"""Side benchmark: PredictorPlus fused train steps on the FB15k-237-shape workload (profiling target)."""
import sys, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from rnnlogic_b200 import KnowledgeGraph
from rnnlogic_b200.predictors import PredictorPlus
variant = sys.argv[1] if len(sys.argv) > 1 else "lstm"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shape, N, R, train, valid, test, rules = bench.build_workload()
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
kw = dict(type="lstm", num_layers=3, hidden_dim=16, entity_feature="bias", aggregator="sum") if variant == "lstm" else \
     dict(type="emb", hidden_dim=16, entity_feature="bias", aggregator="pna")
torch.manual_seed(0)
pm = PredictorPlus(kg, **kw); pm.set_rules([[h] + list(b) for h, b in rules]); pm = pm.cuda()
pm.fused_rnn = "--cudnn" not in sys.argv
opt = torch.optim.Adam(pm.parameters(), lr=0.005)
per = 64
def step(i):
    sb = [batches[(i * per + j) % len(batches)] for j in range(per)]
    opt.zero_grad(set_to_none=True)
    loss, _ = pm.fused_train_step(sb, 0.2, grad_scale=1.0 / per)
    opt.step()
    return sum(len(b) for b in sb)
for i in range(3): step(i)
torch.cuda.synchronize(); t0 = time.perf_counter(); q = 0
for i in range(3, 3 + steps): q += step(i)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("plus", variant, "q/s", q / dt, "ms/step", 1e3 * dt / steps)
