"""Host side of the pipelined end-to-end train loop (submit_train_step / ticket.result()): per-step wall
times, the enqueue-only cost and a cProfile of it.  usage: python scripts/profile_host.py [batches_per_step]"""
import cProfile, pstats, sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from rnnlogic_b200 import KnowledgeGraph
from rnnlogic_b200.predictors import Predictor
from rnnlogic_b200.optim import Adam
per = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape, N, R, train, valid, test, rules = bench.build_workload()
batches = bench.make_batches(train, R, seed=1)
kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
m = Predictor(kg, "bias"); m.set_rules([[h] + list(b) for h, b in rules]); m = m.cuda()
opt = Adam(m.parameters(), lr=0.005)
steps = 40
lists = [[batches[(i * per + j) % len(batches)] for j in range(per)] for i in range(steps + 4)]
def step(gw, gb):
    m.rule_weights.grad, m.bias.grad = gw, gb
    opt.step()
def loop(n0, n1, times=None):
    ticket = m.submit_train_step(lists[n0], 0.2, grad_scale=1.0 / per)
    for s in range(n0, n1):
        t0 = time.perf_counter()
        step(ticket.gw, ticket.gb)
        nxt = m.submit_train_step(lists[s + 1], 0.2, grad_scale=1.0 / per)
        t1 = time.perf_counter()
        ticket.result()
        if times is not None:
            times.append((t1 - t0, time.perf_counter() - t1))
        ticket = nxt
    torch.cuda.synchronize()
loop(0, 3)
times = []
t0 = time.perf_counter(); loop(3, steps, times); t1 = time.perf_counter()
print("wall per step %.3f ms" % ((t1 - t0) / (steps - 3) * 1e3))
print("per step: enqueue ms / wait ms:", " ".join("%.2f/%.2f" % (a * 1e3, b * 1e3) for a, b in times))
def submit_only(n0, n1):
    for s in range(n0, n1):
        t = m.submit_train_step(lists[s], 0.2, grad_scale=1.0 / per)
        step(t.gw, t.gb)
torch.cuda.synchronize()
t0 = time.perf_counter(); submit_only(3, steps); t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue per step %.3f ms" % ((t1 - t0) / (steps - 3) * 1e3))
if "--profile" in sys.argv:
    pr = cProfile.Profile(); pr.enable(); submit_only(3, steps); pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
