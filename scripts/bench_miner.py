"""Rule discovery: GPU kernel (rl_miner.cu) against the CPU restatement of the reference's DFS miner (oracle C, one
thread) on the golden datasets and on the FB15k-237-shape synthetic graph.  usage (GPU box): python scripts/bench_miner.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import _golden as G
from oracle import rnnlogic_oracle as O
from rnnlogic_b200 import miner, synth

def gpu(train, N, R, L, triples=None, table_log2=22):
    miner.mine_rule_keys(train[:64], N, R, 1)                       # warm-up (module load)
    torch.cuda.synchronize()
    t = time.perf_counter()
    keys = miner.mine_rule_keys(train, N, R, L, triples, table_log2=table_log2)
    torch.cuda.synchronize()
    return len(keys), time.perf_counter() - t

for name in ("umls", "kinship"):
    fx = G.load(name)
    tr, N, R = fx["train"].astype(np.int64), int(fx["N"]), int(fx["R"])
    n, tg = gpu(tr, N, R, 3)
    t = time.perf_counter(); want = O.mine_rules(tr, N, R, 3); tc = time.perf_counter() - t
    assert n == len(want)
    print("%-8s L=3: %6d triples -> %6d rules | GPU %.3f s | CPU oracle (1 thread) %.2f s | x%.0f" % (name, len(tr), n, tg, tc, tc / tg))
shape = synth.load_shape("fb15k237")
N, R, train, valid, test = synth.synthetic_kg(shape)
rng = np.random.default_rng(0)
sub = train[rng.permutation(len(train))[:20000]]
for L, n_gpu, n_cpu in ((2, 20000, 2000), (3, 2000, 100)):
    n, tg = gpu(train, N, R, L, sub[:n_gpu], table_log2=27)
    sub_c = sub[:n_cpu]
    t = time.perf_counter(); want = O.mine_rules(train, N, R, L, triples=sub_c); tc = time.perf_counter() - t
    print("fb15k237-shape L=%d: %d of %d triples searched -> %d rules | GPU %.2f s (%.0f triples/s) | CPU oracle %d triples %.2f s (%.0f triples/s)"
          % (L, n_gpu, len(train), n, tg, n_gpu / tg, len(sub_c), tc, len(sub_c) / tc), flush=True)
