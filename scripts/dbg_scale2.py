import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from tests.test_gpu_scale import workload, batches_of, DEV
from rnnlogic_b200.engine import Grounder
kg, okg, rules, cr, train = workload("fb15k237")
sparse = Grounder(kg, cr, DEV)
dense = Grounder(kg, cr, DEV, force_dense=True)
rng = np.random.default_rng(0)
def sync(tag):
    torch.cuda.synchronize(); print('   ok', tag, flush=True)
for b in batches_of(train, kg.relation_size, 4, 1):
    q = int(b[0, 1]); ids = cr.head_rules[q]
    print('batch q', q, 'B', len(b), flush=True)
    if not ids: continue
    etr = torch.from_numpy(kg.edge_index_of(b)).to(DEV)
    h = torch.from_numpy(b[:, 0]).to(DEV)
    sl1 = sparse.make_slots([q], [len(b)], h, None, etr); sync('slots1')
    s1 = sparse.ground(sl1); sync('ground1')
    sl2 = dense.make_slots([q], [len(b)], h, None, etr); sync('slots2')
    s2 = dense.ground(sl2); sync('ground2')
    longest = sorted(ids, key=lambda i: -len(rules[i][1]))[:6]
    pick = sorted(set(int(i) for i in rng.choice(ids, size=min(10, len(ids)), replace=False)) | set(longest))
    print('   pick', pick, [cr.rule_node[i] for i in pick], flush=True)
    c1 = sparse.rule_counts(s1, pick); sync('counts1')
    c2 = dense.rule_counts(s2, pick); sync('counts2')
    assert torch.equal(c1, c2)
    sl3 = sparse.make_slots([q], [5], h[:5].contiguous(), None, etr[:5].contiguous()); sync('slots3')
    s3 = sparse.ground(sl3); sync('ground3')
    c3 = sparse.rule_counts(s3, pick[:4]); sync('counts3')
    assert torch.equal(c3, c1[:4, :5])
print('all ok')
