"""Parity at the named shapes (BASELINE.json configs 3-4: FB15k-237 / WN18RR shaped synthetic KGs with
rule sets of the reference files' shape).  Oracle comparison on sampled rules, plus size-independent
properties on whole batches: sparse-aware == dense expansion bit for bit, batch independence
(a query's counts do not depend on its batch mates), length-1 identities."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
_cache = {}


def workload(name):
    if name not in _cache:
        from rnnlogic_b200 import synth, KnowledgeGraph, CompiledRules
        from oracle import rnnlogic_oracle as O
        shape = synth.load_shape(name)
        N, R, train, valid, test = synth.synthetic_kg(shape)
        rules = synth.synthetic_rules(shape)
        kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
        okg = O.OracleKG(N, R, train, valid[:0], test[:0])
        _cache[name] = (kg, okg, rules, CompiledRules(kg, rules), train)
    return _cache[name]


def batches_of(train, R, n, seed):
    rng = np.random.default_rng(seed)
    out = []
    rels = rng.choice(np.unique(train[:, 1]), size=n, replace=False)
    for r in rels:
        grp = train[train[:, 1] == r]
        out.append(grp[rng.permutation(grp.shape[0])[:32]])
    return out


@pytest.mark.parametrize("name", ["fb15k237", "wn18rr"])
def test_counts_vs_oracle_and_modes_agree(name):
    from rnnlogic_b200.engine import Grounder
    kg, okg, rules, cr, train = workload(name)
    sparse = Grounder(kg, cr, DEV)
    dense = Grounder(kg, cr, DEV, force_dense=True)
    rng = np.random.default_rng(0)
    maxc = 0
    for b in batches_of(train, kg.relation_size, 4, 1):
        q = int(b[0, 1])
        ids = cr.head_rules[q]
        if not ids:
            continue
        etr = torch.from_numpy(kg.edge_index_of(b)).to(DEV)
        h = torch.from_numpy(b[:, 0]).to(DEV)
        s1 = sparse.ground(sparse.make_slots([q], [len(b)], h, None, etr))
        s2 = dense.ground(dense.make_slots([q], [len(b)], h, None, etr))
        longest = sorted(ids, key=lambda i: -len(rules[i][1]))[:6]
        pick = sorted(set(int(i) for i in rng.choice(ids, size=min(10, len(ids)), replace=False)) | set(longest))
        c1 = sparse.rule_counts(s1, pick)
        c2 = dense.rule_counts(s2, pick)
        assert torch.equal(c1, c2)
        got = c1.cpu().numpy()
        for k, rid in enumerate(pick):
            want = okg.grounding(b[:, 0], q, rules[rid][1], etr.cpu().numpy())
            assert np.array_equal(got[k], want), (name, q, rid, rules[rid])
            maxc = max(maxc, int(want.max()))
        # batch independence: the first 5 queries alone give the same rows
        n5 = min(5, len(b))
        s3 = sparse.ground(sparse.make_slots([q], [n5], h[:n5].contiguous(), None, etr[:n5].contiguous()))
        assert torch.equal(sparse.rule_counts(s3, pick[:4]), c1[:4, :n5])
        with pytest.raises(ValueError):
            sparse.make_slots([q], [len(b) + 2], h, None, etr)
    assert maxc >= 1


def test_length1_identities():
    """rule q <- q: without removal the row is the adjacency row; with removal exactly the query's own
    tail loses one path; rule q <- p (p != q) ignores edges_to_remove (data.py:143-146)."""
    from rnnlogic_b200 import KnowledgeGraph
    kg, okg, rules, cr, train = workload("fb15k237")
    b = batches_of(train, kg.relation_size, 1, 3)[0]
    q = int(b[0, 1])
    h = torch.from_numpy(b[:, 0]).to(DEV)
    etr = torch.from_numpy(kg.edge_index_of(b)).to(DEV)
    full = kg.grounding(h, q, [q], None).cpu().numpy()
    cut = kg.grounding(h, q, [q], etr).cpu().numpy()
    deg = {}
    for hh, rr, tt in train[train[:, 1] == q].tolist():
        deg[hh] = deg.get(hh, 0) + 1
    assert np.array_equal(full.sum(1), [deg[int(x)] for x in b[:, 0]])
    diff = full - cut
    assert np.array_equal(diff.sum(1), np.ones(len(b), dtype=np.int64))
    assert all(diff[i, b[i, 2]] == 1 for i in range(len(b)))
    p = (q + 1) % kg.relation_size
    assert np.array_equal(kg.grounding(h, q, [p], etr).cpu().numpy(), kg.grounding(h, q, [p], None).cpu().numpy())


def test_predictor_scores_at_fb_shape_vs_oracle():
    from rnnlogic_b200.predictors import Predictor
    from oracle import rnnlogic_oracle as O
    kg, okg, rules, cr, train = workload("fb15k237")
    m = Predictor(kg, "bias")
    m.set_rules([[h] + list(b) for h, b in rules])
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        m.rule_weights.copy_(torch.randn(m.num_rules, generator=g) * 0.1)
        m.bias.copy_(torch.randn(kg.entity_size, generator=g) * 0.1)
    w, bias = m.rule_weights.detach().clone(), m.bias.detach().clone()
    m = m.cuda()
    table = O.relation2rules(O.parse_rules([[h] + list(b) for h, b in rules]), kg.relation_size)
    b = batches_of(train, kg.relation_size, 2, 5)[1]
    q = int(b[0, 1])
    etr = kg.edge_index_of(b)
    tri = torch.from_numpy(b).to(DEV)
    with torch.no_grad():
        score, mask = m(tri[:, 0], tri[:, 1], torch.from_numpy(etr).to(DEV))
    want, wmask = O.predictor_forward(okg, table[q], w, bias, torch.from_numpy(b[:, 0]), etr, q)
    np.testing.assert_allclose(score.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)
    # fused loss == oracle loss on the same batch
    target = torch.zeros(len(b), kg.entity_size)
    for k, (hh, rr, tt) in enumerate(b.tolist()):
        target[k, torch.tensor(kg.hr2o[kg.encode_hr(hh, rr)])] = 1
    loss_ref = O.ce_loss(want, wmask, O.smoothed_target(target, torch.from_numpy(b[:, 2]), 0.2))
    loss, _ = m.fused_train_step([b], 0.2)
    np.testing.assert_allclose(loss[0].item(), loss_ref.item(), rtol=1e-5)


def test_config5_scaled_graph():
    """BASELINE config 5: N = 1e6 entities, R = 1e3 relations, 2e7 train edges, 1e4 length-3 rules.
    Bit-exact counts vs the C oracle on sampled rules, a fused train step and a filtered-rank call."""
    from rnnlogic_b200 import synth, KnowledgeGraph, CompiledRules
    from rnnlogic_b200.engine import Grounder
    from rnnlogic_b200.predictors import Predictor
    from oracle import rnnlogic_oracle as O
    N, R, train, valid, test = synth.scaled_kg()
    rules = synth.scaled_rules()
    kg = KnowledgeGraph(entity_size=N, relation_size=R, train=train, valid=valid, test=test)
    assert train.shape[0] == 20_000_000 and kg.host["row_dst"].shape[0] > 5_000_000
    okg = O.OracleKG.grounding_only(N, R, train)
    cr = CompiledRules(kg, rules)
    gr = Grounder(kg, cr, DEV)
    heads = [q for q in range(R) if len(cr.head_rules[q]) >= 5][:2]
    rng = np.random.default_rng(0)
    batches = []
    for q in heads:
        grp = train[train[:, 1] == q]
        b = grp[rng.permutation(grp.shape[0])[:32]]
        batches.append(b)
        etr = kg.edge_index_of(b)
        sl = gr.ground(gr.make_slots([q], [len(b)], torch.from_numpy(b[:, 0]).to(DEV), None, torch.from_numpy(etr).to(DEV)))
        pick = cr.head_rules[q][:3]
        got = gr.rule_counts(sl, pick).cpu().numpy()
        for k, rid in enumerate(pick):
            want = okg.grounding(b[:, 0], q, rules[rid][1], etr)
            assert np.array_equal(got[k], want), (q, rid)
    m = Predictor(kg, "bias")
    m.set_rules([[h] + list(b) for h, b in rules])
    with torch.no_grad():
        m.rule_weights.normal_(0, 0.1)
    m = m.cuda()
    loss, tsum = m.fused_train_step(batches, 0.2)
    assert torch.isfinite(loss).all() and (tsum > 0).all()
    assert m.rule_weights.grad is not None and torch.isfinite(m.rule_weights.grad).all()
    vb = [valid[valid[:, 1] == valid[0, 1]][:32]]
    LH = m.fused_rank(vb, "valid")
    assert LH.shape == (len(vb[0]), 2) and (LH[:, 0] >= 1).all() and (LH[:, 1] > LH[:, 0]).all() and (LH[:, 1] <= N + 1).all()
