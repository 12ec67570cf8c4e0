"""GPU parity of kernel (1): frontier expansion vs the reference's golden counts and the oracle.
Bit-exact int64 (north_star)."""
import numpy as np
import pytest
import torch

from tests import _golden as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=G.DATASETS)
def ds(request):
    from rnnlogic_b200 import KnowledgeGraph, parse_rules
    from oracle import rnnlogic_oracle as O
    fx = G.load(request.param)
    kg = KnowledgeGraph(entity_size=int(fx["N"]), relation_size=int(fx["R"]), train=fx["train"],
                        valid=fx["valid"], test=fx["test"])
    okg = O.OracleKG(int(fx["N"]), int(fx["R"]), fx["train"], fx["valid"], fx["test"])
    return request.param, fx, kg, okg, parse_rules(G.rules_of(fx))


def test_grounding_api_matches_reference_golden(ds):
    name, fx, kg, okg, rules = ds
    dev = torch.device("cuda:0")
    for c in range(int(fx["gr_n"])):
        tri, _, etr = G.train_batch_inputs(fx, int(fx["gr%d_batch" % c]))
        head, body = rules[int(fx["gr%d_rule" % c])]
        use = bool(fx["gr%d_etr" % c])
        got = kg.grounding(torch.from_numpy(tri[:, 0]).to(dev), head, body, etr.to(dev) if use else None)
        assert got.dtype == torch.int64 and got.shape == (tri.shape[0], kg.entity_size)
        assert np.array_equal(got.cpu().numpy(), fx["gr%d_counts" % c]), (name, c, body)


@pytest.mark.parametrize("mode", ["auto", "dense", "sparse"])
@pytest.mark.parametrize("bits", [32, 64])
def test_all_rules_of_head_vs_oracle(ds, mode, bits):
    """Every rule of the head through the shared trie == per-rule oracle grounding (with the
    query edge removed), for train batches (B up to 50 -> 2 slots)."""
    from rnnlogic_b200 import CompiledRules
    from rnnlogic_b200.engine import Grounder
    name, fx, kg, okg, rules = ds
    dev = torch.device("cuda:0")
    cr = CompiledRules(kg, rules)
    gr = Grounder(kg, cr, dev, force_dense=(mode == "dense"))
    if mode == "sparse":
        gr.dense_num, gr.dense_den = 1 << 20, 1      # never switch a node to all-rows
    gr.force_bits = bits
    for j in range(0, int(fx["tb_n"]), 3):
        tri, _, etr = G.train_batch_inputs(fx, j)
        q = int(tri[0, 1])
        ids = cr.head_rules[q]
        if not ids:
            continue
        for use in (True, False):
            sl = gr.make_slots([q], [tri.shape[0]], torch.from_numpy(tri[:, 0]).to(dev), None,
                               etr.to(dev) if use else None)
            gr.ground(sl)
            assert int(sl.overflow.item()) == 0
            pick = ids[:: max(1, len(ids) // 40)]
            got = gr.rule_counts(sl, pick).cpu().numpy()
            for i, rid in enumerate(pick):
                want = okg.grounding(tri[:, 0], q, rules[rid][1], etr.numpy() if use else None)
                assert np.array_equal(got[i], want), (name, j, rid, rules[rid])


def test_overflow_falls_back_to_64bit():
    """Complete bipartite-ish graph: counts exceed 2^32 after 8 hops; int64 result must match the
    oracle bit for bit (and wrap like int64 beyond 2^63)."""
    from rnnlogic_b200 import KnowledgeGraph
    from oracle import rnnlogic_oracle as O
    N = 40
    tri = np.array([(a, 0, b) for a in range(N) for b in range(N) if a != b], dtype=np.int64)
    kg = KnowledgeGraph(entity_size=N, relation_size=2, train=tri)
    okg = O.OracleKG(N, 2, tri, np.zeros((0, 3), np.int64), np.zeros((0, 3), np.int64))
    h = torch.arange(5)
    for L in (6, 8, 13):
        body = [0] * L
        got = kg.grounding(h.cuda(), 1, body, None).cpu().numpy()
        want = okg.grounding(h.numpy(), 1, body, None)
        assert np.array_equal(got, want), L
        if L == 8:
            assert want.max() > 2 ** 32


def test_cpu_tensors_fail_loudly():
    from rnnlogic_b200 import KnowledgeGraph, _lib
    kg = KnowledgeGraph(entity_size=4, relation_size=1, train=np.array([[0, 0, 1], [1, 0, 2]]))
    with pytest.raises(_lib.RlError):
        kg.grounding(torch.tensor([0]), 0, [0], None)
