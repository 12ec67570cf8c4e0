"""GPU parity of the candidate-cell path (csrc/rl_cells.cu, rl_tail2.cu) against (a) the dense entity-major path
of round 1 on the same frontier, (b) the oracle, and of KnowledgeGraph.propagate against the oracle's C restatement
of src/data.py:149-173 (bit-exact int64)."""
import ctypes

import numpy as np
import pytest
import torch

from tests import _golden as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _world(seed=3, N=700, R=10, E=9000):
    from rnnlogic_b200 import KnowledgeGraph
    rng = np.random.default_rng(seed)
    tri = np.unique(np.stack([rng.integers(N, size=E), rng.integers(R, size=E), rng.integers(N, size=E)], 1), axis=0)
    rng.shuffle(tri)
    n_tr = int(0.8 * len(tri))
    kg = KnowledgeGraph(entity_size=N, relation_size=R, train=tri[:n_tr], valid=tri[n_tr:n_tr + 500], test=tri[n_tr + 500:])
    rules = []
    for q in range(R):
        rules.append([q])                                             # empty body
        for _ in range(12):
            L = int(rng.integers(1, 4))
            rules.append([q] + [int(x) for x in rng.integers(R, size=L)])
        rules.append([q, q])
        rules.append([q, q])                                          # duplicate rule
    return kg, rules, tri[:n_tr]


@pytest.mark.parametrize("ef", ["bias", "none"])
def test_cell_step_equals_dense_step(ef):
    """loss, target sums and the gradients of rule weights / bias: cells vs dense [S][N][32] matrices."""
    from rnnlogic_b200.predictors import Predictor
    from rnnlogic_b200 import cellpath
    kg, rules, tri = _world()
    torch.manual_seed(5)
    m = Predictor(kg, ef)
    m.set_rules(rules)
    with torch.no_grad():
        m.rule_weights.copy_(torch.randn(len(rules)) * 0.3)
        if ef == "bias":
            m.bias.copy_(torch.randn(kg.entity_size) * 0.5)
    m = m.cuda()
    sk = m._driver(torch.device(DEV))
    batches = [tri[tri[:, 1] == q][:40].astype(np.int64) for q in range(6)]         # 40 > 32: groups of two slots
    sl = sk.gr.make_slots_host(batches, with_etr=True)
    loss_d, tsum_d, msum_d, gw_d, gb_d = m.step_on_slots_dense(sk, sl, 0.2, 0.25)
    loss_d, tsum_d, gw_d = loss_d.clone(), tsum_d.clone(), gw_d.clone()
    gb_d = gb_d.clone() if gb_d is not None else None
    sl2 = sk.gr.make_slots_host(batches, with_etr=True, coo_only=True)
    gbuf = cellpath.GradBuffer(m.fused_params())
    loss_c, tsum_c = m.step_on_slots(sk, sl2, 0.2, 0.25, gbuf)
    flags = sl2.flags.cpu().numpy()
    assert flags[1] == 0 and flags[8] == 0 and flags[0] > 0
    np.testing.assert_allclose(loss_c.cpu().numpy(), loss_d.cpu().numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(tsum_c.cpu().numpy(), tsum_d.cpu().numpy(), rtol=1e-6)
    gw_c = gbuf.view(m.rule_weights).cpu().numpy()
    np.testing.assert_allclose(gw_c, gw_d.cpu().numpy(), rtol=2e-5, atol=1e-7 * max(1.0, float(gw_d.abs().max())))
    if ef == "bias":
        gb_c = gbuf.view(m.bias).cpu().numpy()
        np.testing.assert_allclose(gb_c, gb_d.cpu().numpy(), rtol=2e-5, atol=2e-7 * max(1.0, float(gb_d.abs().max())))
    else:
        got = np.add.reduceat(sl2.slot_ncell.cpu().numpy(), np.arange(0, sl2.S, 2))
        np.testing.assert_array_equal(got, msum_d.cpu().numpy().astype(np.int64))


@pytest.mark.parametrize("ef", ["bias", "none"])
def test_cell_rank_equals_dense_rank(ef):
    from rnnlogic_b200.predictors import Predictor, _valid_lanes
    kg, rules, tri = _world(seed=9)
    torch.manual_seed(2)
    m = Predictor(kg, ef)
    m.set_rules(rules)
    with torch.no_grad():
        m.rule_weights.copy_(torch.randn(len(rules)) * 0.3)
        if ef == "bias":
            m.bias.copy_((torch.randn(kg.entity_size) * 0.5).round(decimals=1))   # many exact ties between bare-bias logits
    m = m.cuda()
    sk = m._driver(torch.device(DEV))
    test = kg.test_array
    batches = [test[test[:, 1] == q][:32].astype(np.int64) for q in range(6) if (test[:, 1] == q).any()]
    for split in ("valid", "test"):
        got = m.fused_rank(batches, split).cpu().numpy()
        sl = sk.gr.make_slots_host(batches, with_etr=False)
        sk.gr.ground(sl)
        Z, nz = sk.predictor_scores(sl, m.rule_weights.detach(), m.bias.detach() if ef == "bias" else None, ef != "bias")
        want = _valid_lanes(sl, sk.filtered_rank(sl, Z, nz, "hr2oo" if split == "valid" else "hr2ooo", ef != "bias")).cpu().numpy()
        np.testing.assert_array_equal(got, want)


def test_pipelined_steps_redo_on_overflow():
    """A complete graph makes length-5 path counts overflow 32 bits: an enqueued step reports it (RlStepOverflow),
    fused_train_step redoes it with 64-bit rows, and the pipelined trainer ends with the same parameters as the
    synchronous one (the flags are read BEFORE the optimizer step)."""
    from rnnlogic_b200 import KnowledgeGraph
    from rnnlogic_b200.data import TrainDataset, ValidDataset, TestDataset
    from rnnlogic_b200.predictors import Predictor, RlStepOverflow
    from rnnlogic_b200.trainer import TrainerPredictor
    from rnnlogic_b200.utils import set_seed
    N = 300
    hh, tt = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    keep = hh != tt
    tri = np.stack([hh[keep], np.zeros(keep.sum(), np.int64), tt[keep]], 1).astype(np.int64)      # 89,700 edges of relation 0
    tri = np.concatenate([tri, np.array([[0, 1, 1], [1, 1, 2], [2, 1, 3]])])
    kg = KnowledgeGraph(entity_size=N, relation_size=2, train=tri, valid=tri[-3:], test=tri[-3:])
    rules = [[1, 0, 0, 0, 0, 0], [1, 0], [1, 1]]                    # ~299^4 = 8e9 paths per cell > 2^32
    results = []
    for pipelined in (True, False):
        set_seed(4)
        m = Predictor(kg, "bias")
        m.set_rules(rules)
        opt = torch.optim.Adam(m.parameters(), lr=0.01)
        tr = TrainDataset(kg, 32)
        tr.batches = [b for b in tr.batches if b[0][1] == 1] * 3
        tr.batch_arrays = [np.array(b, dtype=np.int64).reshape(-1, 3) for b in tr.batches]
        tr.make_batches = lambda: None
        solver = TrainerPredictor(m, tr, ValidDataset(kg, 32), TestDataset(kg, 32), opt, gpus=[0])
        solver.pipelined = pipelined
        if pipelined:
            tk = m.submit_train_step([tr.batch_arrays[0]], 0.2)
            with pytest.raises(RlStepOverflow):
                tk.result()
        solver.train(batch_per_epoch=3, smoothing=0.2, print_every=100)
        results.append({k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()})
    for k in results[0]:
        # the gradient of the length-5 rule is a sum of (p - target) * count over ~300 cells whose counts (~8e9) are almost
        # equal: catastrophic cancellation in fp32, so the order of the atomic adds (which follows the launch shape) shows up at the
        # 5e-2 level after three Adam steps
        np.testing.assert_allclose(results[0][k], results[1][k], rtol=1.5e-1, atol=1e-8)
    assert np.abs(results[0]["rule_weights"]).max() > 0


@pytest.mark.parametrize("name", ["kinship", "syn"])
def test_propagate_matches_oracle(name):
    """KnowledgeGraph.propagate (src/data.py:149-173) from an ARBITRARY int64 frontier, with and without
    edges_to_remove, against the oracle's C restatement -- bit-exact, including int64 wrap-around."""
    from rnnlogic_b200 import KnowledgeGraph
    from oracle import rnnlogic_oracle as O
    fx = G.load(name)
    kg = KnowledgeGraph(entity_size=int(fx["N"]), relation_size=int(fx["R"]), train=fx["train"], valid=fx["valid"], test=fx["test"])
    okg = O.OracleKG.grounding_only(kg.entity_size, kg.relation_size, fx["train"])
    lib = O._lib()
    rng = np.random.default_rng(0)
    N = kg.entity_size
    for rel in range(min(kg.relation_size, 6)):
        e0, e1 = int(okg.rel_ptr[rel]), int(okg.rel_ptr[rel + 1])
        if e1 == e0:
            continue
        for B, with_etr, big in ((5, False, False), (37, True, False), (32, True, True)):
            x = rng.integers(0, 7, size=(N, B)).astype(np.int64)
            if big:
                x[rng.integers(N, size=20), rng.integers(B, size=20)] = np.int64(2) ** 62          # sums wrap like int64
            etr = rng.integers(0, e1 - e0, size=B).astype(np.int64) if with_etr else None
            want = np.empty((N, B), dtype=np.int64)
            lib.oracle_propagate(ctypes.c_int64(N), ctypes.c_int64(B), ctypes.c_int64(e1 - e0),
                                 O._p(np.ascontiguousarray(okg.node_in[e0:e1])), O._p(np.ascontiguousarray(okg.node_out[e0:e1])),
                                 O._p(x), O._p(etr) if etr is not None else None, O._p(want))
            got = kg.propagate(torch.from_numpy(x).to(DEV).unsqueeze(-1), rel,
                               torch.from_numpy(etr).to(DEV) if etr is not None else None)
            assert got.shape == (N, B, 1) and got.dtype == torch.int64
            assert np.array_equal(got.squeeze(-1).cpu().numpy(), want), (name, rel, B)
    # two hops through propagate == grounding of the length-2 body (data.py:136-147)
    h = torch.from_numpy(fx["train"][:16, 0].astype(np.int64)).to(DEV)
    x = torch.nn.functional.one_hot(h, N).transpose(0, 1).unsqueeze(-1)
    y = kg.propagate(kg.propagate(x, 0), 1)
    assert torch.equal(y.squeeze(-1).transpose(0, 1), kg.grounding(h, 2 % kg.relation_size, [0, 1], None))
