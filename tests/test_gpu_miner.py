"""GPU rule discovery (rl_miner.cu) against the reference miner's own output (tests/golden/mined_*.npz)
and the C oracle (oracle_mine_rules): the same rules in the same order."""
import os

import numpy as np
import pytest

from tests import _golden as G

pytestmark = pytest.mark.gpu


class _KG:
    def __init__(self, fx):
        self.entity_size, self.relation_size = int(fx["N"]), int(fx["R"])
        self.train_array = fx["train"].astype(np.int64)


@pytest.mark.parametrize("name", G.DATASETS)
def test_mined_rules_equal_the_reference_miner(name):
    from rnnlogic_b200.miner import mine_rules
    fx = G.load(name)
    gold = dict(np.load(os.path.join(G.GOLDEN, "mined_%s.npz" % name)))
    want = [[int(v) for v in row if v >= 0] for row in gold["rules"]]
    got = mine_rules(_KG(fx), int(gold["max_length"]))
    assert len(got) == len(want)
    assert got == want


@pytest.mark.parametrize("max_length", [0, 1, 2, 4])
def test_other_lengths_match_the_oracle(max_length):
    from oracle import rnnlogic_oracle as O
    from rnnlogic_b200.miner import mine_rules
    fx = G.load("syn")
    kg = _KG(fx)
    rng = np.random.default_rng(max_length)
    sub = kg.train_array[rng.permutation(kg.train_array.shape[0])[:400 if max_length == 4 else 4000]]
    from rnnlogic_b200.miner import mine_rule_keys, decode_rule_keys
    keys = mine_rule_keys(kg.train_array, kg.entity_size, kg.relation_size, max_length, sub, table_log2=10)   # forces a table regrow
    got = decode_rule_keys(keys, kg.relation_size, max_length)
    want = O.mine_rules(kg.train_array, kg.entity_size, kg.relation_size, max_length, triples=sub)
    assert got == want and (max_length == 0 or len(got) > 0)


def test_edge_cases_and_the_pybind_names():
    """h == t -> empty body; r <- r dropped; the triple itself removed; a path stops at its first visit of t; the names of
    miner/pyrnnlogic.cpp (new_rule_miner / run_rule_miner / get_logic_rules)."""
    from oracle import rnnlogic_oracle as O
    from rnnlogic_b200 import miner

    class KG:
        entity_size, relation_size = 3, 3
        train_array = np.array([[0, 0, 1], [1, 1, 2], [0, 2, 2], [2, 0, 2], [2, 1, 0]], dtype=np.int64)
    m = miner.new_rule_miner(KG)
    miner.run_rule_miner(m, 3, 1.0, 8)
    rules = [list(r) for r in miner.get_logic_rules(m)]
    assert rules == O.mine_rules(KG.train_array, 3, 3, 3)
    assert [0] in rules and [2, 0, 1] in rules and [2, 2] not in rules
    half = miner.mine_rules(KG, 3, portion=0.5, seed=1)
    assert set(map(tuple, half)) <= set(map(tuple, rules))


def test_mined_rules_feed_the_predictor():
    """The rule list goes straight into set_rules (the miner FILE has a trailing H column that set_rules cannot read,
    SURVEY quirk 1; the list has not)."""
    import torch
    from rnnlogic_b200 import KnowledgeGraph
    from rnnlogic_b200.miner import mine_rules
    from rnnlogic_b200.predictors import Predictor
    fx = G.load("umls")
    kg = KnowledgeGraph(entity_size=int(fx["N"]), relation_size=int(fx["R"]), train=fx["train"], valid=fx["valid"], test=fx["test"])
    rules = mine_rules(kg, 2)
    model = Predictor(kg, "bias")
    model.set_rules(rules)
    model = model.cuda()
    b = fx["train"][fx["train"][:, 1] == fx["train"][0, 1]][:32].astype(np.int64)
    loss, tsum = model.fused_train_step([b], 0.2)
    assert torch.isfinite(loss).all() and float(tsum[0]) > 0
