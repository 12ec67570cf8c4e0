"""Rule discovery on the GPU: the rule-search half of the reference's C++ miner
(miner/rnnlogic.cpp:350-382 KnowledgeGraph::rule_search, :505-589 RuleMiner::search) behind the names its pybind
module exposes (miner/pyrnnlogic.cpp:161-179: new_rule_miner / run_rule_miner / get_logic_rules).

    rules = mine_rules(graph, max_length=3)          # [[head, b1, ..., bk], ...] in the order of the reference's rule list
    model.set_rules(rules)

The H-score / reasoning-predictor half of the miner binary (ReasoningPredictor, RuleGenerator) is the CPU twin of the
path this package accelerates and is not rebuilt here: train PredictorPlus on the mined rules instead."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib
from .engine import _stream


def _adjacency(train: np.ndarray, N: int):
    """Out-edges by source entity over ALL relations (the miner's e2r2n, rnnlogic.cpp:275-281)."""
    order = np.argsort(train[:, 0], kind="stable")
    tr = train[order]
    ptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(np.bincount(tr[:, 0], minlength=N), out=ptr[1:])
    return ptr.astype(np.int32), np.ascontiguousarray(tr[:, 1], dtype=np.int32), np.ascontiguousarray(tr[:, 2], dtype=np.int32)


def mine_rule_keys(train: np.ndarray, N: int, R: int, max_length: int = 3, triples: np.ndarray = None, device="cuda",
                   table_log2: int = 22) -> np.ndarray:
    """Sorted 64-bit keys of the mined rules (csrc/rl_miner.cu for the packing).  ``triples`` = the train facts that
    are searched (default: all of them; the reference's ``portion`` takes a random prefix of a shuffle)."""
    train = np.ascontiguousarray(train, dtype=np.int64).reshape(-1, 3)
    triples = train if triples is None else np.ascontiguousarray(triples, dtype=np.int64).reshape(-1, 3)
    rel_bits = max(1, int(R - 1).bit_length())
    if rel_bits * (max_length + 1) + 3 > 63:
        raise ValueError("max_length %d with %d relations does not fit a 64-bit rule key" % (max_length, R))
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.RlError("rnnlogic_b200.miner is CUDA-only")
    with torch.cuda.device(dev):
        ptr, rel, dst = (torch.from_numpy(a).to(dev) for a in _adjacency(train, N))
        tri = torch.from_numpy(np.ascontiguousarray(triples, dtype=np.int32)).to(dev)
        flags = torch.zeros(4, dtype=torch.int32, device=dev)
        while True:
            cap = 1 << table_log2
            table = torch.full((cap,), -1, dtype=torch.int64, device=dev)
            flags.zero_()
            _lib.check(_lib.lib().rl_mine_rules(int(tri.shape[0]), tri.data_ptr(), ptr.data_ptr(), rel.data_ptr(), dst.data_ptr(),
                                               int(max_length), rel_bits, table.data_ptr(), cap, flags.data_ptr(), _stream()),
                       "rl_mine_rules")
            keys = table[table != -1]
            full = int(flags[0].item()) != 0 or keys.numel() * 2 > cap          # keep the load factor below 1/2
            if not full:
                break
            table_log2 += 3
            del table, keys
        return torch.sort(keys)[0].cpu().numpy()


def decode_rule_keys(keys: np.ndarray, R: int, max_length: int) -> List[List[int]]:
    b = max(1, int(R - 1).bit_length())
    keys = keys.astype(np.uint64)
    mask = np.uint64((1 << b) - 1)
    head = (keys >> np.uint64(b * max_length + 3)).astype(np.int64)
    length = ((keys >> np.uint64(b * max_length)) & np.uint64(7)).astype(np.int64)
    body = np.stack([((keys >> np.uint64(b * (max_length - 1 - i))) & mask).astype(np.int64) for i in range(max_length)], 1) \
        if max_length else np.zeros((keys.shape[0], 0), np.int64)
    return [[int(h)] + [int(x) for x in bd[:n]] for h, n, bd in zip(head, length, body)]


def mine_rules(graph, max_length: int = 3, portion: float = 1.0, seed: int = 0, device="cuda") -> List[List[int]]:
    """All chain rules ``head <- b1 ... bk`` (k <= max_length) that connect the head and tail of at least one searched
    train fact with that fact removed, without ``r <- r``; ordered like the reference's rule list
    (head relation, then length, then body).  ``portion`` < 1 searches a random subset of the facts (RuleMiner::search's portion)."""
    train = np.asarray(graph.train_array if hasattr(graph, "train_array") else graph.train_facts, dtype=np.int64).reshape(-1, 3)
    triples = train
    if portion < 1.0:
        rng = np.random.default_rng(seed)
        triples = train[rng.permutation(train.shape[0])[:int(train.shape[0] * portion)]]
    keys = mine_rule_keys(train, graph.entity_size, graph.relation_size, max_length, triples, device)
    return decode_rule_keys(keys, graph.relation_size, max_length)


# ---- the names of miner/pyrnnlogic.cpp:161-179 (rule-search subset) ----
class _RuleMiner:
    def __init__(self, graph):
        self.graph, self.rules = graph, []


def new_rule_miner(graph):
    return _RuleMiner(graph)


def run_rule_miner(miner: _RuleMiner, max_length: int, portion: float = 1.0, num_threads: int = 0):
    """num_threads is accepted for signature compatibility; the search runs on the GPU."""
    miner.rules = mine_rules(miner.graph, max_length, portion)


def get_logic_rules(miner: _RuleMiner) -> Sequence[Sequence[int]]:
    return miner.rules
