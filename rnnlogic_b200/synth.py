"""Synthetic knowledge graphs and rule sets of the shapes BASELINE.json names (there is no
network and the reference's FB15k-237 / WN18RR train files are missing).

``shapes/*.json`` hold only aggregate statistics of the reference's inputs (sizes, per-relation
counts, per-head rule-length and trie-depth histograms) made by scripts/make_shapes.py."""
from __future__ import annotations

import json
import os
from typing import List, Tuple

import numpy as np

_SHAPES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shapes")


def load_shape(name: str) -> dict:
    with open(os.path.join(_SHAPES, name + ".json")) as f:
        return json.load(f)


def _draw_triples(rng, n_target, half, N, rw, ew, perm, taken):
    """i.i.d. (r ~ rw, h,t ~ Zipf popularity), no self loops / duplicates / triples
    already in ``taken`` (a list of sorted int64 key arrays, extended in place).  Fully vectorised."""
    out_keys = np.zeros(0, dtype=np.int64)
    order = np.zeros(0, dtype=np.int64)
    cdf_r, cdf_e = np.cumsum(rw), np.cumsum(ew)
    while out_keys.shape[0] < n_target:
        k = int((n_target - out_keys.shape[0]) * 1.2) + 64
        r = np.minimum(np.searchsorted(cdf_r, rng.random(k)), half - 1).astype(np.int64)
        h = perm[np.minimum(np.searchsorted(cdf_e, rng.random(k)), N - 1)].astype(np.int64)
        t = perm[np.minimum(np.searchsorted(cdf_e, rng.random(k)), N - 1)].astype(np.int64)
        key = (r * N + h) * N + t
        ok = h != t
        _, first = np.unique(key, return_index=True)            # first occurrence inside this draw
        uniq = np.zeros(k, dtype=bool)
        uniq[first] = True
        ok &= uniq
        for prev in taken + [out_keys]:
            if prev.shape[0]:
                pos = np.minimum(np.searchsorted(prev, key), prev.shape[0] - 1)
                ok &= prev[pos] != key
        new = key[ok][: n_target - out_keys.shape[0]]             # draw order kept
        out_keys = np.sort(np.concatenate([out_keys, new]))
        order = np.concatenate([order, new])
    taken.append(out_keys)
    key = order
    t = key % N
    h = (key // N) % N
    r = key // (N * N)
    return np.stack([h, r, t], 1)


def _with_inverses(base: np.ndarray, half: int) -> np.ndarray:
    inv = np.stack([base[:, 2], base[:, 1] + half, base[:, 0]], 1)
    return np.stack([base, inv], 1).reshape(-1, 3)


def synthetic_kg(shape: dict, seed: int = None, scale: float = 1.0):
    """(N, R, train, valid, test): calibrated recipe of SURVEY App. C.  Every base triple (h,r,t)
    is followed by its inverse (t, r+R/2, h).  valid/test are synthetic too (same recipe)."""
    seed = shape["seed"] if seed is None else seed
    rng = np.random.default_rng(seed)
    N, R = shape["num_entities"], shape["num_relations"]
    half = R // 2
    rw = np.asarray(shape["eval_count_per_base_relation"], dtype=np.float64) + 1.0
    rw /= rw.sum()
    ew = (np.arange(N) + 1.0) ** (-shape["entity_zipf"])
    ew /= ew.sum()
    perm = rng.permutation(N)
    taken: list = []
    n_valid, n_test = shape["valid_triples"] // 2, shape["test_triples"] // 2
    valid = _draw_triples(rng, int(n_valid * scale), half, N, rw, ew, perm, taken)
    test = _draw_triples(rng, int(n_test * scale), half, N, rw, ew, perm, taken)
    train = _draw_triples(rng, int(shape["train_base_triples"] * scale), half, N, rw, ew, perm, taken)
    if N * N * max(R, 1) < 2 ** 62:                              # one sort of a packed (h, r, t) key: same order as the lexsort
        train = train[np.argsort((train[:, 0] * R + train[:, 1]) * N + train[:, 2], kind="stable")]
    else:
        train = train[np.lexsort((train[:, 2], train[:, 1], train[:, 0]))]
    return N, R, _with_inverses(train, half), _with_inverses(valid, half), _with_inverses(test, half)


def synthetic_rules(shape: dict, seed: int = None) -> List[Tuple[int, List[int]]]:
    """A rule list with the reference file's shape: per head the same number of rules of each body
    length and (up to feasibility) the same number of distinct body prefixes per depth, body
    relations drawn from the file's usage histogram."""
    seed = shape["seed"] if seed is None else seed
    rng = np.random.default_rng(seed + 1)
    R, lmax = shape["num_relations"], shape["max_len"]
    usage = np.asarray(shape["body_relation_usage"], dtype=np.float64) + 1.0
    prob = usage / usage.sum()
    rules: List[Tuple[int, List[int]]] = []
    for q in range(R):
        n_len = shape["rules_per_head_by_len"][q]
        d_dep = shape["trie_nodes_per_head_by_depth"][q]
        rules += [(q, [])] * n_len[0]
        levels: List[List[tuple]] = []
        for depth in range(1, lmax + 1):
            want = d_dep[depth - 1]
            nodes: List[tuple] = []
            if want:
                if depth == 1:
                    rels = rng.choice(R, size=min(want, R), replace=False, p=prob)
                    nodes = [(int(r),) for r in rels]
                else:
                    parents = levels[-1]
                    seen = set()
                    # parents that are not rule ends themselves need a child first
                    need = max(0, len(parents) - n_len[depth - 1])
                    order = list(rng.permutation(len(parents))[:need]) + list(rng.integers(len(parents), size=max(0, want - need)))
                    for pi in order[:want]:
                        for _try in range(8):
                            cand = parents[int(pi)] + (int(rng.choice(R, p=prob)),)
                            if cand not in seen:
                                seen.add(cand)
                                nodes.append(cand)
                                break
            levels.append(nodes)
        for depth in range(1, lmax + 1):
            nodes = levels[depth - 1]
            n = n_len[depth]
            if n == 0 or not nodes:
                continue
            has_child = set(c[:-1] for c in levels[depth]) if depth < lmax else set()
            leaves = [p for p in nodes if p not in has_child]
            inner = [p for p in nodes if p in has_child]
            chosen = leaves[:n]
            if len(chosen) < n:
                extra = [inner[i] for i in rng.permutation(len(inner))[: n - len(chosen)]]
                chosen += extra
            while len(chosen) < n:                      # duplicates, as in the reference file
                chosen.append(nodes[int(rng.integers(len(nodes)))])
            rules += [(q, list(p)) for p in chosen]
    order = rng.permutation(len(rules))
    return [rules[i] for i in order]


def scaled_kg(N=1_000_000, R=1000, E=20_000_000, seed=5):
    """Config 5: relation sizes Zipf(1.0) over R/2 base relations, entity popularity Zipf(0.7)."""
    shape = {"num_entities": N, "num_relations": R, "train_base_triples": E // 2, "entity_zipf": 0.7, "seed": seed,
             "valid_triples": 2000, "test_triples": 2000,
             "eval_count_per_base_relation": (1e6 / (np.arange(R // 2) + 1.0)).tolist()}
    return synthetic_kg(shape)


def scaled_rules(R=1000, n_rules=10_000, length=3, seed=5):
    rng = np.random.default_rng(seed)
    heads = rng.integers(R, size=n_rules)
    bodies = rng.integers(R, size=(n_rules, length))
    return [(int(h), [int(b) for b in body]) for h, body in zip(heads, bodies)]


# ------------------------------------------------------------------------------------------------
# A TYPED synthetic graph: the i.i.d. recipe above draws heads and tails independently of the relation, so a
# hop rho' -> rho almost never composes (the tails of rho' are rarely heads of rho) and frontiers die after one
# hop.  Real KGs are typed: every relation has a domain and a range, and mined rule bodies chain relations whose
# range / domain agree -- which is why they were mined.  Same sizes as the shape, same Zipf popularity, but every
# entity has a type, every base relation a (domain, range) pair, and rule bodies are type-compatible chains that
# start in the head's domain and end in its range.
# ------------------------------------------------------------------------------------------------
def typed_kg(shape: dict, n_types: int = 12, seed: int = None):
    """(N, R, train, valid, test, meta) with meta = {"etype": [N], "domain": [R], "range": [R]}."""
    seed = shape["seed"] if seed is None else seed
    rng = np.random.default_rng(seed + 7)
    N, R = shape["num_entities"], shape["num_relations"]
    half = R // 2
    tw = (np.arange(n_types) + 1.0) ** -0.5
    tw /= tw.sum()
    etype = rng.choice(n_types, size=N, p=tw)
    members = [np.flatnonzero(etype == t) for t in range(n_types)]
    for t in range(n_types):                                    # no empty type
        if members[t].shape[0] == 0:
            etype[t] = t
    members = [np.flatnonzero(etype == t) for t in range(n_types)]
    pop = []                                                    # Zipf popularity inside a type
    for t in range(n_types):
        w = (np.arange(members[t].shape[0]) + 1.0) ** (-shape["entity_zipf"])
        pop.append(np.cumsum(w / w.sum()))
    dom = rng.choice(n_types, size=half, p=tw)
    ran = rng.choice(n_types, size=half, p=tw)
    rw = np.asarray(shape["eval_count_per_base_relation"], dtype=np.float64) + 1.0
    rw /= rw.sum()
    cdf_r = np.cumsum(rw)
    taken = np.zeros(0, dtype=np.int64)

    def draw(n_target):
        nonlocal taken
        out = np.zeros(0, dtype=np.int64)
        while out.shape[0] < n_target:
            k = int((n_target - out.shape[0]) * 1.3) + 64
            r = np.minimum(np.searchsorted(cdf_r, rng.random(k)), half - 1).astype(np.int64)
            u, v = rng.random(k), rng.random(k)
            h = np.empty(k, dtype=np.int64)
            t = np.empty(k, dtype=np.int64)
            for ty in range(n_types):
                sel = dom[r] == ty
                if sel.any():
                    h[sel] = members[ty][np.minimum(np.searchsorted(pop[ty], u[sel]), members[ty].shape[0] - 1)]
                sel = ran[r] == ty
                if sel.any():
                    t[sel] = members[ty][np.minimum(np.searchsorted(pop[ty], v[sel]), members[ty].shape[0] - 1)]
            key = (r * N + h) * N + t
            ok = h != t
            _, first = np.unique(key, return_index=True)
            uniq = np.zeros(k, dtype=bool)
            uniq[first] = True
            ok &= uniq
            for prev in (taken, out):
                if prev.shape[0]:
                    srt = np.sort(prev)
                    pos = np.minimum(np.searchsorted(srt, key), srt.shape[0] - 1)
                    ok &= srt[pos] != key
            out = np.concatenate([out, key[ok][: n_target - out.shape[0]]])
        taken = np.concatenate([taken, out])
        return np.stack([(out // N) % N, out // (N * N), out % N], 1)

    valid = draw(shape["valid_triples"] // 2)
    test = draw(shape["test_triples"] // 2)
    train = draw(shape["train_base_triples"])
    if N * N * max(R, 1) < 2 ** 62:                              # one sort of a packed (h, r, t) key: same order as the lexsort
        train = train[np.argsort((train[:, 0] * R + train[:, 1]) * N + train[:, 2], kind="stable")]
    else:
        train = train[np.lexsort((train[:, 2], train[:, 1], train[:, 0]))]
    meta = {"etype": etype, "domain": np.concatenate([dom, ran]), "range": np.concatenate([ran, dom])}
    return N, R, _with_inverses(train, half), _with_inverses(valid, half), _with_inverses(test, half), meta


def typed_rules(shape: dict, meta: dict, seed: int = None) -> List[Tuple[int, List[int]]]:
    """Per head the shape's number of rules of each body length; bodies are type-compatible chains from the head's
    domain to its range (the last constraint is dropped when no chain of that length satisfies it)."""
    seed = shape["seed"] if seed is None else seed
    rng = np.random.default_rng(seed + 11)
    R, lmax = shape["num_relations"], shape["max_len"]
    dom, ran = meta["domain"], meta["range"]
    n_types = int(max(dom.max(), ran.max())) + 1
    by_dom = [np.flatnonzero(dom == t) for t in range(n_types)]
    by_pair = {(a, b): np.flatnonzero((dom == a) & (ran == b)) for a in range(n_types) for b in range(n_types)}
    rules: List[Tuple[int, List[int]]] = []
    for q in range(R):
        n_len = shape["rules_per_head_by_len"][q]
        rules += [(q, [])] * n_len[0]
        for L in range(1, lmax + 1):
            for _ in range(n_len[L]):
                body, cur = [], int(dom[q])
                for pos in range(L):
                    last = pos == L - 1
                    cand = by_pair[(cur, int(ran[q]))] if last else by_dom[cur]
                    if cand.shape[0] == 0:
                        cand = by_dom[cur] if by_dom[cur].shape[0] else np.arange(R)
                    rel = int(cand[rng.integers(cand.shape[0])])
                    body.append(rel)
                    cur = int(ran[rel])
                rules.append((q, body))
    order = rng.permutation(len(rules))
    return [rules[i] for i in order]
