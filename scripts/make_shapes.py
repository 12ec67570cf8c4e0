#!/usr/bin/env python
"""Extract SHAPE statistics (no triples, no rules) of the reference's FB15k-237 / WN18RR inputs
into rnnlogic_b200/shapes/*.json so that bench.py can synthesise KGs and rule sets of the named
shapes on a box without /root/reference (BASELINE.json: "synthetic KGs and rule sets of the named
shapes").  Run in the build container only."""
import json
import os
import sys
from collections import Counter

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rnnlogic_b200", "shapes")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, e_half, zipf, seed in (("FB15k-237", 272115, 0.7, 237), ("wn18rr", 86835, 0.6, 18)):
        d = os.path.join(REF, name)
        ent = sum(1 for _ in open(os.path.join(d, "entities.dict")))
        rel2id = {}
        for line in open(os.path.join(d, "relations.dict")):
            i, r = line.strip().split("\t")
            rel2id[r] = int(i)
        R = len(rel2id)
        half = R // 2
        cnt = Counter()
        n_eval = {}
        for split in ("valid", "test"):
            n = 0
            for line in open(os.path.join(d, split + ".txt")):
                h, r, t = line.strip().split("\t")
                rid = rel2id[r]
                n += 1
                if rid < half:
                    cnt[rid] += 1
            n_eval[split] = n
        # rule-file shape: per head (#rules of length 0..Lmax, #trie nodes per depth) + body-relation usage
        rules = [tuple(int(v) for v in line.split()) for line in open(os.path.join(d, "rnnlogic_rules.txt"))]
        lmax = max(len(r) - 1 for r in rules)
        per_len = [[0] * (lmax + 1) for _ in range(R)]
        tries = [set() for _ in range(R)]
        usage = [0] * R
        for r in rules:
            per_len[r[0]][len(r) - 1] += 1
            for k in range(2, len(r) + 1):
                tries[r[0]].add(r[1:k])
            for b in r[1:]:
                usage[b] += 1
        per_depth = [[sum(1 for p in tries[q] if len(p) == dd) for dd in range(1, lmax + 1)] for q in range(R)]
        shape = {
            "name": name, "num_entities": ent, "num_relations": R, "train_base_triples": e_half,
            "entity_zipf": zipf, "seed": seed, "valid_triples": n_eval["valid"], "test_triples": n_eval["test"],
            "eval_count_per_base_relation": [cnt[i] for i in range(half)],
            "num_rules": len(rules), "max_len": lmax,
            "rules_per_head_by_len": per_len, "trie_nodes_per_head_by_depth": per_depth,
            "body_relation_usage": usage,
        }
        path = os.path.join(OUT, name.lower().replace("-", "") + ".json")
        with open(path, "w") as f:
            json.dump(shape, f, separators=(",", ":"))
        print(path, os.path.getsize(path), "bytes;", len(rules), "rules;",
              [sum(p[d] for p in per_depth) for d in range(lmax)], "trie nodes per depth")


if __name__ == "__main__":
    main()
