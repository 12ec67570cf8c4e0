#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py [--mined-dir /tmp/miner]

It imports /root/reference/src (data.py, predictors.py, trainer.py ...) on CPU under two
stand-in modules for the third-party packages that are not installed here
(``torch_scatter.scatter`` == zeros().index_add_(), ``easydict.EasyDict``) and
records the reference's own outputs on seeded inputs.  Nothing from the reference's source
is copied: the fixtures hold integer-id inputs and numeric outputs only.  The GPU box has no
/root/reference, so the tests read these files instead.

Mined rule files (``-max-length 3`` output of the reference's C++ miner, README.md:49) are
taken from --mined-dir/{umls,kinship}_mined.txt when present (H column stripped, rules
sorted canonically because the miner's line order is not deterministic).
"""
import argparse
import io
import logging
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_standins():
    ts = types.ModuleType("torch_scatter")

    def scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
        assert dim == 0 and reduce == "sum" and out is None
        res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        return res.index_add_(0, index, src)

    ts.scatter = scatter
    ts.scatter_add = scatter
    ts.scatter_min = ts.scatter_max = ts.scatter_mean = None   # imported, never called
    sys.modules["torch_scatter"] = ts

    ed = types.ModuleType("easydict")

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __getattr__ = dict.__getitem__
        __setattr__ = __setitem__

    ed.EasyDict = EasyDict
    sys.modules["easydict"] = ed
    sys.path.insert(0, os.path.join(REF, "src"))


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def write_dataset_dir(path, N, R, train, valid, test):
    """A dataset directory in the reference's text format from integer ids."""
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "entities.dict"), "w") as f:
        for i in range(N):
            f.write("%d\te%d\n" % (i, i))
    with open(os.path.join(path, "relations.dict"), "w") as f:
        for i in range(R):
            f.write("%d\tr%d\n" % (i, i))
    for name, arr in (("train", train), ("valid", valid), ("test", test)):
        with open(os.path.join(path, name + ".txt"), "w") as f:
            for h, r, t in arr:
                f.write("e%d\tr%d\te%d\n" % (h, r, t))


def synthetic_graph(seed, N=300, half=6, E_half=1400, n_eval=160):
    """Small seeded graph WITH inverse relations at id+half (as in WN18RR / FB15k-237)."""
    rng = np.random.default_rng(seed)
    ew = (np.arange(N) + 1.0) ** -0.7
    ew /= ew.sum()
    perm = rng.permutation(N)
    seen = set()
    base = []
    while len(base) < E_half + 2 * n_eval:
        r = int(rng.integers(half))
        h = int(perm[rng.choice(N, p=ew)])
        t = int(perm[rng.choice(N, p=ew)])
        if h == t or (h, r, t) in seen:
            continue
        seen.add((h, r, t))
        base.append((h, r, t))
    tr, va, te = base[:E_half], base[E_half:E_half + n_eval], base[E_half + n_eval:]

    def with_inv(lst):
        out = []
        for h, r, t in lst:
            out.append((h, r, t))
            out.append((t, r + half, h))
        return np.array(out, dtype=np.int64)

    return N, 2 * half, with_inv(sorted(tr)), with_inv(va), with_inv(te)


def load_mined(path, stride, extra_seed, R, max_extra_len):
    rules = set()
    with open(path) as f:
        for line in f:
            toks = line.split()
            rules.add(tuple(int(v) for v in toks[:-1]))      # strip the trailing H column
    rules = sorted(rules, key=lambda r: (r[0], len(r), r))
    full = rules
    sub = rules[::stride]
    rng = np.random.default_rng(extra_seed)
    extra = []
    for head in range(R):                                      # empty bodies + longer bodies + duplicates
        extra.append((head,))
        for L in range(4, max_extra_len + 1):
            extra.append((head,) + tuple(int(v) for v in rng.integers(R, size=L)))
    extra += sub[:7]                                           # duplicate rules keep separate parameters
    return full, [list(r) for r in sub + extra]


def random_walk_rules(graph, R, seed, per_head, max_len):
    """Rules with support: bodies read off random walks in the train graph."""
    rng = np.random.default_rng(seed)
    out_edges = {}
    for h, r, t in graph.train_facts:
        out_edges.setdefault(h, []).append((r, t))
    rules = []
    for head in range(R):
        rules.append([head])
        facts = [f for f in graph.train_facts if f[1] == head]
        for _ in range(per_head):
            L = int(rng.integers(1, max_len + 1))
            e = facts[int(rng.integers(len(facts)))][0] if facts else int(rng.integers(graph.entity_size))
            body = []
            for _k in range(L):
                nb = out_edges.get(e)
                if not nb:
                    break
                r, e = nb[int(rng.integers(len(nb)))]
                body.append(r)
            rules.append([head] + body)
    rules += rules[3:9]
    return rules


def pick_batches(ds, n, seed):
    rng = np.random.default_rng(seed)
    idx = rng.choice(len(ds.batches), size=min(n, len(ds.batches)), replace=False)
    return [int(i) for i in sorted(idx)]


def sd_to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}     # copy: parameters are trained in place later


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mined-dir", default="/tmp/miner")
    args = ap.parse_args()
    install_standins()
    import data as rdata
    import predictors as rpred
    import trainer as rtrainer

    tmp = tempfile.mkdtemp(prefix="golden_")
    datasets = {}
    for name in ("umls", "kinship"):
        g = rdata.KnowledgeGraph(os.path.join(REF, "data", name))
        datasets[name] = (g, os.path.join(REF, "data", name))
    N, R, tr, va, te = synthetic_graph(7)
    sp = os.path.join(tmp, "syn")
    write_dataset_dir(sp, N, R, tr, va, te)
    datasets["syn"] = (rdata.KnowledgeGraph(sp), sp)

    for name, (g, path) in datasets.items():
        fx = {}
        N, R = g.entity_size, g.relation_size
        fx["N"], fx["R"] = np.int64(N), np.int64(R)
        fx["train"] = np.array(g.train_facts, dtype=np.int32)
        fx["valid"] = np.array(g.valid_facts, dtype=np.int32)
        fx["test"] = np.array(g.test_facts, dtype=np.int32)

        mined = os.path.join(args.mined_dir, name + "_mined.txt")
        if name in ("umls", "kinship"):
            full, rules = load_mined(mined, 9 if name == "umls" else 16, 11, R, 5)
            flat = np.full((len(full), 4), -1, dtype=np.int16)
            for i, r in enumerate(full):
                flat[i, :len(r)] = r
            fx["mined_rules"] = flat                            # all mined rules, pad -1
        else:
            rules = random_walk_rules(g, R, 3, 40, 5)
        lmax = max(len(r) for r in rules)
        flat = np.full((len(rules), lmax), -1, dtype=np.int16)
        for i, r in enumerate(rules):
            flat[i, :len(r)] = r
        fx["rules"] = flat

        # ---- batches (reference dataset builders, seeded) -------------------------------
        seed_all(1)
        bs = 32 if name != "syn" else 50                        # syn exercises B > 32
        train_set = rdata.TrainDataset(g, bs)
        valid_set = rdata.ValidDataset(g, bs)
        test_set = rdata.TestDataset(g, bs)
        tb = pick_batches(train_set, 10, 5)
        vb = pick_batches(valid_set, 6, 6)
        for tag, ds, ids in (("tb", train_set, tb), ("vb", valid_set, vb)):
            fx[tag + "_n"] = np.int64(len(ids))
            for j, i in enumerate(ids):
                fx["%s%d_triples" % (tag, j)] = np.array(ds.batches[i], dtype=np.int32)
                item = ds[i]
                if tag == "tb":
                    fx["tb%d_target" % j] = np.packbits(item[3].numpy().astype(bool), axis=1)
                    fx["tb%d_etr" % j] = item[4].numpy()
                else:
                    fx["vb%d_flag" % j] = np.packbits(item[3].numpy(), axis=1)

        # ---- (a) grounding counts ------------------------------------------------------
        pred = rpred.Predictor(g)
        pred.set_rules(rules)
        rng = np.random.default_rng(17)
        case = 0
        for j, i in enumerate(tb[:6]):
            all_h, all_r, all_t, target, etr = train_set[i]
            q = int(all_r[0])
            rq = pred.relation2rules[q]
            pick = rng.choice(len(rq), size=min(12, len(rq)), replace=False)
            # always include rules whose body contains the head relation (masking matters)
            with_head = [k for k, (_, (hd, body)) in enumerate(rq) if q in body][:6]
            for k in sorted(set(int(v) for v in pick) | set(with_head)):
                index, (hd, body) = rq[k]
                for use_etr in (True, False):
                    c = g.grounding(all_h, q, body, etr if use_etr else None)
                    fx["gr%d_batch" % case] = np.int64(j)
                    fx["gr%d_rule" % case] = np.int64(index)
                    fx["gr%d_etr" % case] = np.int64(use_etr)
                    fx["gr%d_counts" % case] = c.numpy()
                    case += 1
        fx["gr_n"] = np.int64(case)

        # ---- (b) Predictor forward / loss / grads / compute_H ----------------------------
        for ef in ("bias", "none"):
            seed_all(3)
            pred = rpred.Predictor(g, entity_feature=ef)
            pred.set_rules(rules)
            with torch.no_grad():
                pred.rule_weights.copy_(torch.randn(pred.num_rules) * 0.3)
                if ef == "bias":
                    pred.bias.copy_(torch.randn(N) * 0.1)
            fx["pred_%s_w" % ef] = pred.rule_weights.detach().numpy().copy()
            if ef == "bias":
                fx["pred_bias_b"] = pred.bias.detach().numpy().copy()
            for j, i in enumerate(tb[:5]):
                all_h, all_r, all_t, target, etr = train_set[i]
                pred.zero_grad()
                score, mask = pred(all_h, all_r, etr)
                fx["pred_%s_tb%d_score" % (ef, j)] = score.detach().numpy()
                fx["pred_%s_tb%d_mask" % (ef, j)] = mask.numpy()
                if mask.sum().item() != 0:
                    tgt = target * 0.2 + torch.nn.functional.one_hot(all_t, N) * 0.8
                    lp = (torch.softmax(score, dim=1) + 1e-8).log()
                    loss = -(lp[mask] * tgt[mask]).sum() / torch.clamp(tgt[mask].sum(), min=1)
                    loss.backward()
                    fx["pred_%s_tb%d_loss" % (ef, j)] = loss.detach().numpy()
                    fx["pred_%s_tb%d_gw" % (ef, j)] = pred.rule_weights.grad.numpy().copy()
                    if ef == "bias":
                        fx["pred_bias_tb%d_gb" % j] = pred.bias.grad.numpy().copy()
                if ef == "bias":
                    with torch.no_grad():
                        H, idx = pred.compute_H(all_h, all_r, all_t, etr)
                    if H is not None:
                        fx["pred_bias_tb%d_H" % j] = H.numpy()
                        fx["pred_bias_tb%d_Hidx" % j] = idx.numpy()
            for j, i in enumerate(vb[:4]):
                all_h, all_r, all_t, flag = valid_set[i]
                with torch.no_grad():
                    score, mask = pred(all_h, all_r, None)
                fx["pred_%s_vb%d_score" % (ef, j)] = score.numpy()
                fx["pred_%s_vb%d_mask" % (ef, j)] = mask.numpy()

        # ---- (c) PredictorPlus variants -----------------------------------------------------
        rot_path = os.path.join(REF, "data", name, "RotatE_50") if name != "syn" else None
        if name == "syn":
            rot_path = os.path.join(tmp, "syn_rot")
            os.makedirs(rot_path, exist_ok=True)
            D, gamma = 24, 6.0
            rr = np.random.default_rng(9)
            rngv = (gamma + 2.0) / D
            np.save(os.path.join(rot_path, "entity_embedding.npy"),
                    rr.uniform(-rngv, rngv, size=(N, 2 * D)).astype(np.float32))
            np.save(os.path.join(rot_path, "relation_embedding.npy"),
                    rr.uniform(-rngv, rngv, size=(R // 2, D)).astype(np.float32))
            import json
            with open(os.path.join(rot_path, "config.json"), "w") as f:
                json.dump({"hidden_dim": D, "gamma": gamma, "nentity": N}, f)
        variants = [("emb", "sum", "bias"), ("lstm", "sum", "bias"), ("emb", "pna", "bias"),
                    ("gru", "pna", "none"), ("rnn", "sum", "none")]
        if name in ("kinship", "syn"):       # RotatE tables for UMLS/Kinship have R rows, no inverses
            variants.append(("emb", "sum", "RotatE"))
        for vi, (typ, agg, ef) in enumerate(variants):
            seed_all(100 + vi)
            kw = dict(type=typ, num_layers=2 if typ != "lstm" else 3, hidden_dim=16,
                      entity_feature=ef, aggregator=agg)
            if ef == "RotatE":
                kw["embedding_path"] = rot_path
            model = rpred.PredictorPlus(g, **kw)
            model.set_rules(rules)
            if ef == "bias":
                with torch.no_grad():
                    model.bias.copy_(torch.randn(N) * 0.1)
            tag = "plus%d" % vi
            fx[tag + "_cfg"] = np.array([typ, agg, ef, str(kw["num_layers"])])
            if ef == "RotatE":
                fx[tag + "_gamma"] = np.float64(model.RotatE.gamma)
            for k, v in sd_to_np(model.state_dict()).items():
                fx["%s_sd_%s" % (tag, k)] = v
            for j, i in enumerate(tb[:3]):
                all_h, all_r, all_t, target, etr = train_set[i]
                model.zero_grad()
                score, mask = model(all_h, all_r, etr)
                fx["%s_tb%d_score" % (tag, j)] = score.detach().numpy()
                fx["%s_tb%d_mask" % (tag, j)] = mask.numpy()
                if mask.sum().item() != 0:
                    tgt = target * 0.2 + torch.nn.functional.one_hot(all_t, N) * 0.8
                    lp = (torch.softmax(score, dim=1) + 1e-8).log()
                    loss = -(lp[mask] * tgt[mask]).sum() / torch.clamp(tgt[mask].sum(), min=1)
                    loss.backward()
                    fx["%s_tb%d_loss" % (tag, j)] = loss.detach().numpy()
                    for pn, par in model.named_parameters():
                        if par.grad is not None:
                            fx["%s_tb%d_g_%s" % (tag, j, pn)] = par.grad.numpy().copy()
            for j, i in enumerate(vb[:2]):
                all_h, all_r, all_t, flag = valid_set[i]
                with torch.no_grad():
                    score, mask = model(all_h, all_r, None)
                fx["%s_vb%d_score" % (tag, j)] = score.numpy()
                fx["%s_vb%d_mask" % (tag, j)] = mask.numpy()

        # ---- (e) evaluate(): (L,H) rows + metrics through the reference trainer -------------
        for ef in ("bias", "none"):
            seed_all(3)
            pred = rpred.Predictor(g, entity_feature=ef)
            pred.set_rules(rules)
            with torch.no_grad():
                pred.rule_weights.copy_(torch.randn(pred.num_rules) * 0.3)
                if ef == "bias":
                    pred.bias.copy_(torch.randn(N) * 0.1)
            seed_all(1)
            train_set2 = rdata.TrainDataset(g, bs)
            valid_set2 = rdata.ValidDataset(g, bs)
            test_set2 = rdata.TestDataset(g, bs)
            solver = rtrainer.TrainerPredictor(pred, train_set2, valid_set2, test_set2,
                                               torch.optim.Adam(pred.parameters(), lr=0.005), gpus=None)
            captured = {}
            real_tensor = torch.tensor

            def spy(obj, *a, **k):
                if isinstance(obj, list) and obj and isinstance(obj[0], list) and len(obj[0]) == 5:
                    captured["rows"] = [[int(v) for v in row] for row in obj]
                return real_tensor(obj, *a, **k)

            for expectation in (True, False):
                stream = io.StringIO()
                handler = logging.StreamHandler(stream)
                logging.getLogger("").addHandler(handler)
                logging.getLogger("").setLevel(logging.INFO)
                torch.tensor = spy
                try:
                    mrr = solver.evaluate("valid", expectation=expectation)
                finally:
                    torch.tensor = real_tensor
                    logging.getLogger("").removeHandler(handler)
                vals = {}
                for line in stream.getvalue().splitlines():
                    if ":" in line and line.split(":")[0].strip() in ("Data", "Hit1", "Hit3", "Hit10", "MR", "MRR"):
                        vals[line.split(":")[0].strip()] = float(line.split(":")[1])
                fx["eval_%s_%d_mrr" % (ef, expectation)] = np.float64(mrr)
                fx["eval_%s_%d_logged" % (ef, expectation)] = np.array(
                    [vals["Data"], vals["Hit1"], vals["Hit3"], vals["Hit10"], vals["MR"], vals["MRR"]])
            fx["eval_%s_rows" % ef] = np.array(captured["rows"], dtype=np.int64)
            fx["eval_%s_batches" % ef] = np.array(
                [len(b) for b in valid_set2.batches], dtype=np.int64)
            fx["eval_%s_triples" % ef] = np.array(
                [tr_ for b in valid_set2.batches for tr_ in b], dtype=np.int32)

        # ---- (f) one short training run through the reference trainer (config-1 shape) ------
        if name in ("umls", "syn"):
            seed_all(1)
            train_set3 = rdata.TrainDataset(g, bs)
            valid_set3 = rdata.ValidDataset(g, bs)
            test_set3 = rdata.TestDataset(g, bs)
            model = rpred.PredictorPlus(g, type="emb", num_layers=3, hidden_dim=16,
                                        entity_feature="bias", aggregator="sum")
            model.set_rules(rules)
            for k, v in sd_to_np(model.state_dict()).items():
                fx["train_sd0_%s" % k] = v
            optim = torch.optim.Adam(model.parameters(), lr=0.005, weight_decay=0)
            solver = rtrainer.TrainerPredictor(model, train_set3, valid_set3, test_set3, optim, gpus=None)
            solver.train(batch_per_epoch=20, smoothing=0.2, print_every=1000)
            for k, v in sd_to_np(model.state_dict()).items():
                fx["train_sd1_%s" % k] = v
            fx["train_mrr_valid"] = np.float64(solver.evaluate("valid", expectation=True))

        np.savez_compressed(os.path.join(OUT, "golden_%s.npz" % name), **fx)
        print(name, "->", os.path.getsize(os.path.join(OUT, "golden_%s.npz" % name)) // 1024, "KiB,", len(fx), "arrays")


if __name__ == "__main__":
    main()
