"""Helpers shared by the tests: load tests/golden/*.npz (outputs of the unmodified reference,
see tests/golden/make_golden.py) and rebuild the integer inputs they were produced from."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
DATASETS = ("umls", "kinship", "syn")
_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, "golden_%s.npz" % name)))
    return _cache[name]


def rules_of(fx, key="rules"):
    return [[int(v) for v in row if v >= 0] for row in fx[key]]


def train_batch_inputs(fx, j):
    tri = fx["tb%d_triples" % j].astype(np.int64)
    N = int(fx["N"])
    target = np.unpackbits(fx["tb%d_target" % j], axis=1)[:, :N].astype(np.float32)
    return tri, torch.from_numpy(target), torch.from_numpy(fx["tb%d_etr" % j].astype(np.int64))


def valid_batch_inputs(fx, j):
    tri = fx["vb%d_triples" % j].astype(np.int64)
    N = int(fx["N"])
    flag = np.unpackbits(fx["vb%d_flag" % j], axis=1)[:, :N].astype(bool)
    return tri, torch.from_numpy(flag)


def plus_state(fx, tag):
    pre = tag + "_sd_"
    return {k[len(pre):]: torch.from_numpy(v.copy()) for k, v in fx.items() if k.startswith(pre)}


def plus_cfg(fx, tag):
    typ, agg, ef, nl = [str(v) for v in fx[tag + "_cfg"]]
    cfg = dict(type=typ, aggregator=agg, entity_feature=ef, num_layers=int(nl), hidden_dim=16)
    if ef == "RotatE":
        cfg["gamma"] = float(fx[tag + "_gamma"])
    return cfg


def plus_tags(fx):
    return sorted(k[:-4] for k in fx if k.startswith("plus") and k.endswith("_cfg"))
