"""Golden vectors for the rule-discovery path: the rule sets the REFERENCE's own C++ miner
(RuleMiner::search, /root/reference/miner/rnnlogic.cpp:505-589, compiled into oracle/_ref by oracle/Makefile)
mines from the golden datasets with max_length 3 (the README's setting) -> tests/golden/mined_<name>.npz
(rules as rows [head, b1, b2, b3] padded with -1).  Run in the build container (needs /root/reference):
    make -C oracle ref && python tests/golden/make_mined_golden.py"""
import ctypes
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import _golden as G   # noqa: E402


def reference_mined_rules(fx, max_length=3, threads=4):
    N, R = int(fx["N"]), int(fx["R"])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_miner.so"))
    lib.ref_kg_new.restype = ctypes.c_void_p
    lib.ref_mine_rules.restype = ctypes.c_longlong
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "entities.dict"), "w").write("".join("%d\te%d\n" % (i, i) for i in range(N)))
        open(os.path.join(d, "relations.dict"), "w").write("".join("%d\tr%d\n" % (i, i) for i in range(R)))
        for split in ("train", "valid", "test"):
            open(os.path.join(d, split + ".txt"), "w").write("".join("e%d\tr%d\te%d\n" % tuple(x) for x in fx[split].tolist()))
        kg = ctypes.c_void_p(lib.ref_kg_new(d.encode()))
    cap = 1 << 24
    buf = (ctypes.c_int * cap)()
    n = lib.ref_mine_rules(kg, max_length, threads, buf, ctypes.c_longlong(cap))
    assert n <= cap
    flat = np.frombuffer(buf, dtype=np.int32, count=n)
    rules, i = [], 0
    while i < n:
        head, ln = int(flat[i]), int(flat[i + 1])
        rules.append([head] + [int(v) for v in flat[i + 2:i + 2 + ln]])
        i += 2 + ln
    lib.ref_kg_free(kg)
    return rules


if __name__ == "__main__":
    for name in G.DATASETS:
        fx = G.load(name)
        rules = reference_mined_rules(fx)
        arr = np.full((len(rules), 4), -1, dtype=np.int16)
        for k, r in enumerate(rules):
            arr[k, :len(r)] = r
        np.savez_compressed(os.path.join(G.GOLDEN, "mined_%s.npz" % name), rules=arr, max_length=np.int64(3))
        print(name, len(rules), "rules; lengths", np.bincount([len(r) - 1 for r in rules]).tolist())
