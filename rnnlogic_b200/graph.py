"""Knowledge-graph storage for the B200 hot path.

Mirrors the interface of the reference ``KnowledgeGraph`` (src/data.py:9-173): same
constructor, same attributes (``entity_size``, ``relation_size``, ``train_facts``, ``hr2o`` ...,
``relation2ht2index``, ``encode_hr``, ``encode_ht``) and the same ``grounding(h, r, rule,
edges_to_remove) -> int64[B,N]`` operator -- but the adjacency lives on the GPU as a
relation-sorted DCSR by destination (include/rnnlogic_b200.h: rl_graph) and grounding runs the
hand-written frontier-expansion kernel.  Host-side construction is vectorised numpy.
"""
from __future__ import annotations

import ctypes as C
import os
import zlib
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def _answers_csr(N: int, parts: Sequence[np.ndarray]):
    """(keys, ptr, ent): sorted unique keys r*N+h with their de-duplicated tails."""
    tri = np.concatenate([p.reshape(-1, 3) for p in parts], axis=0).astype(np.int64)
    if tri.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(1, np.int32), np.zeros(0, np.int32)
    key = tri[:, 1] * N + tri[:, 0]
    pair = np.unique(np.stack([key, tri[:, 2]], axis=1), axis=0)
    keys, start = np.unique(pair[:, 0], return_index=True)
    ptr = np.concatenate([start, [pair.shape[0]]]).astype(np.int32)
    return keys.astype(np.int64), ptr, pair[:, 1].astype(np.int32)


def _load_triples_native(data_path, n_ent, n_rel):
    """(train, valid, test) int64[E,3] id arrays through the native loader (csrc_host/kg_loader.cpp),
    which replaces the per-line Python parse of src/data.py:43-99."""
    import ctypes
    H = _lib.host_lib()
    h = H.rl_kg_load(os.fsencode(data_path))
    if not h:
        raise KeyError(H.rl_kg_load_error().decode(errors="replace"))
    try:
        if H.rl_kg_num_entities(h) != n_ent or H.rl_kg_num_relations(h) != n_rel:
            raise ValueError("dictionary sizes disagree between the Python and the native loader")
        out = []
        for which in range(3):
            cnt = ctypes.c_int64()
            ptr = H.rl_kg_triples(h, which, ctypes.byref(cnt))
            n = int(cnt.value)
            out.append(np.ctypeslib.as_array(ptr, shape=(n * 3,)).copy().reshape(-1, 3) if n else np.zeros((0, 3), np.int64))
        return out
    finally:
        H.rl_kg_free(h)


class DeviceGraph:
    """Device-resident copy of the graph arrays + the C struct handed to the kernels."""

    def __init__(self, kg: "KnowledgeGraph", device: torch.device):
        self.device = device
        h = kg.host
        self.t = {k: torch.from_numpy(v if v.shape[0] else np.zeros(2, v.dtype)).to(device) for k, v in h.items()}
        t = self.t
        self.struct = _lib.RlGraph(
            kg.entity_size, kg.relation_size, kg.rank_words, int(h["row_dst"].shape[0]), int(h["edge_src"].shape[0]),
            t["dst_ptr"].data_ptr(), t["row_dst"].data_ptr(), t["row_start"].data_ptr(), t["edge_src"].data_ptr(),
            t["rank_tab"].data_ptr(), t["ord_ptr"].data_ptr(), t["ord_h"].data_ptr(), t["ord_t"].data_ptr(),
            t["fsrc_ptr"].data_ptr(), t["frow_start"].data_ptr(), t["fedge_dstrow"].data_ptr(),
            t["srank_tab"].data_ptr())
        self.answers = {}
        for which in ("hr2o", "hr2oo", "hr2ooo"):
            keys, ptr, ent = kg.answers_csr(which)
            tk, tp, te = (torch.from_numpy(a).to(device) for a in (keys, ptr, ent))
            self.answers[which] = (_lib.RlAnswers(int(keys.shape[0]), tk.data_ptr(), tp.data_ptr(), te.data_ptr()),
                                   (tk, tp, te))
        hs = np.zeros(kg.entity_size + 2, dtype=np.float64)
        hs[1:] = np.cumsum(1.0 / np.arange(1, kg.entity_size + 2, dtype=np.float64))
        self.harmonic = torch.from_numpy(hs).to(device)

    def ref(self):
        return C.byref(self.struct)


class KnowledgeGraph(object):
    _CACHE_VERSION = 1
    _FILES = ("entities.dict", "relations.dict", "train.txt", "valid.txt", "test.txt")

    def __init__(self, data_path: Optional[str] = None, *, entity_size: Optional[int] = None,
                 relation_size: Optional[int] = None, train=None, valid=None, test=None, cache: Optional[str] = None):
        """``KnowledgeGraph(data_path)`` reads entities.dict / relations.dict / {train,valid,test}.txt
        exactly like src/data.py:10-108.  The keyword form builds the same object from integer
        id arrays [E,3] of (h, r, t) (synthetic graphs, tests).
        cache (or $RNNLOGIC_B200_KG_CACHE): a directory for a binary image of the parsed triples and the
        device tables (DCSR, rank tables), keyed by the dataset files' sizes and mtimes -- a second
        construction then skips the text parse and every sort."""
        self.data_path = data_path
        self._cache_dir = cache or os.environ.get("RNNLOGIC_B200_KG_CACHE") or None
        self.entity2id, self.relation2id, self.id2entity, self.id2relation = {}, {}, {}, {}
        if data_path is not None:
            with open(os.path.join(data_path, "entities.dict")) as fi:
                for line in fi:
                    i, name = line.strip().split("\t")
                    self.entity2id[name] = int(i)
                    self.id2entity[int(i)] = name
            with open(os.path.join(data_path, "relations.dict")) as fi:
                for line in fi:
                    i, name = line.strip().split("\t")
                    self.relation2id[name] = int(i)
                    self.id2relation[int(i)] = name
            self.entity_size = len(self.entity2id)
            self.relation_size = len(self.relation2id)
            cached = self._cache_load()
            if cached is None:
                train, valid, test = _load_triples_native(data_path, self.entity_size, self.relation_size)
            else:
                train, valid, test = cached["train"], cached["valid"], cached["test"]
        else:
            cached = None
            self.entity_size = int(entity_size)
            self.relation_size = int(relation_size)
            train = np.asarray(train, dtype=np.int64).reshape(-1, 3)
            valid = np.zeros((0, 3), np.int64) if valid is None else np.asarray(valid, dtype=np.int64).reshape(-1, 3)
            test = np.zeros((0, 3), np.int64) if test is None else np.asarray(test, dtype=np.int64).reshape(-1, 3)
        self.train_array, self.valid_array, self.test_array = train, valid, test
        self._lists: Dict[str, list] = {}
        self._dicts: Dict[str, dict] = {}
        self._csr: Dict[str, tuple] = {}
        self._ht2index = None
        self._devices: Dict[str, DeviceGraph] = {}
        self._chains: "OrderedDict[tuple, object]" = OrderedDict()
        if cached is None:
            self._build_host()
            self._cache_store()
        else:
            self._adopt_host(cached)
        if data_path is not None:
            print("Data loading | DONE!")

    # ---- binary image of a parsed dataset ----
    def _cache_key(self):
        st = [os.stat(os.path.join(self.data_path, f)) for f in self._FILES]
        return np.array([self._CACHE_VERSION, self.entity_size, self.relation_size]
                        + [v for x in st for v in (x.st_size, x.st_mtime_ns)], dtype=np.int64)

    def _cache_path(self):
        tag = "%08x" % (zlib.crc32(os.path.abspath(self.data_path).encode()) & 0xFFFFFFFF)
        return os.path.join(self._cache_dir, "kg_%s_%s.npz" % (os.path.basename(os.path.normpath(self.data_path)), tag))

    def _cache_load(self):
        if not self._cache_dir or self.data_path is None:
            return None
        try:
            with np.load(self._cache_path()) as z:
                if not np.array_equal(z["key"], self._cache_key()):
                    return None                                  # a dataset file changed: rebuild
                return {k: z[k] for k in z.files}
        except (OSError, KeyError, ValueError):
            return None

    def _cache_store(self):
        if not self._cache_dir or self.data_path is None:
            return
        blob = {"key": self._cache_key(), "train": self.train_array, "valid": self.valid_array, "test": self.test_array,
                "train_edge_index": self.train_edge_index, "rel_edges": self.rel_edges, "rel_rows": self.rel_rows,
                "rel_sources": self.rel_sources}
        blob.update({"host_" + k: v for k, v in self.host.items()})
        try:
            os.makedirs(self._cache_dir, exist_ok=True)
            tmp = self._cache_path() + ".tmp%d.npz" % os.getpid()
            np.savez(tmp, **blob)
            os.replace(tmp, self._cache_path())                  # atomic: concurrent ranks never see a partial file
        except OSError:
            pass                                                 # read-only location: the cache is optional

    def _adopt_host(self, z):
        self.train_edge_index = z["train_edge_index"]
        self.rel_edges, self.rel_rows, self.rel_sources = z["rel_edges"], z["rel_rows"], z["rel_sources"]
        self.rank_words = (self.entity_size + 31) // 32
        self.host = {k[len("host_"):]: z[k] for k in z if k.startswith("host_")}

    # ---- reference-compatible attributes (lazy: Python lists/dicts are slow at 2e7 edges) ----
    def _facts(self, name):
        if name not in self._lists:
            arr = getattr(self, name + "_array")
            self._lists[name] = [tuple(row) for row in arr.tolist()]
        return self._lists[name]

    train_facts = property(lambda self: self._facts("train"))
    valid_facts = property(lambda self: self._facts("valid"))
    test_facts = property(lambda self: self._facts("test"))

    def _answer_dict(self, which):
        """data.py:49-61,79-99 -- insertion order of the reference (train, then valid, then test)."""
        if which not in self._dicts:
            parts = {"hr2o": ("train",), "hr2oo": ("train", "valid"), "hr2ooo": ("train", "valid", "test")}[which]
            d: Dict[int, List[int]] = {}
            for p in parts:
                for h, r, t in self._facts(p):
                    d.setdefault(self.encode_hr(h, r), []).append(t)
            self._dicts[which] = d
        return self._dicts[which]

    hr2o = property(lambda self: self._answer_dict("hr2o"))
    hr2oo = property(lambda self: self._answer_dict("hr2oo"))
    hr2ooo = property(lambda self: self._answer_dict("hr2ooo"))

    @property
    def relation2ht2index(self):
        """data.py:66-69: per relation, (t*N+h) -> position in the relation's train-order edge list."""
        if self._ht2index is None:
            out = [dict() for _ in range(self.relation_size)]
            for (h, r, t), k in zip(self._facts("train"), self.train_edge_index.tolist()):
                out[r][self.encode_ht(h, t)] = k
            self._ht2index = out
        return self._ht2index

    def encode_hr(self, h, r):
        return r * self.entity_size + h

    def decode_hr(self, index):
        return index % self.entity_size, index // self.entity_size

    def encode_ht(self, h, t):
        return t * self.entity_size + h

    def decode_ht(self, index):
        return index % self.entity_size, index // self.entity_size

    def edge_index_of(self, triples: np.ndarray) -> np.ndarray:
        """Position of each (h, r, t) train triple inside relation r's train-order edge list -- the
        ``edges_to_remove`` value of the reference's TrainDataset (data.py:214-216).  Vectorised
        binary search over the sorted train keys; raises KeyError for a triple not in train."""
        N = self.entity_size
        if getattr(self, "_train_keys", None) is None:
            tr = self.train_array
            key = (tr[:, 1] * N + tr[:, 2]) * N + tr[:, 0]
            o = np.argsort(key, kind="stable")
            self._train_keys, self._train_keys_idx = key[o], self.train_edge_index[o]
        tri = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
        key = (tri[:, 1] * N + tri[:, 2]) * N + tri[:, 0]
        pos = np.searchsorted(self._train_keys, key)
        pos = np.minimum(pos, self._train_keys.shape[0] - 1)
        if not np.array_equal(self._train_keys[pos], key):
            raise KeyError("triple not in train.txt")
        return self._train_keys_idx[pos]

    def answers_csr(self, which):
        if which not in self._csr:
            parts = {"hr2o": (self.train_array,), "hr2oo": (self.train_array, self.valid_array),
                     "hr2ooo": (self.train_array, self.valid_array, self.test_array)}[which]
            self._csr[which] = _answers_csr(self.entity_size, parts)
        return self._csr[which]

    # ---- host-side DCSR construction ---------------------------------------------------------
    def _build_host(self):
        N, R = self.entity_size, self.relation_size
        tr = self.train_array
        E = tr.shape[0]
        if E >= 2 ** 31 or N >= 2 ** 31:
            raise ValueError("graph too large for 32-bit indices")
        h, r, t = tr[:, 0], tr[:, 1], tr[:, 2]
        if E and (h.min() < 0 or h.max() >= N or t.min() < 0 or t.max() >= N or r.min() < 0 or r.max() >= R):
            raise ValueError("triple id out of range")
        # reference edge order: per relation, train.txt order (data.py:63-64)
        order = np.argsort(r, kind="stable")
        rel_sizes = np.bincount(r, minlength=R).astype(np.int64)
        ord_ptr = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(rel_sizes, out=ord_ptr[1:])
        self.train_edge_index = np.empty(E, dtype=np.int64)       # index of train fact i inside its relation
        self.train_edge_index[order] = np.arange(E) - ord_ptr[r[order]]
        # DCSR by destination: sort by (r, t, h) -- one sort of a packed key when it fits 63 bits
        packed = N * N * max(R, 1) < 2 ** 62
        srt = np.argsort((r * N + t) * N + h, kind="stable") if packed else np.lexsort((h, t, r))
        rs, ts, hs = r[srt], t[srt], h[srt]
        rowkey = rs * N + ts
        new_row = np.ones(E, dtype=bool)
        new_row[1:] = rowkey[1:] != rowkey[:-1]
        # duplicates are an input error in the reference (assert at data.py:67); after the sort they are neighbours
        if E > 1 and bool(np.any(~new_row[1:] & (hs[1:] == hs[:-1]))):
            raise AssertionError("duplicate train triple")
        row_first = np.flatnonzero(new_row)
        TR = row_first.shape[0]
        row_rel = rs[row_first]
        row_dst = ts[row_first]
        row_start = np.concatenate([row_first, [E]])
        dst_ptr = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(np.bincount(row_rel, minlength=R), out=dst_ptr[1:])
        # rank table {bits, rows-before-word} per (relation, 32-entity word)
        W = (N + 31) // 32
        bits = np.zeros(R * W, dtype=np.uint32)
        np.bitwise_or.at(bits, row_rel * W + (row_dst >> 5), (np.uint32(1) << (row_dst & 31).astype(np.uint32)))
        pop = np.bitwise_count(bits).astype(np.int64).reshape(R, W) if R * W else np.zeros((R, W), np.int64)
        prefix = np.cumsum(pop, axis=1) - pop
        rank_tab = np.empty((R * W, 2), dtype=np.uint32)
        rank_tab[:, 0] = bits
        rank_tab[:, 1] = prefix.reshape(-1).astype(np.uint32)
        # forward DCSR by source: edges sorted by (r, h, t); each out-edge stores the LOCAL row of its tail
        fs = np.argsort((r * N + h) * N + t, kind="stable") if packed else np.lexsort((t, h, r))
        rf, hf, tf = r[fs], h[fs], t[fs]
        skey = rf * N + hf
        new_src = np.ones(E, dtype=bool)
        new_src[1:] = skey[1:] != skey[:-1]
        src_first = np.flatnonzero(new_src)
        src_rel, src_ent = rf[src_first], hf[src_first]
        frow_start = np.concatenate([src_first, [E]])
        fsrc_ptr = np.zeros(R + 1, dtype=np.int64)
        np.cumsum(np.bincount(src_rel, minlength=R), out=fsrc_ptr[1:])
        # local destination row of every edge: position of (r,t) among the relation's rows
        fedge_dstrow = np.searchsorted(rs[row_first] * N + row_dst, rf * N + tf) - dst_ptr[rf]
        sbits = np.zeros(R * W, dtype=np.uint32)
        np.bitwise_or.at(sbits, src_rel * W + (src_ent >> 5), (np.uint32(1) << (src_ent & 31).astype(np.uint32)))
        spop = np.bitwise_count(sbits).astype(np.int64).reshape(R, W) if R * W else np.zeros((R, W), np.int64)
        sprefix = np.cumsum(spop, axis=1) - spop
        srank_tab = np.empty((R * W, 2), dtype=np.uint32)
        srank_tab[:, 0] = sbits
        srank_tab[:, 1] = sprefix.reshape(-1).astype(np.uint32)
        # per-relation statistics (algorithmic-bytes model)
        self.rel_edges = rel_sizes
        self.rel_rows = np.diff(dst_ptr)
        self.rel_sources = np.diff(fsrc_ptr)                      # distinct (relation, source) pairs
        self.rank_words = W
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.host = {
            "dst_ptr": i32(dst_ptr), "row_dst": i32(row_dst), "row_start": i32(row_start), "edge_src": i32(hs),
            "rank_tab": np.ascontiguousarray(rank_tab.reshape(-1)),
            "ord_ptr": i32(ord_ptr), "ord_h": i32(h[order]), "ord_t": i32(t[order]),
            "fsrc_ptr": i32(fsrc_ptr), "frow_start": i32(frow_start), "fedge_dstrow": i32(fedge_dstrow),
            "srank_tab": np.ascontiguousarray(srank_tab.reshape(-1)),
        }

    def device_graph(self, device) -> DeviceGraph:
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.RlError("rnnlogic_b200 is CUDA-only (sm_100a): no CPU fallback for grounding (got %s)" % device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = str(device)
        if key not in self._devices:
            self._devices[key] = DeviceGraph(self, device)
        return self._devices[key]

    # ---- a6: grounding operator, API of src/data.py:136-147 ------------------------------------
    def grounding(self, h, r, rule, edges_to_remove):
        """int64[B,N] path counts of body ``rule`` from each h[b]; the query's own edge
        (``edges_to_remove[b]``, an index into relation r's train-order edge list) is cut on the
        hops whose relation equals r.  Bit-exact with the reference; CUDA tensors only."""
        from .engine import ground_chain
        return ground_chain(self, h, int(r), [int(x) for x in rule], edges_to_remove)

    def propagate(self, x, relation, edges_to_remove=None):
        """One hop from an arbitrary frontier (src/data.py:149-173): x int64[N,B,D] -> int64[N,B,D] with
        out[t] = sum over relation edges (s -> t) of x[s]; ``edges_to_remove[b]`` (index into the relation's
        train-order edge list) drops that edge's message for query b.  CUDA tensors only, bit-exact."""
        from .engine import _stream
        _lib.require_cuda(x, "x")
        if x.dim() != 3 or x.shape[0] != self.entity_size:
            raise ValueError("propagate expects x of shape [num_entities, B, D]")
        N, B, D = x.shape
        dg = self.device_graph(x.device)
        xi = x.to(torch.int64).contiguous()
        etr = None
        if edges_to_remove is not None:
            etr = edges_to_remove.to(x.device, torch.int64).contiguous()
            if etr.numel() != B:
                raise ValueError("edges_to_remove has %d entries for %d queries" % (etr.numel(), B))
            if D != 1:                                        # data.py:165-169 indexes message.view(-1, D) rows
                etr = etr.repeat_interleave(D)
        out = torch.empty_like(xi)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().rl_propagate_dense(dg.ref(), int(relation), int(B * D), xi.data_ptr(),
                                                     etr.data_ptr() if etr is not None else None, out.data_ptr(),
                                                     _stream()), "rl_propagate_dense")
        return out.to(x.dtype)
