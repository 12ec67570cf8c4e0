"""Rule compiler: rule list -> per-head prefix tries + work tables (include/rnnlogic_b200.h:
rl_rules).  Replaces the ``relation2rules`` table of the reference predictors
(src/predictors.py:46-49, 186-189) for the GPU kernels.

Sharing body prefixes is exact: the edge mask of a hop depends only on (head, hop relation)
(src/data.py:143-146), so two rules of one head with the same prefix have the same frontier.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

LANES = _lib.LANES


def parse_rules(inp) -> List[Tuple[int, List[int]]]:
    """list[[head, body...]] or a file of whitespace-separated ints (predictors.py:27-43)."""
    rules: List[Tuple[int, List[int]]] = []
    if type(inp) == list:
        for rule in inp:
            rules.append((int(rule[0]), [int(v) for v in rule[1:]]))
    elif type(inp) == str:
        with open(inp, "r") as fi:
            for line in fi:
                toks = [int(v) for v in line.strip().split()]
                rules.append((toks[0], toks[1:]))
    else:
        raise ValueError
    return rules


class CompiledRules:
    def __init__(self, graph, rules: Sequence[Tuple[int, Sequence[int]]]):
        self.graph = graph
        R = graph.relation_size
        self.num_rules = len(rules)
        self.max_len = max([len(b) for _, b in rules] + [1])
        rel_rows = graph.rel_rows
        # ---- tries: native compiler (csrc_host/kg_loader.cpp: rl_compile_tries) ---------------
        # nodes numbered by head, then depth, then first appearance in the rule list
        n = self.num_rules
        heads_a = np.fromiter((int(h) for h, _ in rules), dtype=np.int64, count=n)
        body_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.fromiter((len(b) for _, b in rules), dtype=np.int64, count=n), out=body_ptr[1:])
        total = int(body_ptr[-1])
        body_a = np.fromiter((int(x) for _, b in rules for x in b), dtype=np.int64, count=total)
        cap = max(1, total)
        node_rel, node_parent = np.empty(cap, np.int64), np.empty(cap, np.int64)
        node_depth, node_head = np.empty(cap, np.int64), np.empty(cap, np.int64)
        head_node_ptr = np.zeros(R + 1, dtype=np.int64)
        rule_node = np.full(max(1, n), -1, dtype=np.int64)
        vp_ = lambda a: a.ctypes.data_as(C.c_void_p)
        nn = int(_lib.host_lib().rl_compile_tries(n, R, vp_(heads_a), vp_(body_ptr), vp_(body_a), vp_(node_rel), vp_(node_parent),
                                                  vp_(node_depth), vp_(node_head), vp_(head_node_ptr), vp_(rule_node)))
        if nn < 0:
            raise ValueError("rule %d uses a relation id outside [0, %d)" % (-1 - nn, R))
        self.num_nodes = nn
        node_rel, node_parent, node_depth, node_head = node_rel[:nn], node_parent[:nn], node_depth[:nn], node_head[:nn]
        rule_node = rule_node[:n]
        zero_rules: List[List[int]] = [[] for _ in range(R)]
        for idx in np.flatnonzero(rule_node < 0):
            zero_rules[int(heads_a[idx])].append(int(idx))
        node_rows = rel_rows[node_rel] if self.num_nodes else np.zeros(0, np.int64)
        # arena rows per head and per-node offset inside the head's arena
        csum = np.zeros(self.num_nodes + 1, dtype=np.int64)
        np.cumsum(node_rows, out=csum[1:])
        head_row_base = csum[head_node_ptr[:-1]]
        node_row_off = csum[:-1] - (head_row_base[node_head] if self.num_nodes else 0)
        self.head_rows = csum[head_node_ptr[1:]] - csum[head_node_ptr[:-1]]
        self.head_nodes = np.diff(head_node_ptr)
        # chunks of <= 32 rows
        nch = (node_rows + LANES - 1) // LANES
        chunk_node = np.repeat(np.arange(self.num_nodes), nch)
        cstart = np.zeros(self.num_nodes + 1, dtype=np.int64)
        np.cumsum(nch, out=cstart[1:])
        chunk_row0 = (np.arange(chunk_node.shape[0]) - cstart[chunk_node]) * LANES
        self.num_chunks = int(chunk_node.shape[0])
        L1 = self.max_len + 1
        # lvl_ptr[q, d] = first chunk of the first node with (head q, depth > d)
        nkey = node_head * (L1 + 1) + node_depth                     # ascending in node order
        want = (np.arange(R)[:, None] * (L1 + 1) + np.arange(1, L1 + 1)[None, :]).reshape(-1)
        first_node = np.searchsorted(nkey, want, side="left")
        lvl_ptr = cstart[first_node].reshape(R, L1)
        self.level_chunks = np.diff(lvl_ptr, axis=1)                  # [R, max_len] chunks per (head, depth)
        self.level_nodes = np.diff(first_node.reshape(R, L1), axis=1)  # [R, max_len] nodes per (head, depth)
        self.head_chunks = lvl_ptr[:, -1] - lvl_ptr[:, 0]
        # rule ends: (rule id, node) pairs of the rules with a non-empty body
        t_rule = np.flatnonzero(rule_node >= 0).astype(np.int64)
        t_node = rule_node[t_rule]
        self.num_terms = int(t_rule.shape[0])
        self.head_terms = np.bincount(node_head[t_node], minlength=R) if t_rule.shape[0] else np.zeros(R, np.int64)
        self.rule_node = np.full(self.num_rules, -1, dtype=np.int64)
        self.rule_node[t_rule] = t_node
        node_nterm = np.bincount(t_node, minlength=max(1, self.num_nodes))
        # rules ending at a node (CSR by node) and the item capacity of a head = rows of its rule-end nodes
        o_n = np.lexsort((t_rule, t_node)) if t_rule.shape[0] else np.zeros(0, np.int64)
        node_term_rule = t_rule[o_n] if t_rule.shape[0] else np.zeros(0, np.int64)
        node_term_ptr = np.zeros(max(1, self.num_nodes) + 1, dtype=np.int64)
        np.cumsum(node_nterm, out=node_term_ptr[1:])
        self.rule_nterm = np.full(self.num_rules, -1, dtype=np.int64)      # position of a rule in node_term_rule (-1: empty body)
        self.rule_nterm[node_term_rule] = np.arange(node_term_rule.shape[0])
        self.head_nterm0 = node_term_ptr[np.minimum(head_node_ptr[:-1], max(1, self.num_nodes))]   # first term of each head
        term_rows = np.where(node_nterm[:self.num_nodes] > 0, node_rows, 0) if self.num_nodes else np.zeros(0, np.int64)
        tr_sum = np.zeros(self.num_nodes + 1, dtype=np.int64)
        np.cumsum(term_rows, out=tr_sum[1:])
        self.head_item_cap = tr_sum[head_node_ptr[1:]] - tr_sum[head_node_ptr[:-1]]
        zr_ptr = np.zeros(R + 1, dtype=np.int64)
        np.cumsum([len(z) for z in zero_rules], out=zr_ptr[1:])
        zr_rule = np.array([i for z in zero_rules for i in z], dtype=np.int64)
        # rule ids per head in rule-file order (reference relation2rules order)
        self.rules = list(rules)
        self.head_rules: List[List[int]] = [[] for _ in range(R)]
        for idx, (head, _) in enumerate(rules):
            self.head_rules[head].append(idx)
        self.head_rule_array = [np.asarray(v, dtype=np.int64) for v in self.head_rules]
        # algorithmic bytes per head for a 32-lane slot: c = i = 4 bytes
        E, D, U = graph.rel_edges, graph.rel_rows, graph.rel_sources
        per_node = 4 * E[node_rel] + 8 * D[node_rel] + 4 * LANES * ((node_depth > 1) * U[node_rel] + D[node_rel]) \
            if self.num_nodes else np.zeros(0, np.int64)
        nb = np.zeros(self.num_nodes + 1, dtype=np.int64)
        np.cumsum(per_node, out=nb[1:])
        self.head_ground_bytes = nb[head_node_ptr[1:]] - nb[head_node_ptr[:-1]]
        self.node_depth, self.node_rel_host, self.node_head = node_depth, node_rel, node_head
        self.head_node_ptr = head_node_ptr
        # packed node records (one 32-byte load in the kernels)
        has_p = node_parent >= 0
        psafe = np.where(has_p, node_parent, 0)
        rec = np.zeros((max(1, self.num_nodes), 8), dtype=np.int64)
        if self.num_nodes:
            dst_ptr = np.concatenate([[0], np.cumsum(rel_rows)])
            rec[:, 0] = node_rel
            rec[:, 1] = node_parent
            rec[:, 2] = np.where(has_p, node_rel[psafe], -1)
            rec[:, 3] = dst_ptr[node_rel]
            rec[:, 4] = node_rows
            rec[:, 5] = cstart[:-1]
            rec[:, 6] = np.where(has_p, cstart[:-1][psafe], 0)
            rec[:, 7] = node_nterm[:self.num_nodes]
        node_prow_off = np.where(has_p, node_row_off[psafe], 0) if self.num_nodes else np.zeros(0, np.int64)
        # distinct (parent relation, relation) hops: each gets a pair table on the device (DeviceRules)
        if self.num_nodes and has_p.any():
            pkey = np.where(has_p, node_rel[psafe] * R + node_rel, -1)
            ukeys, inv = np.unique(pkey[has_p], return_inverse=True)
            self.pair_prel, self.pair_rel = ukeys // R, ukeys % R
            pbase = np.zeros(ukeys.shape[0] + 1, dtype=np.int64)
            np.cumsum(rel_rows[self.pair_rel] + 1, out=pbase[1:])
            self.pair_base = pbase
            node_pair_off = np.full(self.num_nodes, -1, dtype=np.int64)
            node_pair_off[has_p] = pbase[:-1][inv]
        else:
            self.pair_prel = self.pair_rel = np.zeros(0, np.int64)
            self.pair_base = np.zeros(1, dtype=np.int64)
            node_pair_off = np.full(max(1, self.num_nodes), -1, dtype=np.int64)
        self.node_pair_off = node_pair_off
        # symbolic work items: one per 32 parent bitmap words (1024 parent rows); one per node at depth 1
        if self.num_nodes:
            pwords = np.where(has_p, (node_rows[psafe] + 31) // 32, 1)
            n_items = np.maximum(1, (pwords + 31) // 32)
            sym_node = np.repeat(np.arange(self.num_nodes), n_items)
            istart = np.zeros(self.num_nodes + 1, dtype=np.int64)
            np.cumsum(n_items, out=istart[1:])
            sym_w0 = (np.arange(sym_node.shape[0]) - istart[sym_node]) * 32
            lvl_sym_ptr = istart[first_node].reshape(R, L1)
        else:
            sym_node = sym_w0 = np.zeros(0, np.int64)
            lvl_sym_ptr = np.zeros((R, L1), dtype=np.int64)
        self.level_sym_items = np.diff(lvl_sym_ptr, axis=1)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        if self.num_chunks >= 2 ** 31 or self.num_nodes >= 2 ** 31:
            raise ValueError("rule set too large for 32-bit work tables")
        if self.num_nodes and int(self.head_item_cap.max()) >= 2 ** 26:
            raise ValueError("a head's rule-end nodes have >= 2^26 rows: item indices no longer fit the coordinate keys")
        self.host = {
            "node_rel": i32(node_rel), "node_parent": i32(node_parent),
            "node_row_off": np.ascontiguousarray(node_row_off, dtype=np.int64),
            "head_node_ptr": i32(head_node_ptr), "lvl_ptr": i32(lvl_ptr.reshape(-1)),
            "chunk_node": i32(chunk_node), "chunk_row0": i32(chunk_row0), "zr_ptr": i32(zr_ptr), "zr_rule": i32(zr_rule),
            "node_chunk0": i32(cstart[:-1]),
            "node_rec": i32(rec.reshape(-1)), "node_prow_off": np.ascontiguousarray(node_prow_off, dtype=np.int64),
            "lvl_sym_ptr": i32(lvl_sym_ptr.reshape(-1)), "sym_node": i32(sym_node), "sym_w0": i32(sym_w0),
            "node_term_ptr": i32(node_term_ptr), "node_term_rule": i32(node_term_rule),
            "node_pair_off": np.ascontiguousarray(node_pair_off, dtype=np.int64),
        }
        self._devices = {}

    def device_rules(self, device) -> "DeviceRules":
        key = str(device)
        if key not in self._devices:
            self._devices[key] = DeviceRules(self, device)
        return self._devices[key]


class DeviceRules:
    def __init__(self, cr: CompiledRules, device):
        self.device = device
        # zero-length arrays still need a valid (non-null) pointer for the arg checks
        self.t = {k: torch.from_numpy(v if v.shape[0] else np.zeros(1, v.dtype)).to(device) for k, v in cr.host.items()}
        t = self.t
        self._build_pair_tables(cr, device)
        self.struct = _lib.RlRules(
            cr.num_nodes, cr.num_rules, cr.max_len, cr.num_chunks, cr.num_terms, int(cr.host["zr_rule"].shape[0]),
            t["node_rel"].data_ptr(), t["node_row_off"].data_ptr(),
            t["head_node_ptr"].data_ptr(), t["lvl_ptr"].data_ptr(), t["chunk_node"].data_ptr(),
            t["chunk_row0"].data_ptr(), t["zr_ptr"].data_ptr(), t["zr_rule"].data_ptr(),
            t["node_chunk0"].data_ptr(),
            t["node_rec"].data_ptr(), t["node_prow_off"].data_ptr(), t["lvl_sym_ptr"].data_ptr(),
            t["sym_node"].data_ptr(), t["sym_w0"].data_ptr(), t["node_term_ptr"].data_ptr(),
            t["node_term_rule"].data_ptr(), t["node_pair_off"].data_ptr(), t["pair_ptr"].data_ptr(), t["pair_ent"].data_ptr())

    def _build_pair_tables(self, cr: CompiledRules, device):
        """(parent relation, relation) -> per destination row, the parent rows with an edge into it (rl_pair_table):
        count on the device, prefix-sum, fill.  Query-independent, built once per rule set and device."""
        from .engine import _stream
        t = self.t
        P = int(cr.pair_prel.shape[0])
        total = int(cr.pair_base[-1])
        if P == 0 or torch.device(device).type != "cuda":
            t["pair_ptr"] = torch.zeros(2, dtype=torch.int32, device=device)
            t["pair_ent"] = torch.zeros(1, dtype=torch.int32, device=device)
            return
        dg = cr.graph.device_graph(device)
        with torch.cuda.device(device):
            prel = torch.from_numpy(cr.pair_prel.astype(np.int32)).to(device)
            rel = torch.from_numpy(cr.pair_rel.astype(np.int32)).to(device)
            base = torch.from_numpy(cr.pair_base).to(device)
            cnt = torch.empty(total, dtype=torch.int32, device=device)
            _lib.check(_lib.lib().rl_pair_table(dg.ref(), P, prel.data_ptr(), rel.data_ptr(), base.data_ptr(), None,
                                                cnt.data_ptr(), _stream()), "rl_pair_table")
            csum = torch.cumsum(cnt, 0, dtype=torch.int64)
            n_ent = int(csum[-1].item())
            if n_ent >= 2 ** 31:
                raise ValueError("pair tables need %d entries: 32-bit offsets exhausted" % n_ent)
            ptr = (csum - cnt).to(torch.int32)                      # exclusive prefix sum
            del csum, cnt
            ent = torch.empty(max(1, n_ent), dtype=torch.int32, device=device)
            _lib.check(_lib.lib().rl_pair_table(dg.ref(), P, prel.data_ptr(), rel.data_ptr(), base.data_ptr(), ptr.data_ptr(),
                                                ent.data_ptr(), _stream()), "rl_pair_table")
            torch.cuda.current_stream().synchronize()                # prel / rel / base die here
        t["pair_ptr"], t["pair_ent"] = ptr, ent
        self.pair_entries = n_ent

    def ref(self):
        return C.byref(self.struct)
