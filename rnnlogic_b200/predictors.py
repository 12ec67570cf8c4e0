"""Reasoning predictors with the reference's Python interface (src/predictors.py:17-271), running
on the hand-written sm_100a kernels.

``Predictor`` / ``PredictorPlus`` keep ``set_rules(list | path)``, ``forward(all_h, all_r,
edges_to_remove) -> (score fp32[B,N], mask bool[B,N])``, ``compute_H`` and the reference's
``state_dict`` keys.  They are CUDA-only: a CPU tensor raises (no fallback)."""
from __future__ import annotations

import logging
import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .engine import Grounder
from .hotpath import ScoreKernels
from .rules import CompiledRules, parse_rules

LANES = _lib.LANES


class _RuleModel(nn.Module):
    """Shared plumbing: rule parsing/compilation and per-device kernel drivers."""

    force_dense = False      # True: expand every row of every trie node (plain dense SpMM)

    def _load_rules(self, input, who):
        rules = parse_rules(input)
        logging.info("%s: read %d rules from %s." % (who, len(rules), "list" if type(input) == list else "file"))
        self.rules = [(h, b) for h, b in rules]
        self.num_rules = len(self.rules)
        self.relation2rules = [[] for _ in range(self.num_relations)]
        for index, rule in enumerate(self.rules):
            self.relation2rules[rule[0]].append([index, rule])
        self.compiled = CompiledRules(self.graph, self.rules)
        self._drivers = {}

    def _driver(self, device) -> ScoreKernels:
        if device.type != "cuda":
            raise _lib.RlError("rnnlogic_b200 predictors are CUDA-only (sm_100a); got a %s batch. There is no CPU "
                               "fallback: move the model and the batch to a CUDA device." % device)
        key = str(device)
        if key not in self._drivers:
            self._drivers[key] = ScoreKernels(Grounder(self.graph, self.compiled, device, self.force_dense))
        return self._drivers[key]

    def _ground(self, all_h, all_r, edges_to_remove):
        query_r = all_r[0].item()
        assert (all_r != query_r).sum() == 0
        sk = self._driver(all_r.device)
        sl = sk.gr.make_slots([query_r], [all_h.size(0)], all_h, None, edges_to_remove)
        sk.gr.ground(sl)
        return query_r, sk, sl


class _PredictorScoreFn(torch.autograd.Function):
    """score = sum_rule w_rule * count_rule (+ bias): kernels (2a) forward / (2c) backward."""

    @staticmethod
    def forward(ctx, rule_weights, bias, sk, sl, fill_neg_inf):
        Z, nzmask = sk.predictor_scores(sl, rule_weights.detach().contiguous(),
                                        None if bias is None else bias.detach().contiguous(), fill_neg_inf)
        score, nz = sk.to_dense(sl, Z, nzmask)
        ctx.sk, ctx.sl, ctx.has_bias = sk, sl, bias is not None
        ctx.nrules = rule_weights.shape[0]
        ctx.mark_non_differentiable(nz)
        return score, nz

    @staticmethod
    def backward(ctx, gscore, _gnz):
        sk, sl = ctx.sk, ctx.sl
        G = sk.from_dense(sl, gscore.contiguous())
        grad_w = torch.zeros(ctx.nrules, dtype=torch.float32, device=gscore.device)
        grad_b = torch.zeros(sk.N, dtype=torch.float32, device=gscore.device) if ctx.has_bias else None
        sk.predictor_backward(sl, G, None, grad_w, grad_b)
        return grad_w, grad_b, None, None, None


class Predictor(_RuleModel):
    """Linear rule-weight predictor (src/predictors.py:17-119)."""

    def __init__(self, graph, entity_feature="bias"):
        super(Predictor, self).__init__()
        self.graph = graph
        self.num_entities = graph.entity_size
        self.num_relations = graph.relation_size
        self.entity_feature = entity_feature
        if entity_feature == "bias":
            self.bias = nn.parameter.Parameter(torch.zeros(self.num_entities))

    def set_rules(self, input):
        self._load_rules(input, "Predictor")
        self.rule_weights = nn.parameter.Parameter(torch.zeros(self.num_rules, device=self._param_device()))

    def _param_device(self):
        return self.bias.device if self.entity_feature == "bias" else torch.device("cpu")

    def forward(self, all_h, all_r, edges_to_remove):
        query_r, sk, sl = self._ground(all_h, all_r, edges_to_remove)
        use_bias = self.entity_feature == "bias"
        score, nz = _PredictorScoreFn.apply(self.rule_weights, self.bias if use_bias else None, sk, sl,
                                            not use_bias)
        if not bool(nz.any().item()):                                   # predictors.py:67-71
            if use_bias:
                return score, torch.ones_like(nz)
            return torch.full_like(score, float("inf")), torch.zeros_like(nz)
        if use_bias:
            return score, torch.ones_like(nz)
        return score, nz

    @torch.no_grad()
    def compute_H(self, all_h, all_r, all_t, edges_to_remove):
        """E-step rule scores (src/predictors.py:82-119)."""
        query_r, sk, sl = self._ground(all_h, all_r, edges_to_remove)
        ids = self.compiled.head_rules[query_r]
        if len(ids) == 0:
            return None, None
        device = all_r.device
        all_t = all_t.to(device)
        B, N = all_h.size(0), self.num_entities
        neg = torch.zeros(B, N, dtype=torch.bool, device=device)
        pos_cnt = torch.empty(len(ids), B, dtype=torch.float32, device=device)
        sum_cnt = torch.empty(len(ids), B, dtype=torch.float32, device=device)
        for c0 in range(0, len(ids), 64):
            chunk = ids[c0:c0 + 64]
            x = sk.gr.rule_counts(sl, chunk)                              # [k,B,N] int64
            neg |= (x != 0).any(0)
            xf = x.float()
            pos_cnt[c0:c0 + len(chunk)] = xf.gather(2, all_t.view(1, B, 1).expand(len(chunk), B, 1)).squeeze(2)
            sum_cnt[c0:c0 + len(chunk)] = xf.sum(2)
        w = self.rule_weights[torch.tensor(ids, device=device)].unsqueeze(1)
        pos_score = pos_cnt * w                                           # pos_index has exactly one entry per row
        neg_n = torch.clamp(neg.sum(1), min=1).unsqueeze(0)
        # (score * neg_index).sum(1): counts are zero outside neg_index, so the plain row sum is the same
        neg_score = sum_cnt * w / neg_n
        H = torch.softmax((pos_score - neg_score).t(), dim=-1).sum(0)
        return H, torch.tensor(ids, dtype=torch.long, device=device)


# ------------------------------------------------------------------------------------------------
# fused paths used by rnnlogic_b200.trainer (no dense [B,N] target / flag tensors, no autograd)
# ------------------------------------------------------------------------------------------------
def _group_ptr(sl, device):
    """DEVICE int32[n_groups+1] slot ranges of the reference batches, or None when every batch fits
    one slot (then slot == group)."""
    sizes = sl.group_sizes
    if all(n <= LANES for n in sizes):
        return None, len(sizes)
    ptr = np.zeros(len(sizes) + 1, dtype=np.int32)
    np.cumsum([(n + LANES - 1) // LANES for n in sizes], out=ptr[1:])
    return torch.from_numpy(ptr).to(device), len(sizes)


_POP8 = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.float32)


def _group_mask_sum(sl, nzmask, ng):
    """mask.sum() of each reference batch (trainer.py:87,96) from the per-entity lane bitmasks."""
    per_slot = _POP8.to(nzmask.device)[nzmask.view(torch.uint8).long()].view(sl.S, -1).sum(1)
    if ng == sl.S:
        return per_slot
    gid = torch.from_numpy(np.repeat(np.arange(ng), [(n + LANES - 1) // LANES for n in sl.group_sizes])).to(nzmask.device)
    return torch.zeros(ng, dtype=torch.float32, device=nzmask.device).index_add_(0, gid, per_slot)


def _predictor_step_on_slots(self, sk, sl, smoothing, grad_scale=1.0, bits=32):
    """Enqueue ground -> aggregate -> CE -> backward for prepared slots (no host sync).
    Returns the device tensors (loss[ng], tsum[ng], mask_sum[ng] | None, grad_w, grad_b)."""
    device = sk.device
    use_bias = self.entity_feature == "bias"
    gptr, ng = _group_ptr(sl, device)
    sk.gr._run(sl, bits)
    Z, nzmask = sk.predictor_scores(sl, self.rule_weights.detach(), self.bias.detach() if use_bias else None,
                                    not use_bias)
    loss, tsum, G = sk.softmax_ce(sl, Z, nzmask, smoothing, not use_bias, gptr, ng, want_grad=True)
    gw = torch.zeros_like(self.rule_weights)
    gb = torch.zeros_like(self.bias) if use_bias else None
    scale = None
    if grad_scale != 1.0:
        scale = torch.full((sl.S,), float(grad_scale), dtype=torch.float32, device=device)
    sk.predictor_backward(sl, G, scale, gw, gb)
    msum = None if use_bias else _group_mask_sum(sl, nzmask, ng)
    return loss, tsum, msum, gw, gb


def _predictor_fused_train(self, batches, smoothing, grad_scale=1.0):
    """One fused step over a list of single-relation train batches (trainer.py:68-93 for each):
    ground -> aggregate -> log(softmax+1e-8) CE -> backward.  Gradients of ``grad_scale * sum of
    the batch losses`` are ACCUMULATED into .grad.  Returns (loss[n_batches], target_sum[n_batches])
    as host float tensors -- one device->host read per step."""
    device = self.rule_weights.device
    sk = self._driver(device)
    use_bias = self.entity_feature == "bias"
    sl = sk.gr.make_slots_host(batches, with_etr=True)
    ng = len(batches)
    for bits in ((sk.gr.force_bits,) if sk.gr.force_bits else (32, 64)):
        loss, tsum, msum, gw, gb = _predictor_step_on_slots(self, sk, sl, smoothing, grad_scale, bits)
        parts = [loss, tsum] + ([msum] if msum is not None else [])
        host = torch.cat(parts + [sl.overflow.float()]).cpu()            # the step's one sync
        if host[-1].item() == 0 or bits == 64:
            break
    for p, g in ((self.rule_weights, gw), (self.bias if use_bias else None, gb)):
        if p is not None:
            p.grad = g if p.grad is None else p.grad.add_(g)
    self.last_h2d_bytes = sl.h2d_bytes
    self.last_d2h_bytes = int(host.numel() * 4)
    self.last_mask_sum = None if use_bias else host[2 * ng:3 * ng].tolist()
    return host[:ng], host[ng:2 * ng]


@torch.no_grad()
def _predictor_fused_rank(self, batches, split):
    """(L,H) int64[Q,2] of a list of single-relation eval batches (trainer.py:173,189-201)."""
    device = self.rule_weights.device
    sk = self._driver(device)
    use_bias = self.entity_feature == "bias"
    sl = sk.gr.make_slots_host(batches, with_etr=False)
    sk.gr.ground(sl)
    Z, nzmask = sk.predictor_scores(sl, self.rule_weights.detach(), self.bias.detach() if use_bias else None,
                                    not use_bias)
    LH = sk.filtered_rank(sl, Z, nzmask, "hr2oo" if split == "valid" else "hr2ooo", not use_bias)
    if not use_bias:
        # predictors.py:67-71 quirk: a batch with no candidate at all returns +inf logits and an
        # all-False mask -> every query of it ranks (1, N+1); nzmask already yields exactly that.
        pass
    return _valid_lanes(sl, LH)


def _valid_lanes(sl, LH):
    idx = np.concatenate([s * LANES + np.arange(n) for s, n in enumerate(sl.nq)])
    return LH[torch.from_numpy(idx).to(LH.device)]


Predictor.fused_train_step = _predictor_fused_train
Predictor.step_on_slots = _predictor_step_on_slots
Predictor.fused_rank = _predictor_fused_rank
