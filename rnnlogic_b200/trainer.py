"""TrainerPredictor with the reference's interface (src/trainer.py:10-289) on the fused B200 path.

Differences that do not change results at world size 1:
  * batches are read as integer triples from ``dataset.batches``; targets (data.py:207-212) and
    eval filters (data.py:250-254) come from the device-resident answer lists, so no dense
    [B,N] tensor is built on the host or copied per batch;
  * forward, loss and backward are the fused kernels; the optimizer stays torch (Adam);
  * evaluate() ranks each batch as it is scored instead of keeping the split's [Q,N] logits.
Multi-GPU: one process per GPU, batches sharded by DistributedSampler (as the reference), the KG
and rules replicated, ONE all-reduce per step of the flattened gradients (DDP-mean semantics,
trainer.py:56-60), and one gather of the (h,r,t,L,H) rows at the end of evaluate().  The exchange of step i
runs after the grounding of step i+1 was enqueued (same stream: it does not overlap it; the point of the
early enqueue is that the GPU is never idle while the host reads step i back)."""
import logging
import os
from itertools import islice

import numpy as np
import torch
from torch import distributed as dist
from torch.utils import data as torch_data

from . import _lib, comm


def shard_indices(n_items, world_size, rank, shuffle=True, seed=0, epoch=0):
    """The batch indices DistributedSampler(dataset, world_size, rank) yields (trainer.py:52,
    :150): shuffled with torch.Generator(seed+epoch), padded by wrap-around to a multiple of
    world_size, strided by rank."""
    class _Sized:
        def __len__(self):
            return n_items
    sampler = torch_data.DistributedSampler(_Sized(), world_size, rank, shuffle=shuffle, seed=seed)
    sampler.set_epoch(epoch)
    return list(iter(sampler))


def allreduce_mean_grads(params, world_size, local_step=True):
    """The path's one collective: a single flat all-reduce of every gradient, then / world_size
    (DDP mean semantics, trainer.py:56-60).  A usage flag per parameter rides along so that a
    parameter no rank used keeps grad=None (find_unused_parameters=True behaviour), and so does this rank's
    "my batch had candidates" flag: the return value says whether ANY rank wants an optimizer step, so that
    every rank takes the same decision and the replicas never diverge (the reference skips the step per rank,
    trainer.py:87, and would dead-lock DDP there)."""
    if world_size == 1 or not params:
        return bool(local_step)
    device = params[0].device
    used = torch.tensor([0.0 if p.grad is None else 1.0 for p in params] + [1.0 if local_step else 0.0], device=device)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params]
                     + [used])
    comm.all_reduce_sum_(flat)
    n_par = len(params)
    flat[:-(n_par + 1)] /= world_size
    used = flat[-(n_par + 1):].cpu().tolist()
    off = 0
    for p, u in zip(params, used[:-1]):
        n = p.numel()
        if u > 0:
            p.grad = flat[off:off + n].view_as(p).clone()
        off += n
    return used[-1] > 0


def snake_deal(costs, world_size, rank):
    """Indices of the items one rank takes when ``costs`` are dealt largest-first in snake order
    (0,1,..,W-1,W-1,..,1,0,0,1,..): the ranks' shares are disjoint, cover everything, have the same size when
    len(costs) is a multiple of the world size, and nearly equal total cost -- so a per-step gradient
    exchange does not wait for one unlucky rank.  (A throughput-mode helper; the reference's own schedule is
    shard_indices.)"""
    order = sorted(range(len(costs)), key=lambda j: (-costs[j], j))
    period = 2 * world_size
    return [j for k, j in enumerate(order) if k % period == rank or k % period == period - 1 - rank]


class TrainerPredictor(object):
    slots_per_step = 1      # train batches per optimizer step (1 = the reference's schedule)
    pipelined = True        # enqueue the grounding of step i+1 before reading step i back (flags are checked before
                            # every optimizer step; an overflowed step is redone with 64-bit rows / larger arrays)
    eval_batches_per_call = 64
    use_graphs = True       # slots_per_step == 1: replay one captured CUDA graph per head relation for the whole fused step

    def __init__(self, model, train_set, valid_set, test_set, optimizer, scheduler=None, gpus=None, num_worker=0):
        self.rank = comm.get_rank()
        self.world_size = comm.get_world_size()
        self.gpus = gpus
        self.num_worker = num_worker
        if gpus is None:
            if not torch.cuda.is_available():
                raise _lib.RlError("rnnlogic_b200 is CUDA-only: `gpus: null` (CPU training, trainer.py:18-19) "
                                   "has no fallback here and no CUDA device is visible")
            logging.warning("gpus is None: the B200 path has no CPU mode, using cuda:0")
            gpus = [0] * self.world_size
        if len(gpus) != self.world_size:
            error_msg = "World size is %d but found %d GPUs in the argument"
            if self.world_size == 1:
                error_msg += ". Did you launch with `python -m torch.distributed.launch`?"
            raise ValueError(error_msg % (self.world_size, len(gpus)))
        g = gpus[self.rank % len(gpus)]
        self.device = torch.device(g if isinstance(g, str) else "cuda:%d" % int(g))
        if self.world_size > 1 and not dist.is_initialized():
            if self.rank == 0:
                logging.info("Initializing distributed process group")
            torch.cuda.set_device(self.device)
            comm.init_process_group("nccl", init_method="env://")
        if self.rank == 0:
            logging.info("Preprocess training set")
        torch.cuda.set_device(self.device)
        self.model = model.cuda(self.device)
        self.train_set, self.valid_set, self.test_set = train_set, valid_set, test_set
        self.optimizer = optimizer
        self.scheduler = scheduler

    def _allreduce_grads(self, local_step=True):
        return allreduce_mean_grads([p for p in self.model.parameters() if p.requires_grad], self.world_size, local_step)

    def train(self, batch_per_epoch, smoothing, print_every):
        if comm.get_rank() == 0:
            logging.info(">>>>> Predictor: Training")
        self.train_set.make_batches()
        # the graph, the rule set and the batches are ~10^6 long-lived Python objects: keep them out of the cyclic
        # collector's generations for the loop (a generation-2 pass over them is a 30-70 ms host stall per occurrence)
        import gc
        freeze = type(self.model).__name__ == "Predictor"        # PredictorPlus steps were bimodal with frozen generations
        if freeze:
            gc.collect()
            gc.freeze()
        try:
            self._train_loop(batch_per_epoch, smoothing, print_every)
        finally:
            if freeze:
                gc.unfreeze()
        if self.scheduler:
            self.scheduler.step()

    def _train_loop(self, batch_per_epoch, smoothing, print_every):
        order = shard_indices(len(self.train_set), self.world_size, self.rank, epoch=0)
        batch_per_epoch = batch_per_epoch or len(order)
        order = order[:batch_per_epoch]
        model = self.model
        model.train()
        total_loss, total_size = 0.0, 0.0
        k = max(1, int(self.slots_per_step))
        use_mask = getattr(model, "entity_feature", "bias") not in ("bias", "RotatE")
        N = self.train_set.graph.entity_size
        done = 0
        steps = [[self.train_set.batch_arrays[i] for i in order[s0:s0 + k]] for s0 in range(0, len(order), k)]
        # Software pipeline: the grounding of step i+1 (parameter-independent) is enqueued BEFORE the host waits
        # for step i's losses and flags, so the GPU has work queued while the host reads step i back, exchanges
        # its gradients and steps the optimizer.  The flags are always checked before the optimizer step: a step
        # whose 32-bit counts overflowed (or whose cell arrays were too small) is redone synchronously first.
        from .predictors import RlStepOverflow
        from .predictors import _used_params
        from . import cellpath
        graphs = bool(self.use_graphs and k == 1 and getattr(model, "supports_pipeline", False))
        pipelined = bool(self.pipelined and getattr(model, "supports_pipeline", False)) and not graphs
        ticket = None
        if pipelined and steps:
            from .data import StepPrefetcher
            packed = StepPrefetcher(model.pack_train_step, steps, depth=2)      # host packing on a loader thread
            ticket = model.prepare_train_step(next(packed)).finish(smoothing, grad_scale=1.0 / len(steps[0]))
        gs_next = None                                                 # captured step whose grounding half is already enqueued
        as_batch = lambda b: np.asarray(b, dtype=np.int64).reshape(-1, 3)
        for si, batches in enumerate(steps):
            self.optimizer.zero_grad(set_to_none=True)
            res = None
            if graphs:
                # Reference schedule on per-head CUDA graphs: score step i, enqueue the (parameter-independent) grounding of
                # step i+1 behind it, and only then wait for step i's loss and flags.
                gs, gs_next = gs_next, None
                if gs is None:
                    gs = cellpath.graph_step_for(model, as_batch(batches[0]), smoothing)
                    if gs is not None:
                        gs.launch_ground(as_batch(batches[0]))
                if gs is not None:
                    gs.launch_score()
                    if si + 1 < len(steps):
                        nb = as_batch(steps[si + 1][0])
                        gs_next = cellpath.graph_step_for(model, nb, smoothing)
                        if gs_next is not None:
                            gs_next.launch_ground(nb)
                    res = cellpath.graph_step_result(model, gs)
            if res is not None:                                        # the whole step was two graph replays + one event wait
                loss, tsum, cand = [res[0]], [res[1]], [res[2]]
                res[3].assign(_used_params(model, res[2]))
            elif graphs:                                               # eager redo (overflow / larger cell arrays / no capture)
                loss, tsum = model.fused_train_step(batches, smoothing, grad_scale=1.0 / len(batches))
                cand = getattr(model, "last_mask_sum", None)
                if gs_next is not None:                                # the redo used the shared frontier workspace
                    gs_next.launch_ground(as_batch(steps[si + 1][0]))
            elif pipelined:
                prep = model.prepare_train_step(next(packed)) if si + 1 < len(steps) else None
                try:
                    loss, tsum = ticket.result()
                    ticket.gbuf.assign(ticket.used)
                    cand = ticket.mask_sum
                except RlStepOverflow:
                    loss, tsum = model.fused_train_step(batches, smoothing, grad_scale=1.0 / len(batches))
                    cand = model.last_ticket.mask_sum
                    if prep is not None:                               # the redo used the shared frontier workspace
                        prep = model.prepare_train_step(prep.batches)
            else:
                loss, tsum = model.fused_train_step(batches, smoothing, grad_scale=1.0 / len(batches))
                cand = getattr(model, "last_mask_sum", None)          # per-batch mask.sum() in mask mode
            local_step = not (use_mask and cand is not None and all(c == 0 for c in cand))
            if self._allreduce_grads(local_step):                     # trainer.py:87: no candidates anywhere -> no step
                self.optimizer.step()
            if pipelined:
                ticket = prep.finish(smoothing, grad_scale=1.0 / len(steps[si + 1])) if prep is not None else None
            self.optimizer.zero_grad(set_to_none=True)
            for j, b in enumerate(batches):
                msum = float(cand[j]) if (use_mask and cand is not None) else float(len(b) * N)
                if msum != 0:
                    total_loss += float(loss[j])
                    total_size += msum
                done += 1
                if done % print_every == 0:
                    if comm.get_rank() == 0:
                        logging.info("{} {} {:.6f} {:.1f}".format(done, len(order), total_loss / print_every,
                                                                  total_size / print_every))
                    total_loss, total_size = 0.0, 0.0

    @torch.no_grad()
    def compute_H(self, print_every):
        if comm.get_rank() == 0:
            logging.info(">>>>> Predictor: Computing H scores of rules")
        order = shard_indices(len(self.train_set), self.world_size, self.rank, epoch=0)
        model = self.model
        model.eval()
        all_H_score = torch.zeros(model.num_rules, device=self.device)
        n_train = len(model.graph.train_facts)
        for batch_id, i in enumerate(order):
            data = self.train_set.batches[i]
            all_h = torch.tensor([d[0] for d in data], device=self.device)
            all_r = torch.tensor([d[1] for d in data], device=self.device)
            all_t = torch.tensor([d[2] for d in data], device=self.device)
            etr = torch.tensor(self.train_set.edges_to_remove(data), device=self.device)
            H, index = model.compute_H(all_h, all_r, all_t, etr)
            if H is not None and index is not None:
                all_H_score[index] += H / n_train
            if (batch_id + 1) % print_every == 0 and comm.get_rank() == 0:
                logging.info("{} {}".format(batch_id + 1, len(order)))
        comm.all_reduce_sum_(all_H_score)                       # == comm.stack(...).sum(0), trainer.py:139-141
        return all_H_score.data.cpu().numpy().tolist()

    @torch.no_grad()
    def evaluate(self, split, expectation=True):
        if comm.get_rank() == 0:
            logging.info(">>>>> Predictor: Evaluating on {}".format(split))
        test_set = getattr(self, "%s_set" % split)
        order = shard_indices(len(test_set), self.world_size, self.rank, epoch=0)
        model = self.model
        model.eval()
        rows = []
        step = max(1, int(self.eval_batches_per_call))
        for s0 in range(0, len(order), step):
            batches = [test_set.batch_arrays[i] for i in order[s0:s0 + step]]
            LH = model.fused_rank(batches, split)
            tri = torch.from_numpy(np.concatenate(batches)).to(self.device)
            rows.append(torch.cat([tri, LH], dim=1))
        ranks = torch.cat(rows, dim=0) if rows else torch.zeros(0, 5, dtype=torch.long, device=self.device)
        ranks = comm.cat_rows(ranks)                              # trainer.py:204-205
        res = summarize_ranks(model, ranks, expectation, test_set.graph.entity_size)
        if comm.get_rank() == 0:
            logging.info("Data : {}".format(res["data"]))
            logging.info("Hit1 : {:.6f}".format(res["hit1"]))
            logging.info("Hit3 : {:.6f}".format(res["hit3"]))
            logging.info("Hit10: {:.6f}".format(res["hit10"]))
            logging.info("MR   : {:.6f}".format(res["mr"]))
            logging.info("MRR  : {:.6f}".format(res["mrr"]))
        self.last_eval = res
        return res["mrr"]

    def load(self, checkpoint, load_optimizer=True):
        if comm.get_rank() == 0:
            logging.info("Load checkpoint from %s" % checkpoint)
        checkpoint = os.path.expanduser(checkpoint)
        state = torch.load(checkpoint, map_location=self.device)
        self.model.load_state_dict(state["model"])
        if load_optimizer:
            self.optimizer.load_state_dict(state["optimizer"])
            for state in self.optimizer.state.values():
                for k, v in state.items():
                    if isinstance(v, torch.Tensor):
                        state[k] = v.to(self.device)
        comm.synchronize()

    def save(self, checkpoint):
        if comm.get_rank() == 0:
            logging.info("Save checkpoint to %s" % checkpoint)
        checkpoint = os.path.expanduser(checkpoint)
        if self.rank == 0:
            torch.save({"model": self.model.state_dict(), "optimizer": self.optimizer.state_dict()}, checkpoint)
        comm.synchronize()


def dedup_weights(rows: np.ndarray) -> np.ndarray:
    """trainer.py:207-209 keeps ONE (L,H) per (h,r,t) (the last row wins); weight 1 marks it."""
    w = np.zeros(rows.shape[0], dtype=np.float64)
    if rows.shape[0]:
        key = rows[::-1, :3]
        _, first = np.unique(key, axis=0, return_index=True)
        w[rows.shape[0] - 1 - first] = 1.0
    return w


def summarize_ranks(model, ranks: torch.Tensor, expectation: bool, num_entities: int):
    """trainer.py:207-238 on the gathered int64[Q,5] rows: de-dup by (h,r,t), tie expectation in
    closed form on the device (rl_rank_metrics), divide by the number of ROWS."""
    host = ranks.cpu().numpy()
    w = dedup_weights(host)
    sk = next(iter(model._drivers.values()))
    sums = sk.rank_metrics(ranks[:, 3:5].contiguous(), torch.from_numpy(w).to(ranks.device), expectation).cpu().numpy()
    n = max(1, host.shape[0])
    return {"data": int(w.sum()), "hit1": float(sums[0] / n), "hit3": float(sums[1] / n), "hit10": float(sums[2] / n),
            "mr": float(sums[3] / n), "mrr": float(sums[4] / n)}
