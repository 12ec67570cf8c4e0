"""Batch builders with the reference's interface (src/data.py:175-293).

Batches are single-relation groups of <= batch_size triples, shuffled with Python's ``random``
in the same call order as the reference so that a seeded run visits the same batches.
``__getitem__`` still returns the reference's dense tensors (API compatibility); the B200
trainer does not use them -- it reads ``batches[idx]`` and lets the kernels look targets and
filters up in the device-resident answer lists (rl_answers)."""
import random

import numpy as np
import torch
from torch.utils.data import Dataset

from .graph import KnowledgeGraph  # noqa: F401  (``from data import KnowledgeGraph`` in the run scripts)


class TrainDataset(Dataset):
    def __init__(self, graph, batch_size):
        self.graph = graph
        self.batch_size = batch_size
        self.r2instances = [[] for r in range(self.graph.relation_size)]
        for h, r, t in self.graph.train_facts:
            self.r2instances[r].append((h, r, t))
        self.make_batches()

    def make_batches(self):
        for r in range(self.graph.relation_size):
            random.shuffle(self.r2instances[r])
        self.batches = list()
        for r, instances in enumerate(self.r2instances):
            for k in range(0, len(instances), self.batch_size):
                self.batches.append(instances[k:min(k + self.batch_size, len(instances))])
        random.shuffle(self.batches)
        self.batch_arrays = [np.array(b, dtype=np.int64).reshape(-1, 3) for b in self.batches]

    def __len__(self):
        return len(self.batches)

    def edges_to_remove(self, data):
        return [self.graph.relation2ht2index[r][self.graph.encode_ht(h, t)] for h, r, t in data]

    def __getitem__(self, idx):
        data = self.batches[idx]
        all_h = torch.LongTensor([_[0] for _ in data])
        all_r = torch.LongTensor([_[1] for _ in data])
        all_t = torch.LongTensor([_[2] for _ in data])
        target = torch.zeros(len(data), self.graph.entity_size)
        for k, (h, r, t) in enumerate(data):
            target[k][torch.LongTensor(self.graph.hr2o[self.graph.encode_hr(h, r)])] = 1
        return all_h, all_r, all_t, target, torch.LongTensor(self.edges_to_remove(data))


class _EvalDataset(Dataset):
    split = "valid"
    answers = "hr2oo"

    def __init__(self, graph, batch_size):
        self.graph = graph
        self.batch_size = batch_size
        facts = getattr(self.graph, self.split + "_facts")
        r2instances = [[] for r in range(self.graph.relation_size)]
        for h, r, t in facts:
            r2instances[r].append((h, r, t))
        self.batches = list()
        for r, instances in enumerate(r2instances):
            random.shuffle(instances)
            for k in range(0, len(instances), self.batch_size):
                self.batches.append(instances[k:min(k + self.batch_size, len(instances))])
        self.batch_arrays = [np.array(b, dtype=np.int64).reshape(-1, 3) for b in self.batches]

    def __len__(self):
        return len(self.batches)

    def __getitem__(self, idx):
        data = self.batches[idx]
        all_h = torch.LongTensor([_[0] for _ in data])
        all_r = torch.LongTensor([_[1] for _ in data])
        all_t = torch.LongTensor([_[2] for _ in data])
        known = getattr(self.graph, self.answers)
        mask = torch.ones(len(data), self.graph.entity_size).bool()
        for k, (h, r, t) in enumerate(data):
            mask[k][torch.LongTensor(known[self.graph.encode_hr(h, r)])] = 0
        return all_h, all_r, all_t, mask


class ValidDataset(_EvalDataset):
    split = "valid"
    answers = "hr2oo"


class TestDataset(_EvalDataset):
    split = "test"
    answers = "hr2ooo"


class StepPrefetcher(object):
    """Runs ``pack_fn`` (e.g. ``model.pack_train_step``) over an iterable of steps on a loader thread, ``depth``
    steps ahead of the consumer -- the host-side packing of step k+2 then overlaps the enqueue of step k+1 and
    the wait for step k.  Iterating yields the packed steps in order; an exception in the thread is re-raised."""

    _END = object()

    def __init__(self, pack_fn, steps, depth=2):
        import queue
        import threading
        self._q = queue.Queue(maxsize=max(1, int(depth)))
        self._stop = threading.Event()

        def work():
            try:
                for st in steps:
                    item = pack_fn(st)
                    while not self._stop.is_set():
                        try:
                            self._q.put(item, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if self._stop.is_set():
                        return
                self._q.put(self._END)
            except BaseException as e:                      # surfaces in the consumer
                self._q.put(e)

        self._thread = threading.Thread(target=work, daemon=True)
        self._thread.start()

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is self._END:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    def close(self):
        self._stop.set()
