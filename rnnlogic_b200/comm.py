"""Process-group helpers with the names the reference scripts import (src/comm.py:13-80).  Only
what the predictor path uses: rank / world size, group setup, barrier.  The data path itself has
one collective -- the flat gradient all-reduce in trainer.py -- plus small metric reductions."""
import multiprocessing
import os

import torch
from torch import distributed as dist

cpu_group = None
gpu_group = None


def get_rank():
    if dist.is_initialized():
        return dist.get_rank()
    if "RANK" in os.environ:
        return int(os.environ["RANK"])
    return 0


def get_world_size():
    if dist.is_initialized():
        return dist.get_world_size()
    if "WORLD_SIZE" in os.environ:
        return int(os.environ["WORLD_SIZE"])
    return 1


def get_group(device):
    group = cpu_group if device.type == "cpu" else gpu_group
    if group is None:
        raise ValueError("%s group is not initialized. Use comm.init_process_group() to initialize it"
                         % device.type.upper())
    return group


def init_process_group(backend, init_method=None, **kwargs):
    global cpu_group, gpu_group
    dist.init_process_group(backend, init_method, **kwargs)
    gpu_group = dist.group.WORLD
    cpu_group = dist.new_group(backend="gloo") if backend == "nccl" else gpu_group


def get_cpu_count():
    return multiprocessing.cpu_count()


def synchronize():
    if get_world_size() > 1:
        dist.barrier()


def all_reduce_sum_(tensor):
    """In-place SUM all-reduce of one flat tensor on its device's group (no-op at world 1)."""
    if get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=get_group(tensor.device))
    return tensor


def cat_rows(tensor):
    """Concatenate a [n_i, k] tensor over ranks (replaces comm.cat, src/comm.py:228-256)."""
    world = get_world_size()
    if world == 1:
        return tensor
    group = get_group(tensor.device)
    n = torch.tensor([tensor.shape[0]], dtype=torch.long, device=tensor.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(s.item() for s in sizes))
    pad = torch.zeros((m,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    pad[:tensor.shape[0]] = tensor
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)], dim=0)
