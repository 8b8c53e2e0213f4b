"""Multi-GPU plumbing: one process per GPU (torchrun), particles replicated, primary tiles sharded,
per-bin histograms summed with one all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).

The path has no data-path exchange step: cell pairs are independent work units whose only shared
output is a histogram of a few hundred 8-byte slots, so the collective is a single latency-bound
all-reduce per call."""
from __future__ import annotations

import numpy as np


def make_allreduce(dist, device=None):
    """Returns reduce_fn(npairs, sum_sep, sum_w) summing the histograms over the default group."""
    import torch

    def reduce_fn(npairs, sum_sep, sum_w):
        # uint64 counts travel as int64 (exact: counts are far below 2^63)
        t_n = torch.from_numpy(npairs.view(np.int64).copy())
        t_f = torch.from_numpy(np.concatenate([sum_sep, sum_w]))
        if device is not None:
            t_n = t_n.to(device)
            t_f = t_f.to(device)
        dist.all_reduce(t_n, op=dist.ReduceOp.SUM)
        dist.all_reduce(t_f, op=dist.ReduceOp.SUM)
        n = npairs.size
        npairs[:] = t_n.cpu().numpy().view(np.uint64)
        f = t_f.cpu().numpy()
        sum_sep[:] = f[:n]
        sum_w[:] = f[n:]

    return reduce_fn


def enable_distributed(dist=None, device=None):
    """Shard this process's pair counting by torch.distributed rank.  Call after init_process_group."""
    from . import _lib

    if dist is None:
        import torch.distributed as dist  # noqa: PLC0415
    rank, world = dist.get_rank(), dist.get_world_size()
    _lib.set_shard(rank, world, make_allreduce(dist, device) if world > 1 else None)
    return rank, world


def disable_distributed():
    from . import _lib

    _lib.set_shard(0, 1, None)
