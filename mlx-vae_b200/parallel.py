"""Data-parallel plumbing for the training step (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the
CPU tests).  The reference is single-device; this is new.

The molecule batch is sharded evenly over ranks, weights are replicated.  Two exchanges per step:
  1. forward: all-reduce(sum) of the loss kernel's batch statistics (2L+5 doubles) so every rank evaluates the
     GLOBAL-batch MI term and scales its gradients by 1/global_B, 1/global_tokens (losses/info.py:33-41 couples
     the batch, F6);
  2. backward: all-reduce(sum) of the flat gradient buffers — the decoder's is issued as soon as the decoder reverse
     pass is enqueued and overlaps the encoder BPTT; the encoder's follows.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Rows [lo, hi) of a global batch of n owned by `rank` (even split; the remainder goes to the low ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradSync:
    """All-reduce helper bound to a process group.  Works on CUDA (nccl) and CPU (gloo) tensors."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.world = dist.get_world_size(group) if self.enabled else 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        self._pending: List = []
        self.bytes_reduced = 0

    def allreduce_stats(self, stats: torch.Tensor):
        """Blocking w.r.t. the stream order (the loss kernel's second phase is enqueued right after)."""
        if self.enabled:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=self.group)
            self.bytes_reduced += stats.numel() * stats.element_size()

    def allreduce_async(self, flat: torch.Tensor):
        if self.enabled:
            self._pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.bytes_reduced += flat.numel() * flat.element_size()

    def wait(self):
        for w in self._pending:
            w.wait()
        self._pending.clear()

    def broadcast(self, t: torch.Tensor, src: int = 0):
        """Rank `src` of the group overwrites `t` everywhere (initial parameters / optimizer moments of the replicas)."""
        if self.enabled:
            dist.broadcast(t, src=dist.get_global_rank(self.group, src) if self.group is not None else src,
                           group=self.group)
