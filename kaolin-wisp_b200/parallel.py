"""Multi-GPU plumbing: the render path shards by rays with no data-path collective; the only exchange is one
gradient all-reduce per optimiser step.  One process per GPU, torch.distributed (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_range(num_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of `num_items` rays / image rows for `rank` (image tiles per GPU)."""
    base, rem = divmod(num_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(origins: torch.Tensor, dirs: torch.Tensor, rank: int, world: int):
    b, e = shard_range(origins.shape[0], rank, world)
    return origins[b:e], dirs[b:e]


class GradientReducer:
    """All-reduce(mean) of the gradients of `params`: the large hash table is reduced in place, the small decoder
    tensors are coalesced into one flat bucket so that a step issues exactly two collectives."""

    def __init__(self, params: Iterable[torch.nn.Parameter], big_threshold: int = 1 << 20):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.big = [p for p in self.params if p.numel() >= big_threshold]
        self.small = [p for p in self.params if p.numel() < big_threshold]

    def reduce(self, group=None) -> None:
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        for p in self.big:
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
                p.grad.div_(world)
        grads = [p.grad for p in self.small if p.grad is not None]
        if grads:
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)
            o = 0
            for g in grads:
                g.copy_(flat[o:o + g.numel()].view_as(g)); o += g.numel()


def sync_sample_counts(num_samples: int, num_rays: int, device, group=None) -> Tuple[int, int]:
    """Global (sum over ranks) hit-sample and ray counts, so that every rank derives the same adaptive ray budget
    (multiview_trainer.py:95-109 uses tracer.prev_num_samples for that)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return num_samples, num_rays
    t = torch.tensor([num_samples, num_rays], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0]), int(t[1])
