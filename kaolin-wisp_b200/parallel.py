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


class HostPrefetcher:
    """Pinned host batches -> device, one batch ahead, on a side stream (the trainer-side analogue of the reference's
    DataLoader(pin_memory=True) + `.to(device)` in the step, multiview_trainer.py:73-82): the host->device copy of batch
    i+1 overlaps the render step of batch i.  Iterating yields tuples of device tensors that are safe to use on the
    current stream."""

    def __init__(self, batches: Iterable, device, stream=None):
        self.batches, self.device = batches, torch.device(device)
        self.stream = stream if stream is not None else (torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None)
        self.staged, self.staged_event = None, None       # the batch after the one being consumed (device tensors, copy event)

    def _stage(self, batch):
        if self.stream is None:
            return tuple(t.to(self.device) for t in batch), None
        with torch.cuda.stream(self.stream):
            out = tuple(t.to(self.device, non_blocking=True) for t in batch)
            ev = torch.cuda.Event(); ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._stage(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._stage(next(it))         # enqueue the next copy before handing out the current batch
            except StopIteration:
                nxt = None
            self.staged, self.staged_event = nxt if nxt is not None else (None, None)
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
                for t in cur:
                    t.record_stream(torch.cuda.current_stream(self.device))
            yield cur
