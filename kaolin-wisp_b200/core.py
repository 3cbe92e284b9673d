"""Data carriers across the tracer boundary: host-side mirrors of wisp.core.Rays (wisp/core/rays.py:19-198) and
wisp.core.RenderBuffer (wisp/core/render_buffer.py:21-200), reduced to what the render path reads and writes.
When the real wisp package is importable, wisp_b200.install() keeps wisp's own classes instead of these."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Union

import torch

INFINITY = torch.finfo().max


@dataclass
class Rays:
    origins: torch.Tensor
    dirs: torch.Tensor
    dist_min: Union[float, torch.Tensor] = 0.0
    dist_max: Union[float, torch.Tensor] = INFINITY

    def __len__(self) -> int:
        if self.origins.shape != self.dirs.shape:
            raise Exception(f"Rays.origins shape should match Rays.dirs shape, but got {self.origins.shape} and {self.dirs.shape}.")
        return self.origins.shape[0]

    @property
    def shape(self):
        return self.origins.shape[:-1]

    def to(self, *args, **kwargs) -> "Rays":
        mv = lambda v: v.to(*args, **kwargs) if torch.is_tensor(v) else v
        return Rays(mv(self.origins), mv(self.dirs), mv(self.dist_min), mv(self.dist_max))

    def split(self, split_size) -> list:
        return [Rays(o, d, self.dist_min, self.dist_max) for o, d in zip(self.origins.split(split_size), self.dirs.split(split_size))]

    def reshape(self, *dims) -> "Rays":
        return Rays(self.origins.reshape(*dims), self.dirs.reshape(*dims), self.dist_min, self.dist_max)


class RenderBuffer:
    """Per-ray output channels.  `+` concatenates along dim 0 like the reference (render_buffer.py:167-200)."""

    def __init__(self, rgb=None, alpha=None, depth=None, hit=None, **extra):
        self._ch: Dict[str, Optional[torch.Tensor]] = dict(rgb=rgb, alpha=alpha, depth=depth, hit=hit, **extra)

    def __getattr__(self, name):
        ch = self.__dict__.get("_ch", {})
        if name in ch:
            return ch[name]
        raise AttributeError(name)

    @property
    def channels(self):
        return {k for k, v in self._ch.items() if v is not None}

    def __add__(self, other: "RenderBuffer") -> "RenderBuffer":
        out = {}
        for k in set(self._ch) | set(other._ch):
            a, b = self._ch.get(k), other._ch.get(k)
            out[k] = a if b is None else b if a is None else torch.cat([a, b], 0)
        return RenderBuffer(**out)

    def reshape(self, *dims) -> "RenderBuffer":
        return RenderBuffer(**{k: (v.reshape(*dims, *v.shape[1:]) if v is not None else None) for k, v in self._ch.items()})

    def cpu(self) -> "RenderBuffer":
        return RenderBuffer(**{k: (v.cpu() if v is not None else None) for k, v in self._ch.items()})
