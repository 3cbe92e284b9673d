"""Host-side SPC (structured point cloud / octree) setup utilities.

Octree *construction* is a one-off setup step that the reference delegates to Kaolin
(spc_ops.unbatched_points_to_octree / scan_octrees / generate_points; call sites
wisp/accelstructs/octree_as.py:122-144, wisp/ops/spc/conversions.py:72-88, wisp/ops/spc/constructors.py:14-28).
It is outside the hot path, so it is written with plain torch ops and works on any
device.  The format: one byte per non-leaf node, breadth first, bit c = 4x+2y+z set iff child
c exists; prefix = exclusive sum of popcounts; pyramid[0,l] = #points of level l, pyramid[1,l] = their offset.
"""
from __future__ import annotations

from typing import Tuple

import torch

_POPC = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32)


def morton3(points: torch.Tensor, level: int) -> torch.Tensor:
    p = points.long()
    m = torch.zeros(p.shape[0], dtype=torch.int64, device=p.device)
    for i in range(level):
        m |= ((p[:, 0] >> i) & 1) << (3 * i + 2)
        m |= ((p[:, 1] >> i) & 1) << (3 * i + 1)
        m |= ((p[:, 2] >> i) & 1) << (3 * i)
    return m


def points_to_octree(points: torch.Tensor, level: int) -> torch.Tensor:
    """Quantised points [N,3] in [0, 2^level) -> octree bytes (uint8).  unbatched_points_to_octree(sorted=False)."""
    cur = torch.unique(morton3(points, level))          # sorted
    levels = []
    for _ in range(level):
        parent = cur >> 3
        child = (cur & 7)
        up, inv = torch.unique_consecutive(parent, return_inverse=True)
        byte = torch.zeros(up.shape[0], dtype=torch.int32, device=cur.device)
        byte.index_put_((inv,), (1 << child).to(torch.int32), accumulate=True)   # distinct children -> sum == or
        levels.append(byte.to(torch.uint8))
        cur = up
    if not levels:
        return torch.zeros(0, dtype=torch.uint8, device=points.device)
    return torch.cat(levels[::-1])


def create_dense_octree(level: int, device="cpu") -> torch.Tensor:
    """wisp.ops.spc.create_dense_octree (constructors.py:14-28): all 8^l cells of every level occupied."""
    n = sum(8 ** l for l in range(level))
    return torch.full((n,), 255, dtype=torch.uint8, device=device)


def scan_octree(octree: torch.Tensor) -> Tuple[int, torch.Tensor, torch.Tensor]:
    """-> (max_level, pyramid int32 [2, max_level+2] on CPU, prefix int32 [n+1] on octree.device).  scan_octrees."""
    n = octree.shape[0]
    popc = _POPC.to(octree.device)[octree.long()]
    prefix = torch.zeros(n + 1, dtype=torch.int32, device=octree.device)
    if n:
        prefix[1:] = torch.cumsum(popc, 0)
    pc = prefix.cpu()
    cnts, offs = [], []
    start, count, level = 0, 1, 0
    while True:
        cnts.append(count); offs.append(start)
        if start >= n:
            break
        end = start + count
        if end > n:
            raise ValueError("malformed octree")
        children = int(pc[end] - pc[start])
        start, count, level = end, children, level + 1
    ml = level
    pyramid = torch.zeros(2, ml + 2, dtype=torch.int32)
    pyramid[0, : ml + 1] = torch.tensor(cnts, dtype=torch.int32)
    pyramid[1, : ml + 1] = torch.tensor(offs, dtype=torch.int32)
    pyramid[1, ml + 1] = offs[ml] + cnts[ml]
    return ml, pyramid, prefix


def generate_points(octree: torch.Tensor, pyramid: torch.Tensor) -> torch.Tensor:
    """-> point hierarchy int16 [total,3]: level l+1 = for each level-l point, for each set bit c ascending,
    2*p + (c>>2&1, c>>1&1, c&1).  generate_points."""
    dev = octree.device
    ml = pyramid.shape[1] - 2
    pts = [torch.zeros(1, 3, dtype=torch.int16, device=dev)]
    ar = torch.arange(8, device=dev)
    offs = torch.stack([(ar >> 2) & 1, (ar >> 1) & 1, ar & 1], -1).to(torch.int16)
    for l in range(ml):
        b = octree[int(pyramid[1, l]): int(pyramid[1, l]) + int(pyramid[0, l])].long()
        mask = ((b[:, None] >> ar[None]) & 1).bool()
        idx = torch.nonzero(mask)                           # row-major: (node, c) ascending
        pts.append(2 * pts[l][idx[:, 0]] + offs[idx[:, 1]])
    return torch.cat(pts)


def octree_to_spc(octree: torch.Tensor):
    """wisp.ops.spc.octree_to_spc (conversions.py:72-88) -> points, pyramid, prefix."""
    ml, pyramid, prefix = scan_octree(octree)
    return generate_points(octree, pyramid), pyramid, prefix


def quantize_points(x: torch.Tensor, level: int) -> torch.Tensor:
    """spc_ops.quantize_points: floor(clamp(2^level (x+1)/2, 0, 2^level-1)) -> int16."""
    res = 2 ** level
    return torch.floor(torch.clamp(res * (x + 1.0) / 2.0, 0, res - 1.0)).short()


def make_trilinear_spc(points: torch.Tensor, pyramid: torch.Tensor):
    """wisp.ops.spc.make_trilinear_spc (constructors.py:31-47 -> kaolin unbatched_make_dual / unbatched_make_trinkets).
    -> points_dual int16 [Td,3] (per level: Morton-sorted unique cell corners), pyramid_dual int32 [2, L+2],
       trinkets int32 [T,8] (LEVEL-LOCAL dual index of corner j = 4x+2y+z), parents int32 [T] (-1 for the root)."""
    dev = points.device
    L = pyramid.shape[-1] - 2
    ar = torch.arange(8, device=dev)
    corners = torch.stack([(ar >> 2) & 1, (ar >> 1) & 1, ar & 1], -1).long()
    pts = points.long()
    duals = []
    trinkets = torch.zeros(pts.shape[0], 8, dtype=torch.int32, device=dev)
    parents = torch.full((pts.shape[0],), -1, dtype=torch.int32, device=dev)
    pyr = torch.zeros(2, L + 2, dtype=torch.int32)
    off = 0
    for l in range(L + 1):
        s, c = int(pyramid[1, l]), int(pyramid[0, l])
        cell = pts[s:s + c]
        cor = (cell[:, None, :] + corners[None]).reshape(-1, 3)
        key = morton3(cor, l + 1)
        uk, inv = torch.unique(key, return_inverse=True)           # sorted
        first = torch.full((uk.shape[0],), cor.shape[0], dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inv, torch.arange(cor.shape[0], device=dev), reduce="amin")
        duals.append(cor[first].to(torch.int16))
        trinkets[s:s + c] = inv.reshape(c, 8).to(torch.int32)
        pyr[0, l] = uk.shape[0]; pyr[1, l] = off; off += uk.shape[0]
        if l > 0:
            ps, pc = int(pyramid[1, l - 1]), int(pyramid[0, l - 1])
            pkey = morton3(pts[ps:ps + pc], l)
            parents[s:s + c] = (ps + torch.searchsorted(pkey, morton3(cell >> 1, l))).to(torch.int32)
    pyr[1, L + 1] = off
    return torch.cat(duals), pyr, trinkets, parents
