"""Oracle definitions for the OctreeGrid / NGLOD-SDF path (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

[KAOLIN-EXT, PARITY UNPINNED] kaolin.ops.spc.{unbatched_make_dual, unbatched_make_trinkets, unbatched_interpolate_trilinear,
coords_to_trilinear_coeffs} have no source under /root/reference; these numpy restatements follow SURVEY.md Appendix A and the
call sites wisp/ops/spc/constructors.py:31-47 and wisp/models/grids/octree_grid.py:130-163.  Also restated here:
find_depth_bound (wisp/csrc/render/find_depth_bound_cuda.cu:16-45).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def _morton_wide(p: np.ndarray, bits: int) -> np.ndarray:
    p = p.astype(np.int64)
    m = np.zeros(p.shape[0], dtype=np.int64)
    for i in range(bits):
        m |= ((p[:, 0] >> i) & 1) << (3 * i + 2)
        m |= ((p[:, 1] >> i) & 1) << (3 * i + 1)
        m |= ((p[:, 2] >> i) & 1) << (3 * i)
    return m


CORNERS = np.array([[(j >> 2) & 1, (j >> 1) & 1, j & 1] for j in range(8)], dtype=np.int64)


def make_trilinear_spc(spc: O.SPC):
    """-> points_dual int16 [Td,3], pyramid_dual int64 [2, L+2], trinkets int32 [T,8] (LEVEL-LOCAL dual index of corner j,
    z fastest), parents int32 [T] (global index of the parent cell, -1 for the root)."""
    L = spc.max_level
    pts = spc.points.astype(np.int64)
    duals, trinkets, parents = [], np.zeros((pts.shape[0], 8), np.int32), np.full(pts.shape[0], -1, np.int32)
    pyr = np.zeros((2, L + 2), np.int64)
    off = 0
    for l in range(L + 1):
        s, c = int(spc.pyramid[1, l]), int(spc.pyramid[0, l])
        cell = pts[s:s + c]
        corners = (cell[:, None, :] + CORNERS[None]).reshape(-1, 3)
        key = _morton_wide(corners, l + 1)
        uk, inv = np.unique(key, return_inverse=True)
        first = np.zeros(uk.shape[0], np.int64); first[inv[::-1]] = np.arange(corners.shape[0])[::-1]
        duals.append(corners[first].astype(np.int16))
        trinkets[s:s + c] = inv.reshape(c, 8).astype(np.int32)
        pyr[0, l] = uk.shape[0]; pyr[1, l] = off; off += uk.shape[0]
        if l > 0:
            ps, pc = int(spc.pyramid[1, l - 1]), int(spc.pyramid[0, l - 1])
            pkey = _morton_wide(pts[ps:ps + pc], l)
            ckey = _morton_wide(cell >> 1, l)
            parents[s:s + c] = (ps + np.searchsorted(pkey, ckey)).astype(np.int32)
    pyr[1, L + 1] = off
    return np.concatenate(duals), pyr, trinkets, parents


def trilinear_coeffs(coords: np.ndarray, points: np.ndarray, level: int) -> np.ndarray:
    """coords_to_trilinear_coeffs: u = 2^level (c*0.5+0.5) - point; [(1-ux)(1-uy)(1-uz), (1-ux)(1-uy)uz, ...] z fastest."""
    u = (2.0 ** level) * (coords.astype(np.float64) * 0.5 + 0.5) - points.astype(np.float64)
    u = u.astype(np.float32); iu = (1.0 - u).astype(np.float32)
    out = np.zeros((coords.shape[0], 8), np.float32)
    for j in range(8):
        cx = u[:, 0] if (j & 4) else iu[:, 0]; cy = u[:, 1] if (j & 2) else iu[:, 1]; cz = u[:, 2] if (j & 1) else iu[:, 2]
        out[:, j] = (cx * cy) * cz
    return out


def interpolate_trilinear(coords: np.ndarray, pidx: np.ndarray, points: np.ndarray, trinkets: np.ndarray, feats: np.ndarray, level: int,
                          half: bool = True) -> np.ndarray:
    """unbatched_interpolate_trilinear for one sample per cell: out = sum_j feats[trinkets[pidx, j]] * coef_j; pidx == -1 -> 0.
    half=True reproduces the call site's `feats.half()` ... `.float()` (octree_grid.py:147-149): features and result rounded to fp16."""
    N, Fd = coords.shape[0], feats.shape[1]
    out = np.zeros((N, Fd), np.float32)
    ok = pidx >= 0
    if ok.any():
        p = pidx[ok].astype(np.int64)
        cf = trilinear_coeffs(coords[ok], points[p], level)
        f = feats.astype(np.float16).astype(np.float32) if half else feats.astype(np.float32)
        acc = np.zeros((p.shape[0], Fd), np.float32)
        for j in range(8):
            acc = acc + f[trinkets[p, j].astype(np.int64)] * cf[:, j:j + 1]
        out[ok] = acc.astype(np.float16).astype(np.float32) if half else acc
    return out


def octree_grid_interpolate(spc: O.SPC, trinkets: np.ndarray, features, active_lods, coords: np.ndarray, lod_idx: int, multiscale: str,
                            half: bool = True) -> np.ndarray:
    """OctreeGrid.interpolate (octree_grid.py:165-219)."""
    base = active_lods[0]
    pidx = O.query(spc, coords, active_lods[lod_idx], with_parents=True)[:, base:]
    feats = [interpolate_trilinear(coords, pidx[:, i], spc.points, trinkets, features[i], active_lods[i], half) for i in range(lod_idx + 1)]
    if lod_idx == 0:
        return feats[0]
    out = np.concatenate(feats, -1)
    if multiscale == "sum":
        out = out.reshape(out.shape[0], lod_idx + 1, -1).sum(-2)
    return out


def find_depth_bound(query: np.ndarray, curr_idxes: np.ndarray, depth: np.ndarray) -> np.ndarray:
    """wisp/csrc/render/find_depth_bound_cuda.cu:16-45 including its quirks: the output starts at -1 (find_depth_bound.cpp),
    the scan of pack i stops at the CURRENT cursor of pack i+1, and the last pack is bounded by num_packs (not num_nugs)."""
    P = query.shape[0]
    out = np.full(P, -1, np.int32)
    q = query.reshape(-1)
    for t in range(P):
        if curr_idxes[t] <= -1:
            continue
        i = int(curr_idxes[t]); mx = P if t == P - 1 else int(curr_idxes[t + 1])
        mx = mx & 0xFFFFFFFF                                     # `uint max_iidx = ...`: a -1 cursor of the next pack wraps around
        while i < mx:
            if i >= depth.shape[0]:
                break
            en, ex = depth[i, 0], depth[i, 1]
            if (q[t] >= en and q[t] <= ex) or q[t] < en:
                out[t] = i
                break
            i += 1
    return out
