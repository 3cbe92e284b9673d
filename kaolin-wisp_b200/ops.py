"""Tensor-level operators over the C ABI: the functions the wisp-facing classes (and wisp itself, once patched by
wisp_b200.install()) call.  Each one names the reference operator it stands in for.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _cabi as A

# Optional per-stage device timing (bench.py): when PROFILE is a list, every native stage appends
# (name, start_event, end_event) recorded on the launching (current torch) stream.
PROFILE: Optional[list] = None


class _stage:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1))
        return False


# --------------------------------------------------------------------------------------------------------------
# octree handle: the SPC tensors the reference's OctreeAS keeps (octree_as.py:58-62) + the optional dense bitmask
# --------------------------------------------------------------------------------------------------------------
COARSE_LEVEL = 5      # level of the dilated mask the marcher uses to skip empty 32-candidate words (0 disables it)


@dataclass
class OctreeTensors:
    octree: torch.Tensor          # uint8 [nbytes]
    prefix: torch.Tensor          # int32 [nbytes+1]
    points: torch.Tensor          # int16 [total,3]
    pyramid: torch.Tensor         # int32 [2, max_level+2] (CPU)
    max_level: int
    bits: Optional[torch.Tensor] = None
    bits_level: int = -1
    bbox: Optional[tuple] = None          # (lo[3], hi[3]) of the occupied cells of `bits_level`, normalised coords
    coarse: Optional[torch.Tensor] = None # dilated occupancy of `coarse_level` (word-skipping in the marcher), or None
    coarse_level: int = 0

    def desc(self) -> A.OctreeDesc:
        d = A.OctreeDesc()
        d.octree, d.prefix, d.nbytes, d.max_level = self.octree.data_ptr(), self.prefix.data_ptr(), self.octree.shape[0], self.max_level
        d.bits = self.bits.data_ptr() if self.bits is not None else None
        d.bits_level = self.bits_level
        d.has_bbox = 0
        if self.bbox is not None:
            d.has_bbox = 1
            for a in range(3):
                d.bbox_lo[a], d.bbox_hi[a] = self.bbox[0][a], self.bbox[1][a]
        d.coarse_bits = self.coarse.data_ptr() if self.coarse is not None else None
        d.coarse_level = self.coarse_level if self.coarse is not None else 0
        return d

    def ensure_bits(self, level: int) -> None:
        """Dense occupancy bitmask of `level` (<= 10): 8^level bits, built once per octree by wb_octree_build_bits."""
        if level > 10 or (self.bits is not None and self.bits_level == level):
            return
        A.require_device(self.octree)
        words = (8 ** level + 31) // 32
        bits = torch.zeros(words, dtype=torch.int32, device=self.octree.device)
        start, cnt = int(self.pyramid[1, level]), int(self.pyramid[0, level])
        lvl = self.points[start:start + cnt].contiguous()
        A.check(A.lib().wb_octree_build_bits(A.ptr(lvl), C.c_int64(cnt), C.c_int32(level), A.ptr(bits), A.stream()))
        self.bits, self.bits_level = bits, level
        self.coarse, self.coarse_level = None, 0
        cl = min(level - 2, COARSE_LEVEL)
        if cl >= 2 and cnt > 0:
            coarse = torch.zeros((8 ** cl + 31) // 32, dtype=torch.int32, device=self.octree.device)
            A.check(A.lib().wb_octree_build_coarse(A.ptr(lvl), C.c_int64(cnt), C.c_int32(level), C.c_int32(cl), A.ptr(coarse), A.stream()))
            self.coarse, self.coarse_level = coarse, cl
        if cnt > 0:     # one-off (per octree) host read of the occupied extent; exact dyadic cell faces
            mn, mx = lvl.min(0).values.cpu().tolist(), lvl.max(0).values.cpu().tolist()
            res = float(2 ** level)
            self.bbox = ([2.0 * m / res - 1.0 for m in mn], [2.0 * (m + 1) / res - 1.0 for m in mx])


def query(oct: OctreeTensors, coords: torch.Tensor, level: int, with_parents: bool = False) -> torch.Tensor:
    """spc_ops.unbatched_query(octree, prefix, coords, level, with_parents)  (octree_as.py:162)."""
    A.require_device(coords)
    c = A.f32c(coords)
    N = c.shape[0]
    out = torch.empty((N, level + 1) if with_parents else (N,), dtype=torch.int32, device=c.device)
    d = oct.desc()
    A.check(A.lib().wb_query(C.byref(d), A.ptr(c), C.c_int64(N), C.c_int32(level), C.c_int32(int(with_parents)), A.ptr(out), A.stream()))
    return out


# --------------------------------------------------------------------------------------------------------------
# raymarch 'ray'
# --------------------------------------------------------------------------------------------------------------
@dataclass
class MarchState:
    """Intermediate state of one raymarch: per-ray hit bitmask, counts and offsets (device), total (host)."""
    rays: A.RaysDesc
    keep: list
    n: int
    jitter: Optional[torch.Tensor]
    seed: int
    hitmask: Optional[torch.Tensor]
    counts: Optional[torch.Tensor]
    offsets: torch.Tensor            # int64 [R+1]: sample range of every ray
    total: int
    records: Optional[tuple] = None  # (t, delta, ray) already produced by the 'voxel' / 'uniform' samplers


def march_count(oct: OctreeTensors, origins, dirs, dist_min, dist_max, num_samples: int, level: int,
                jitter: Optional[torch.Tensor] = None, seed: int = 0, defer_total: bool = False):
    """Sample culling of OctreeAS._raymarch_ray (octree_as.py:272-288) without materialising candidates."""
    A.require_device(origins)
    oct.ensure_bits(level)
    rays, keep = A.make_rays(origins, dirs, dist_min, dist_max)
    R = rays.num_rays
    dev = origins.device
    nw = (num_samples + 31) // 32
    hitmask = torch.empty((R, nw), dtype=torch.int32, device=dev)
    counts = torch.empty(R, dtype=torch.int32, device=dev)
    offsets = torch.empty(R + 1, dtype=torch.int64, device=dev)
    jit = None if jitter is None else A.f32c(jitter)
    if jit is not None and tuple(jit.shape) != (R, num_samples):
        raise A.WispB200Error(f"jitter must be [{R}, {num_samples}]")
    od = oct.desc()
    L = A.lib()
    with _stage("march_count"):
        A.check(L.wb_raymarch_ray_count(C.byref(od), C.c_int32(level), C.byref(rays), C.c_int32(num_samples), A.ptr(jit),
                                        C.c_uint32(seed & 0xFFFFFFFF), A.ptr(hitmask), A.ptr(counts), A.stream()))
    wsb = int(L.wb_scan_workspace_bytes(C.c_int64(R)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    with _stage("scan"):
        A.check(L.wb_scan_counts(A.ptr(counts), C.c_int64(R), A.ptr(offsets), A.ptr(ws), C.c_int64(wsb), A.stream()))
    ms = MarchState(rays, keep + [jit], num_samples, jit, seed & 0xFFFFFFFF, hitmask, counts, offsets, -1)
    if defer_total:                        # pre-march on a side stream: the total travels to pinned memory, no host sync here
        host = _pinned_slot()
        host.copy_(offsets[-1:], non_blocking=True)
        ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream(dev))
        return PendingMarch(ms, host, ev, torch.cuda.current_stream(dev))
    ms.total = int(offsets[-1].item())     # the one host sync of the path (the reference syncs in torch.nonzero, octree_as.py:288)
    return ms


_PINNED_RING: list = []
_PINNED_NEXT = 0


def _pinned_slot() -> torch.Tensor:
    """One of 8 page-locked int64 slots allocated once per process (cudaHostAlloc in the training loop costs milliseconds and
    synchronises); a slot is reused only 8 pre-marches later, long after its copy has been consumed."""
    global _PINNED_NEXT
    if not _PINNED_RING:
        buf = torch.empty(8, dtype=torch.int64, pin_memory=True)
        _PINNED_RING.extend(buf[i:i + 1] for i in range(8))
    _PINNED_NEXT = (_PINNED_NEXT + 1) % 8
    return _PINNED_RING[_PINNED_NEXT]


@dataclass
class PendingMarch:
    """A raymarch whose kernels were enqueued on a side stream (PackedRFTracer.premarch): finalize() makes the consumer stream
    wait for them and reads the sample total (normally long since on the host by the time the batch is rendered)."""
    ms: MarchState
    host_total: torch.Tensor
    event: "torch.cuda.Event"
    stream: "torch.cuda.Stream"

    def finalize(self) -> MarchState:
        cur = torch.cuda.current_stream(self.ms.offsets.device)
        cur.wait_event(self.event)
        for t in (self.ms.hitmask, self.ms.counts, self.ms.offsets):
            t.record_stream(cur)
        self.event.synchronize()
        self.ms.total = int(self.host_total[0])
        return self.ms


def march_fill_reference_layout(ms: MarchState, device):
    """ASRaymarchResults tensors (base_as.py:57-84)."""
    S = ms.total
    ridx = torch.empty(S, dtype=torch.int64, device=device)
    samples = torch.empty((S, 3), dtype=torch.float32, device=device)
    depth = torch.empty((S, 1), dtype=torch.float32, device=device)
    deltas = torch.empty((S, 1), dtype=torch.float32, device=device)
    boundary = torch.empty(S, dtype=torch.bool, device=device)
    if S > 0:
        A.check(A.lib().wb_raymarch_ray_fill(C.byref(ms.rays), C.c_int32(ms.n), A.ptr(ms.jitter), C.c_uint32(ms.seed), A.ptr(ms.hitmask),
                                             A.ptr(ms.offsets), A.ptr(ridx), A.ptr(samples), A.ptr(depth), A.ptr(deltas), A.ptr(boundary), A.stream()))
    return ridx, samples, depth, deltas, boundary


_CAP_FLOOR = 0      # high-water capacity of the per-sample buffers (samples); only grows


def reserve_samples(S: int) -> int:
    """Pre-size the per-sample buffers for batches of up to S hit samples (trainer warm-up: the reference sizes its ray batch
    from prev_num_samples the same way, multiview_trainer.py:95-109).  Returns the capacity now in force."""
    global _CAP_FLOOR
    _CAP_FLOOR = max(_CAP_FLOOR, _bucket_raw(int(S)))
    return _CAP_FLOOR


def _bucket_raw(S: int) -> int:
    if S <= 0:
        return 0
    gran = max(1 << 16, 1 << max(0, S.bit_length() - 4))
    return (S + gran - 1) // gran * gran


def _bucket(S: int) -> int:
    """Capacity for a per-sample buffer: S rounded up to 1/8 of its power of two (<= 12.5 % slack, >= 64 Ki samples), and never
    below the largest capacity handed out so far.  The sample count changes with every batch of rays; one stable size lets
    the caching allocator hand the same blocks back instead of going to cudaMalloc / cudaFree (a device sync, ~10 ms) whenever
    S crosses a bucket boundary.  A batch that outgrows the high-water mark raises it with 1/16 headroom."""
    global _CAP_FLOOR
    if S <= 0:
        return 0
    b = _bucket_raw(S)
    if b > _CAP_FLOOR:
        _CAP_FLOOR = _bucket_raw(S + S // 16)
    return _CAP_FLOOR


def _empty_s(S: int, tail: tuple, dtype, device):
    """torch.empty((S, *tail)) carved from a bucketed allocation."""
    return torch.empty((_bucket(S),) + tuple(tail), dtype=dtype, device=device)[:S]


def march_fill_records(ms: MarchState, device):
    """Fused-path sample records: depth t, delta, ray index (12 B/sample)."""
    if ms.records is not None:
        return ms.records
    S = ms.total
    rec_t = _empty_s(S, (), torch.float32, device)
    rec_delta = _empty_s(S, (), torch.float32, device)
    rec_ray = _empty_s(S, (), torch.int32, device)
    if S > 0:
        with _stage("march_fill"):
            A.check(A.lib().wb_rf_march_fill(C.byref(ms.rays), C.c_int32(ms.n), A.ptr(ms.jitter), C.c_uint32(ms.seed), A.ptr(ms.hitmask),
                                             A.ptr(ms.offsets), A.ptr(rec_t), A.ptr(rec_delta), A.ptr(rec_ray), A.stream()))
    return rec_t, rec_delta, rec_ray


# --------------------------------------------------------------------------------------------------------------
# raytrace + 'voxel' / 'uniform' samplers
# --------------------------------------------------------------------------------------------------------------
def _scan(counts: torch.Tensor) -> torch.Tensor:
    n = counts.shape[0]
    offsets = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    L = A.lib()
    wsb = int(L.wb_scan_workspace_bytes(C.c_int64(n)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=counts.device)
    A.check(L.wb_scan_counts(A.ptr(counts), C.c_int64(n), A.ptr(offsets), A.ptr(ws), C.c_int64(wsb), A.stream()))
    return offsets


def raytrace(oct: OctreeTensors, origins, dirs, level: int):
    """spc_render.unbatched_raytrace(..., return_depth=True, with_exit=True) (octree_as.py:183-185)
    -> ridx int32 [Ng], pidx int32 [Ng], depth f32 [Ng,2], ray_offsets int64 [R+1]."""
    A.require_device(origins)
    rays, keep = A.make_rays(origins, dirs, 0.0, 0.0)
    R, dev, L = rays.num_rays, origins.device, A.lib()
    od = oct.desc()
    counts = torch.empty(R, dtype=torch.int32, device=dev)
    with _stage("raytrace_count"):
        A.check(L.wb_raytrace_count(C.byref(od), C.c_int32(level), C.byref(rays), A.ptr(counts), A.stream()))
    offsets = _scan(counts)
    Ng = int(offsets[-1].item())
    ridx = torch.empty(Ng, dtype=torch.int32, device=dev); pidx = torch.empty(Ng, dtype=torch.int32, device=dev)
    depth = torch.empty((Ng, 2), dtype=torch.float32, device=dev)
    if Ng > 0:
        with _stage("raytrace_fill"):
            A.check(L.wb_raytrace_fill(C.byref(od), C.c_int32(level), C.byref(rays), A.ptr(offsets), A.ptr(ridx), A.ptr(pidx), A.ptr(depth), A.stream()))
    return ridx, pidx, depth, offsets


def uniform_scale(num_samples: int) -> int:
    """octree_as.py:336-338."""
    import math
    return int(math.ceil(1.0 / (2.0 * math.sqrt(3.0) / num_samples)))


def march_nuggets(oct: OctreeTensors, origins, dirs, level: int, num_samples: int, kind: str, reference_layout: bool,
                  jitter: Optional[torch.Tensor] = None, seed: int = 0):
    """'voxel' / 'uniform' sampling on top of raytrace.  Returns (MarchState with records, reference-layout dict or None)."""
    ridx_n, pidx_n, depth_n, ray_off = raytrace(oct, origins, dirs, level)
    rays, keep = A.make_rays(origins, dirs, 0.0, 0.0)
    R, dev, L = rays.num_rays, origins.device, A.lib()
    Ng = ridx_n.shape[0]
    if kind == 'voxel':
        S = Ng * num_samples
        sample_off = None
    else:
        scale = uniform_scale(num_samples)
        cnt = torch.empty(Ng, dtype=torch.int32, device=dev)
        A.check(L.wb_raymarch_uniform_count(A.ptr(depth_n), C.c_int64(Ng), C.c_int32(scale), A.ptr(cnt), A.stream()))
        sample_off = _scan(cnt)
        S = int(sample_off[-1].item())
    t = torch.empty(S, dtype=torch.float32, device=dev); dl = torch.empty(S, dtype=torch.float32, device=dev)
    rr = torch.empty(S, dtype=torch.int32, device=dev)
    ref = None
    ridx = samples = boundary = None
    if reference_layout:
        ridx = torch.empty(S, dtype=torch.int64, device=dev); samples = torch.empty((S, 3), dtype=torch.float32, device=dev)
        boundary = torch.empty(S, dtype=torch.bool, device=dev)
    jit = None if jitter is None else A.f32c(jitter)
    with _stage("march_" + kind):
        if kind == 'voxel':
            A.check(L.wb_raymarch_voxel_fill(C.byref(rays), A.ptr(ridx_n), A.ptr(depth_n), C.c_int64(Ng), C.c_int32(num_samples), A.ptr(jit),
                                             C.c_uint32(seed & 0xFFFFFFFF), A.ptr(ridx), A.ptr(samples), A.ptr(t), A.ptr(dl), A.ptr(boundary), A.ptr(rr), A.stream()))
            offsets = ray_off * num_samples
        else:
            A.check(L.wb_raymarch_uniform_fill(C.byref(rays), A.ptr(ridx_n), A.ptr(depth_n), C.c_int64(Ng), C.c_int32(scale), A.ptr(sample_off), A.ptr(ray_off),
                                               A.ptr(ridx), A.ptr(samples), A.ptr(t), A.ptr(dl), A.ptr(boundary), A.ptr(rr), A.stream()))
            offsets = sample_off[ray_off]
    ms = MarchState(rays, keep, num_samples, None, seed & 0xFFFFFFFF, None, None, offsets.contiguous(), S, records=(t, dl, rr))
    if reference_layout:
        ref = dict(ridx=ridx, samples=samples, depth_samples=t[:, None], deltas=dl[:, None], boundary=boundary)
    return ms, ref


# --------------------------------------------------------------------------------------------------------------
# hash grid interpolate (unfused drop-in for wisp.ops.grid.hashgrid)
# --------------------------------------------------------------------------------------------------------------
class HashGridInterpolate(torch.autograd.Function):
    """wisp.ops.grid.HashGridInterpolate (ops/grid.py:77-126) over wb_hashgrid_fwd / wb_hashgrid_bwd.
    Differences by design: all LODs in one launch; the table is read as fp32 master (no per-call .half() copy,
    ops/grid.py:88-89) and gradients accumulate in fp32."""

    @staticmethod
    def forward(ctx, coords, resolutions, codebook_bitwidth, lod_idx, codebook, codebook_first_idx):
        if codebook.shape[-1] % 2 == 1:
            raise Exception("The codebook feature dimension needs to be a multiple of 2.")   # ops/grid.py:83-84
        assert coords.shape[-1] == 3, "only the 3D hash grid is on the accelerated path"
        A.require_device(codebook)
        c = A.f32c(coords)
        res = [int(r) for r in torch.as_tensor(resolutions).reshape(-1).tolist()]
        begin = [int(b) for b in codebook_first_idx.tolist()] if torch.is_tensor(codebook_first_idx) else list(codebook_first_idx)
        table = A.f32c(codebook.detach())
        desc = A.make_grid_desc(table, res, begin, 2 ** codebook_bitwidth)
        feats = torch.empty((c.shape[0], len(res) * table.shape[1]), dtype=torch.float32, device=c.device)
        A.check(A.lib().wb_hashgrid_fwd(A.ptr(c), C.c_int64(c.shape[0]), C.byref(desc), A.ptr(feats), A.stream()))
        ctx.save_for_backward(c, table)
        ctx.meta = (res, begin, codebook_bitwidth)
        return feats

    @staticmethod
    def backward(ctx, grad_output):
        c, table = ctx.saved_tensors
        res, begin, bw = ctx.meta
        desc = A.make_grid_desc(table, res, begin, 2 ** bw)
        g = A.f32c(grad_output)
        gt = torch.zeros_like(table)                      # hashgrid_interpolate.cpp:85
        A.check(A.lib().wb_hashgrid_bwd(A.ptr(c), C.c_int64(c.shape[0]), C.byref(desc), A.ptr(g), A.ptr(gt), A.stream()))
        return None, None, None, None, gt, None


def hashgrid(coords, codebook_bitwidth, lod_idx, codebook):
    """wisp.ops.grid.hashgrid (ops/grid.py:128-144); `codebook` is a MultiTable."""
    batch, dim = coords.shape
    feats = HashGridInterpolate.apply(coords.contiguous(), codebook.resolutions, codebook_bitwidth, lod_idx,
                                      codebook.feats, codebook.begin_idxes)
    feature_dim = codebook.feats.shape[1] * len(codebook.resolutions)
    return feats.reshape(batch, feature_dim)


# --------------------------------------------------------------------------------------------------------------
# triplanar grid interpolate
# --------------------------------------------------------------------------------------------------------------
class TriplaneInterpolate(torch.autograd.Function):
    """All LODs x 3 planes of TriplanarGrid.interpolate in one launch (triplanar_grid.py:98-143, 205-223).
    planes: fmx, fmy, fmz of LOD 0, then LOD 1, ... each [1, fdim, res+1, res+1].  Returns [N, num_lods*3*fdim]."""

    @staticmethod
    def forward(ctx, coords, num_lods, *planes):
        A.require_device(planes[0])
        c = A.f32c(coords)
        N, fdim = c.shape[0], planes[0].shape[1]
        pl = [A.f32c(p.detach()) for p in planes[:3 * num_lods]]
        res = (C.c_int32 * num_lods)(*[pl[3 * l].shape[-1] - 1 for l in range(num_lods)])
        ptrs = (C.c_void_p * (3 * num_lods))(*[p.data_ptr() for p in pl])
        feats = torch.empty((N, num_lods * 3 * fdim), dtype=torch.float32, device=c.device)
        A.check(A.lib().wb_triplane_fwd(A.ptr(c), C.c_int64(N), C.c_int32(num_lods), C.c_int32(fdim), res, ptrs, A.ptr(feats), A.stream()))
        ctx.save_for_backward(c, *pl)
        ctx.num_lods, ctx.nplanes = num_lods, len(planes)
        return feats

    @staticmethod
    def backward(ctx, grad_output):
        c, *pl = ctx.saved_tensors
        num_lods, fdim = ctx.num_lods, pl[0].shape[1]
        g = A.f32c(grad_output)
        gp = [torch.zeros_like(p) for p in pl]
        res = (C.c_int32 * num_lods)(*[pl[3 * l].shape[-1] - 1 for l in range(num_lods)])
        ptrs = (C.c_void_p * (3 * num_lods))(*[p.data_ptr() for p in pl])
        gptrs = (C.c_void_p * (3 * num_lods))(*[p.data_ptr() for p in gp])
        A.check(A.lib().wb_triplane_bwd(A.ptr(c), C.c_int64(c.shape[0]), C.c_int32(num_lods), C.c_int32(fdim), res, ptrs, A.ptr(g), gptrs, A.stream()))
        return (None, None, *gp, *([None] * (ctx.nplanes - len(gp))))


# --------------------------------------------------------------------------------------------------------------
# octree grid interpolate + SDF-tracer helpers
# --------------------------------------------------------------------------------------------------------------
class OctreeInterpolate(torch.autograd.Function):
    """OctreeGrid.interpolate for LODs 0..lod_idx in one launch (octree_grid.py:165-219)."""

    @staticmethod
    def forward(ctx, coords, oct, trinkets, base_lod, multiscale, half_round, *feats):
        A.require_device(feats[0])
        c = A.f32c(coords)
        N, F, nl = c.shape[0], feats[0].shape[1], len(feats)
        fl = [A.f32c(f.detach()) for f in feats]
        ptrs = (C.c_void_p * nl)(*[f.data_ptr() for f in fl])
        out = torch.empty((N, F if multiscale == 'sum' else nl * F), dtype=torch.float32, device=c.device)
        od = oct.desc()
        A.check(A.lib().wb_octree_interp_fwd(C.byref(od), A.ptr(oct.points), A.ptr(trinkets), A.ptr(c), C.c_int64(N), C.c_int32(F), C.c_int32(base_lod),
                                             C.c_int32(nl), C.c_int32(1 if multiscale == 'sum' else 0), C.c_int32(int(half_round)), ptrs, A.ptr(out), A.stream()))
        ctx.save_for_backward(c, trinkets, *fl)
        ctx.meta = (oct, base_lod, multiscale)
        return out

    @staticmethod
    def backward(ctx, g):
        c, trinkets, *fl = ctx.saved_tensors
        oct, base_lod, multiscale = ctx.meta
        nl, F = len(fl), fl[0].shape[1]
        gf = [torch.zeros_like(f) for f in fl]
        ptrs = (C.c_void_p * nl)(*[f.data_ptr() for f in fl]); gptrs = (C.c_void_p * nl)(*[f.data_ptr() for f in gf])
        od = oct.desc()
        A.check(A.lib().wb_octree_interp_bwd(C.byref(od), A.ptr(oct.points), A.ptr(trinkets), A.ptr(c), C.c_int64(c.shape[0]), C.c_int32(F), C.c_int32(base_lod),
                                             C.c_int32(nl), C.c_int32(1 if multiscale == 'sum' else 0), ptrs, A.ptr(A.f32c(g)), gptrs, A.stream()))
        return (None, None, None, None, None, None, *gf)


def find_depth_bound(query: torch.Tensor, nug_depth: torch.Tensor, info=None, curr_idxes: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wisp.ops.geometric.find_depth_bound (geometric.py:15-22)."""
    if curr_idxes is None:
        curr_idxes = torch.nonzero(info)[..., 0].int()
    A.require_device(query)
    q, ci, dp = A.f32c(query).reshape(-1), curr_idxes.int().contiguous(), A.f32c(nug_depth)
    out = torch.empty(q.shape[0], dtype=torch.int32, device=q.device)
    A.check(A.lib().wb_find_depth_bound(A.ptr(q), A.ptr(ci), A.ptr(dp), C.c_int64(q.shape[0]), C.c_int64(dp.shape[0]), A.ptr(out), A.stream()))
    return out


def finitediff_gradient(x: torch.Tensor, f, eps: float = 0.005) -> torch.Tensor:
    """wisp.ops.differential.finitediff_gradient (gradients.py:29-45)."""
    ex = torch.tensor([eps, 0.0, 0.0], device=x.device); ey = torch.tensor([0.0, eps, 0.0], device=x.device); ez = torch.tensor([0.0, 0.0, eps], device=x.device)
    grad = torch.cat([f(x + ex) - f(x - ex), f(x + ey) - f(x - ey), f(x + ez) - f(x - ez)], dim=-1)
    return grad / (eps * 2.0)


# --------------------------------------------------------------------------------------------------------------
# packed compositing
# --------------------------------------------------------------------------------------------------------------
def _bg3(bg) -> "C.Array":
    v = [float(x) for x in (bg.detach().cpu().reshape(-1).tolist() if torch.is_tensor(bg) else bg)]
    return (C.c_float * 3)(*v[:3])


class CompositeFn(torch.autograd.Function):
    """exponential_integration + sum_reduce + per-ray scatter of PackedRFTracer.trace (packed_rf_tracer.py:136-165).
    shaded [S,4] = (r,g,b,sigma); returns rgb [R,3], depth [R,1], alpha [R,1], hit [R] (bool)."""

    @staticmethod
    def forward(ctx, shaded, depth, deltas, offsets, bg):
        A.require_device(shaded)
        R = offsets.shape[0] - 1
        dev = shaded.device
        sh, dp, dl = A.f32c(shaded), A.f32c(depth).reshape(-1), A.f32c(deltas).reshape(-1)
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
        dout = torch.empty((R, 1), dtype=torch.float32, device=dev)
        alpha = torch.empty((R, 1), dtype=torch.float32, device=dev)
        hit = torch.empty(R, dtype=torch.bool, device=dev)
        bgv = _bg3(bg)
        A.check(A.lib().wb_composite_fwd(A.ptr(sh), A.ptr(dp), A.ptr(dl), A.ptr(offsets), C.c_int64(R), bgv,
                                         A.ptr(rgb), A.ptr(dout), A.ptr(alpha), A.ptr(hit), A.stream()))
        ctx.save_for_backward(sh, dp, dl, offsets)
        ctx.bg = bgv
        ctx.mark_non_differentiable(hit)
        return rgb, dout, alpha, hit

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_alpha, _g_hit):
        sh, dp, dl, offsets = ctx.saved_tensors
        R = offsets.shape[0] - 1
        g_sh = torch.zeros_like(sh)
        A.check(A.lib().wb_composite_bwd(A.ptr(sh), A.ptr(dp), A.ptr(dl), A.ptr(offsets), C.c_int64(R), ctx.bg,
                                         A.ptr(A.f32c(g_rgb)), A.ptr(A.f32c(g_depth).reshape(-1)) if g_depth is not None else None,
                                         A.ptr(A.f32c(g_alpha).reshape(-1)) if g_alpha is not None else None, A.ptr(g_sh), None, A.stream()))
        return g_sh, None, None, None, None


# --------------------------------------------------------------------------------------------------------------
# fused render path
# --------------------------------------------------------------------------------------------------------------
@dataclass
class NefSpec:
    """Static description of a NeuralRadianceField(HashGrid) that the fused path supports."""
    resolutions: List[int]
    begin_idxes: List[int]
    codebook_size: int
    feature_dim: int
    multiscale: str
    lod_idx: int
    pos_mode: int
    pos_freq: int
    view_mode: int
    view_freq: int
    has_bias: bool
    dens_dims: List[int]
    col_dims: List[int]

    def desc(self, table: torch.Tensor, dens_flat: torch.Tensor, col_flat: torch.Tensor) -> A.NefDesc:
        d = A.make_grid_desc(table, self.resolutions, self.begin_idxes, self.codebook_size, self.multiscale, self.lod_idx)
        d.pos_mode, d.pos_freq, d.view_mode, d.view_freq = self.pos_mode, self.pos_freq, self.view_mode, self.view_freq
        d.has_bias = int(self.has_bias)
        d.dens_layers = len(self.dens_dims) - 1
        d.col_layers = len(self.col_dims) - 1
        if d.dens_layers > A.WB_MAX_LAYERS or d.col_layers > A.WB_MAX_LAYERS:
            raise A.WispB200Error("decoder too deep for the fused path")
        for i, v in enumerate(self.dens_dims):
            d.dens_dims[i] = v
        for i, v in enumerate(self.col_dims):
            d.col_dims[i] = v
        d.dens_params, d.col_params = dens_flat.data_ptr(), col_flat.data_ptr()
        return d


def _flatten(params: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1).float() for p in params]) if params else torch.zeros(0)


class RFTraceFn(torch.autograd.Function):
    """PackedRFTracer.trace + NeuralRadianceField.rgba + HashGrid.interpolate as one native pipeline:
         march(count/scan/fill) -> shade (gather + decoders fused) -> composite        (forward)
         composite_bwd -> shade_bwd (decoder recompute, table scatter)                  (backward)
    Inputs: table, then the decoder parameters in packing order [W0, b0?, W1, b1?, ...] for density then colour.
    """

    @staticmethod
    def forward(ctx, ms: MarchState, spec: NefSpec, n_dens: int, bg, precision: int, want_grad: bool, table, *params):
        A.require_device(table)
        L = A.lib()
        dev = table.device
        tb = A.f32c(table.detach())
        dens_flat, col_flat = _flatten(params[:n_dens]), _flatten(params[n_dens:])
        desc = spec.desc(tb, dens_flat, col_flat)
        nblob = int(L.wb_rf_param_blob_floats(C.byref(desc), C.c_int32(precision)))
        if nblob < 0:
            raise A.WispB200Error(L.wb_last_error().decode())
        blob = torch.empty(nblob, dtype=torch.float32, device=dev)
        A.check(L.wb_rf_pack_params(C.byref(desc), C.c_int32(precision), A.ptr(blob), A.stream()))
        rec_t, rec_delta, rec_ray = march_fill_records(ms, dev)
        S, R = ms.total, ms.rays.num_rays
        shaded = _empty_s(S, (4,), torch.float32, dev)
        # grad mode is always off inside Function.forward and needs_input_grad ignores torch.no_grad(): the caller (rf_trace)
        # tells us whether a backward pass can follow, so that inference neither saves features nor needs the backward tiles
        need_grad = bool(want_grad) and any(ctx.needs_input_grad[6:])
        Scap = _bucket(S)                                  # byte sizes for the bucketed capacity (layouts still use S)
        wsb = int(L.wb_rf_workspace_bytes(C.byref(desc), C.c_int32(precision), C.c_int64(R), C.c_int64(Scap), C.c_int32(0)))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb > 0 else None
        fb = int(L.wb_rf_feat_bytes(C.byref(desc), C.c_int32(precision), C.c_int64(Scap))) if need_grad else 0
        if fb < 0 or wsb < 0:        # e.g. the tensor-core backward of this decoder does not fit: fail at the forward, not mid-backward
            raise A.WispB200Error(L.wb_last_error().decode())
        feat = torch.empty(fb, dtype=torch.uint8, device=dev) if fb > 0 else None
        with _stage("shade_fwd"):
            A.check(L.wb_rf_shade_fwd(C.byref(desc), A.ptr(blob), C.c_int32(precision), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray),
                                      C.c_int64(S), A.ptr(shaded), A.ptr(feat), A.ptr(ws), A.stream()))
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
        depth = torch.empty((R, 1), dtype=torch.float32, device=dev)
        alpha = torch.empty((R, 1), dtype=torch.float32, device=dev)
        hit = torch.empty(R, dtype=torch.bool, device=dev)
        bgv = _bg3(bg)
        with _stage("composite_fwd"):
            A.check(L.wb_composite_fwd(A.ptr(shaded), A.ptr(rec_t), A.ptr(rec_delta), A.ptr(ms.offsets), C.c_int64(R), bgv,
                                       A.ptr(rgb), A.ptr(depth), A.ptr(alpha), A.ptr(hit), A.stream()))
        ctx.ms, ctx.spec, ctx.n_dens, ctx.bg, ctx.precision = ms, spec, n_dens, bgv, precision
        ctx.param_shapes = [p.shape for p in params]
        ctx.feat = feat
        ctx.save_for_backward(tb, dens_flat, col_flat, blob, rec_t, rec_delta, rec_ray, shaded)
        ctx.mark_non_differentiable(hit)
        return rgb, depth, alpha, hit

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_alpha, _g_hit):
        tb, dens_flat, col_flat, blob, rec_t, rec_delta, rec_ray, shaded = ctx.saved_tensors
        ms, spec = ctx.ms, ctx.spec
        L = A.lib()
        S, R = ms.total, ms.rays.num_rays
        desc = spec.desc(tb, dens_flat, col_flat)
        g_sh = _empty_s(S, (4,), torch.float32, shaded.device)
        gd = A.f32c(g_depth).reshape(-1) if g_depth is not None else None
        ga = A.f32c(g_alpha).reshape(-1) if g_alpha is not None else None
        grgb = A.f32c(g_rgb)
        absmax = torch.zeros(1, dtype=torch.float32, device=shaded.device) if ctx.precision == 1 else None
        with _stage("composite_bwd"):
            A.check(L.wb_composite_bwd(A.ptr(shaded), A.ptr(rec_t), A.ptr(rec_delta), A.ptr(ms.offsets), C.c_int64(R), ctx.bg,
                                       A.ptr(grgb), A.ptr(gd), A.ptr(ga), A.ptr(g_sh), A.ptr(absmax), A.stream()))
        g_table = torch.zeros_like(tb)
        g_dens = torch.zeros_like(dens_flat)
        g_col = torch.zeros_like(col_flat)
        scale = None
        if ctx.precision == 1 and S > 0:
            # power-of-two loss scale computed on the device (no host sync): largest |gradient| -> ~64 in fp16
            amax = absmax[0].clamp_min(1e-30)              # max |g_shaded|, gathered by the compositing backward itself
            # 2^k assembled from the exponent bits (torch.exp2 is a jiterator op: NVRTC compile at first use)
            k = torch.floor(torch.log2(64.0 / amax)).clamp(-20.0, 60.0).to(torch.int32)
            scale = ((k + 127) << 23).view(torch.float32).reshape(1).contiguous()
        wsb = int(L.wb_rf_workspace_bytes(C.byref(desc), C.c_int32(ctx.precision), C.c_int64(R), C.c_int64(_bucket(S)), C.c_int32(1)))
        ws = torch.empty(wsb, dtype=torch.uint8, device=tb.device) if wsb > 0 else None
        if ctx.precision == 1 and S > 0:
            with _stage("decoder_bwd"):
                A.check(L.wb_rf_decoder_bwd(C.byref(desc), A.ptr(blob), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray), C.c_int64(S), A.ptr(g_sh),
                                            A.ptr(scale), A.ptr(ctx.feat), A.ptr(ws), A.ptr(g_dens), A.ptr(g_col), A.stream()))
            with _stage("table_scatter"):
                A.check(L.wb_rf_table_scatter(C.byref(desc), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray), C.c_int64(S), A.ptr(scale), A.ptr(ws),
                                              A.ptr(g_table), A.stream()))
        else:
            with _stage("shade_bwd"):
                A.check(L.wb_rf_shade_bwd(C.byref(desc), A.ptr(blob), C.c_int32(ctx.precision), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray),
                                          C.c_int64(S), A.ptr(g_sh), A.ptr(scale), A.ptr(ctx.feat), A.ptr(ws), A.ptr(g_table), A.ptr(g_dens), A.ptr(g_col), A.stream()))
        grads = []
        for flat, shapes in ((g_dens, ctx.param_shapes[:ctx.n_dens]), (g_col, ctx.param_shapes[ctx.n_dens:])):
            o = 0
            for shp in shapes:
                n = int(torch.Size(shp).numel())
                grads.append(flat[o:o + n].reshape(shp)); o += n
        return (None, None, None, None, None, None, g_table, *grads)


def precision_supported(spec, nef, precision: int, backward: bool) -> bool:
    """Host-side query (wb_rf_precision_supported): can this decoder configuration run at `precision`?"""
    tb = A.f32c(nef.grid.codebook.feats.detach())
    dens, col = _flatten(nef.decoder_density.packed_params()), _flatten(nef.decoder_color.packed_params())
    desc = spec.desc(tb, dens, col)
    return bool(A.lib().wb_rf_precision_supported(C.byref(desc), C.c_int32(precision), C.c_int32(1 if backward else 0)))


def rf_trace(ms: MarchState, spec: NefSpec, table: torch.Tensor, dens_params: Sequence[torch.Tensor],
             col_params: Sequence[torch.Tensor], bg, precision: int = 0):
    """-> rgb [R,3], depth [R,1], alpha [R,1], hit [R]."""
    return RFTraceFn.apply(ms, spec, len(dens_params), bg, precision, torch.is_grad_enabled(), table, *dens_params, *col_params)
