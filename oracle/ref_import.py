"""Import the UNMODIFIED reference Python (wisp) in this container (TEST INFRASTRUCTURE).

/root/reference cannot be imported as is: `kaolin` (un-vendored pip dependency, INSTALL.md:8,14),
the compiled extension `wisp._C` and a handful of config/GUI packages are absent.  This module
installs import stubs so that the reference's *Python* hot path --
    wisp.models.Pipeline, wisp.tracers.PackedRFTracer, wisp.accelstructs.OctreeAS,
    wisp.models.grids.HashGrid, wisp.models.nefs.NeuralRadianceField, wisp.ops.grid --
runs on CPU exactly as written, with only the external native calls (Kaolin SPC ops and the
wisp._C kernels) answered by the oracle's definitions.  oracle/make_golden.py uses it to generate
tests/golden/*.npz; tests then compare oracle and CUDA outputs with what the reference glue produced.

Only usable where /root/reference exists (the build container), never on the GPU box.
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import torch

from . import oracle as O

REF_ROOT = "/root/reference"
_STUB_ROOTS = ("kaolin", "hydra_zen", "hydra", "omegaconf", "attrdict", "skimage", "matplotlib", "tinyobjloader",
               "polyscope", "imgui", "glumpy", "OpenGL", "pycuda", "cuda", "wandb", "tensorboard", "lpips", "pyexr",
               "cv2", "PIL", "tyro", "docstring_parser", "apex", "tinycudann", "glfw", "plyfile", "pandas_stub",
               "torchvision", "scipy_stub", "tqdm_stub", "moviepy", "pydispatch")


# wisp sub-packages that are off the hot path and do not import under this interpreter
# (wisp/framework/state.py uses py3.8-era mutable dataclass defaults; renderer/gfx need OpenGL).
_STUB_WISP = ("wisp._C", "wisp.framework", "wisp.renderer", "wisp.gfx", "wisp.config", "wisp.trainers", "wisp.datasets")


class _Dummy:
    """Stand-in for any attribute of a stubbed module (usable as base class, decorator or callable)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # behave as a transparent decorator
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()

    def __mro_entries__(self, bases):
        return (object,)

    def __iter__(self):
        return iter(())

    def __or__(self, o):
        return self

    def __ror__(self, o):
        return self

    def __getitem__(self, k):
        return _Dummy()


class _Meta(type):
    """Metaclass of stub classes: unknown class attributes resolve to dummies (e.g. dispatcher.send)."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return _Meta(name, (), {"__init__": lambda self, *a, **k: None,
                               "__call__": lambda self, *a, **k: _Dummy(),
                               "__class_getitem__": classmethod(lambda cls, k: cls)})


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS or any(fullname == m or fullname.startswith(m + ".") for m in _STUB_WISP):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ------------------------------------------------------------------------------------------------
# Oracle-backed definitions of the external native calls  
# ------------------------------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


def _spc_from(octree):
    return O.octree_to_spc(_np(octree).astype(np.uint8))


def _unbatched_query(octree, prefix, coords, level, with_parents=False):
    spc = _spc_from(octree)
    return torch.from_numpy(O.query(spc, _np(coords).astype(np.float32), level, with_parents))


def _scan_octrees(octree, lengths):
    spc = _spc_from(octree)
    return spc.max_level, torch.from_numpy(spc.pyramid.astype(np.int32))[None], torch.from_numpy(spc.prefix.copy())


def _generate_points(octree, pyramid, prefix):
    return torch.from_numpy(_spc_from(octree).points.copy())


def _unbatched_get_level_points(points, pyramid, level):
    return points[pyramid[1, level]: pyramid[1, level] + pyramid[0, level]]


def _unbatched_points_to_octree(points, level, sorted=False):
    return torch.from_numpy(O.points_to_octree(_np(points), level))


def _points_to_corners(p):
    offs = torch.tensor([[(j >> 2) & 1, (j >> 1) & 1, j & 1] for j in range(8)], dtype=p.dtype)
    return p[..., None, :] + offs


def _mark_pack_boundaries(ridx):
    b = torch.ones_like(ridx, dtype=torch.bool)
    if ridx.shape[0] > 1:
        b[1:] = ridx[1:] != ridx[:-1]
    return b


class _SumReduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, boundary):
        pid = torch.cumsum(boundary.long(), 0) - 1
        ctx.save_for_backward(pid)
        out = torch.from_numpy(O.sum_reduce(_np(feats), _np(boundary)))
        return out

    @staticmethod
    def backward(ctx, g):
        (pid,) = ctx.saved_tensors
        return g[pid], None


def _sum_reduce(feats, boundary):
    return _SumReduce.apply(feats.contiguous(), boundary)


def _cumsum(feats, boundary, exclusive=False, reverse=False):
    from . import torch_twin as TW
    if reverse:
        raise NotImplementedError
    return TW.cumsum_pack(feats, boundary, exclusive)


def _exponential_integration(feats, tau, boundary, exclusive=True):
    alpha = 1.0 - torch.exp(-tau.contiguous())
    transmittance = torch.exp(-1.0 * _cumsum(tau.contiguous(), boundary.contiguous(), exclusive=exclusive))
    transmittance = transmittance * alpha
    feats_out = _sum_reduce(transmittance * feats.contiguous(), boundary.contiguous())
    return feats_out, transmittance


def _hashgrid_interpolate_cuda(coords, codebook, codebook_first_idx, resolution, codebook_bitwidth):
    res = [int(r) for r in resolution.reshape(-1)]
    return torch.from_numpy(O.hashgrid_fwd(_np(coords), _np(codebook).astype(np.float32), res, codebook_bitwidth))


def _hashgrid_interpolate_backward_cuda(coords, grad_output, codebook, codebook_first_idx, resolution,
                                        codebook_bitwidth, feature_dim, require_grad_coords):
    res = [int(r) for r in resolution.reshape(-1)]
    g = O.hashgrid_bwd(_np(coords), _np(grad_output), codebook.shape[0], res, codebook_bitwidth)
    return [torch.empty(0), torch.from_numpy(g)]


def _unbatched_raytrace(octree, points, pyramid, prefix, origins, dirs, level, return_depth=True, with_exit=False):
    rt = O.raytrace(_spc_from(octree), _np(origins), _np(dirs), level)
    depth = rt["depth"] if with_exit else rt["depth"][:, :1]
    return torch.from_numpy(rt["ridx"]), torch.from_numpy(rt["pidx"]), torch.from_numpy(np.ascontiguousarray(depth))


def _uniform_sample_cuda(scale, ridx, depth, insum):
    """wisp/csrc/ops/uniform_sample_cuda.cu:18-59 restated with numpy (one 'thread' per nugget)."""
    ridx_n, depth_n, insum_n = _np(ridx), _np(depth), _np(insum)
    V = ridx_n.shape[0]
    total = int(insum_n[-1]) if V > 0 else 0
    new_ridx = np.zeros(total, np.int64); ds = np.zeros((total, 1), np.float32); boundary = np.zeros(total, bool)
    inv_scale = np.float32(1.0) / np.float32(scale)
    for t in range(V):
        base = int(insum_n[t - 1]) if t > 0 else 0
        n = int(insum_n[t]) - base
        first = np.ceil(np.float32(scale) * depth_n[t, 0]).astype(np.float32)
        bval = True if t == 0 else bool(ridx_n[t] != ridx_n[t - 1])
        f = np.float32(0.0)
        for i in range(n):
            ds[base + i, 0] = inv_scale * (first + f); f += np.float32(1.0)
            new_ridx[base + i] = ridx_n[t]; boundary[base + i] = bval; bval = False
    return [torch.from_numpy(new_ridx), torch.from_numpy(ds), torch.from_numpy(boundary)]


class _SpcView:
    def __init__(self, points, pyramid):
        self.points = _np(points); self.pyramid = _np(pyramid).astype(np.int64); self.max_level = self.pyramid.shape[-1] - 2


def _unbatched_make_dual(points, pyramid):
    from . import octree_grid as OG
    pd, pyr, _, _ = OG.make_trilinear_spc(_SpcView(points, pyramid))
    return torch.from_numpy(pd), torch.from_numpy(pyr.astype(np.int32))


def _unbatched_make_trinkets(points, pyramid, points_dual, pyramid_dual):
    from . import octree_grid as OG
    _, _, tr, par = OG.make_trilinear_spc(_SpcView(points, pyramid))
    return torch.from_numpy(tr), torch.from_numpy(par)


class _InterpTrilinear(torch.autograd.Function):
    """kaolin.ops.spc.unbatched_interpolate_trilinear: fwd via the oracle, bwd = index_add of coef*grad."""

    @staticmethod
    def forward(ctx, coords, pidx, points, trinkets, feats, level):
        from . import octree_grid as OG
        N, Ns = coords.shape[:2]
        c = _np(coords).reshape(-1, 3).astype(np.float32)
        p = np.repeat(_np(pidx).astype(np.int64), Ns)
        half = feats.dtype == torch.float16
        out = OG.interpolate_trilinear(c, p, _np(points), _np(trinkets), _np(feats).astype(np.float32), level, half=half)
        ctx.meta = (c, p, _np(points), _np(trinkets), level, feats.shape, feats.dtype)
        return torch.from_numpy(out).reshape(N, Ns, -1).to(feats.dtype)

    @staticmethod
    def backward(ctx, g):
        from . import octree_grid as OG
        c, p, points, trinkets, level, shape, dtype = ctx.meta
        gf = np.zeros(shape, np.float32)
        ok = p >= 0
        gn = _np(g).reshape(-1, shape[1]).astype(np.float32)
        if ok.any():
            cf = OG.trilinear_coeffs(c[ok], points[p[ok]], level)
            for j in range(8):
                np.add.at(gf, trinkets[p[ok], j].astype(np.int64), gn[ok] * cf[:, j:j + 1])
        return None, None, None, None, torch.from_numpy(gf).to(dtype), None


def _unbatched_interpolate_trilinear(coords, pidx, points, trinkets, feats, level):
    return _InterpTrilinear.apply(coords, pidx, points, trinkets, feats, level)


def _coords_to_trilinear_coeffs(coords, points, level):
    """kaolin.ops.spc.coords_to_trilinear_coeffs: [..., 3] coords and cell points -> [..., 8] coefficients."""
    from . import octree_grid as OG
    shp = coords.shape[:-1]
    cf = OG.trilinear_coeffs(_np(coords).reshape(-1, 3).astype(np.float32), _np(points).reshape(-1, 3), level)
    return torch.from_numpy(cf).reshape(*shp, 8)


def _find_depth_bound_cuda(query, curr_idxes, depth):
    from . import octree_grid as OG
    return torch.from_numpy(OG.find_depth_bound(_np(query), _np(curr_idxes), _np(depth)))


_installed = False


def install():
    """Make `import wisp` work against /root/reference with oracle-backed externals."""
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # wisp/__init__.py eagerly imports every sub-package (GUI, config, trainers...): register the package
    # shell ourselves and import only the hot-path sub-packages from the unmodified sources.
    pkg = types.ModuleType("wisp")
    pkg.__path__ = [REF_ROOT + "/wisp"]
    pkg.__file__ = REF_ROOT + "/wisp/__init__.py"
    sys.modules["wisp"] = pkg
    import kaolin.ops.spc as spc_ops          # noqa: stubs
    import kaolin.render.spc as spc_render
    import kaolin._C.render.spc as kaolin_C_render_spc
    import wisp._C as wisp_C
    import wisp._C.ops as wisp_C_ops
    import wisp._C.render as wisp_C_render
    spc_ops.unbatched_query = _unbatched_query
    spc_ops.scan_octrees = _scan_octrees
    spc_ops.generate_points = _generate_points
    spc_ops.unbatched_get_level_points = _unbatched_get_level_points
    spc_ops.unbatched_points_to_octree = _unbatched_points_to_octree
    spc_ops.points_to_corners = _points_to_corners
    spc_render.mark_pack_boundaries = _mark_pack_boundaries
    spc_render.mark_first_hit = _mark_pack_boundaries
    spc_render.unbatched_raytrace = _unbatched_raytrace
    kaolin_C_render_spc.inclusive_sum_cuda = lambda t: torch.cumsum(t, 0).int()
    wisp_C_ops.uniform_sample_cuda = _uniform_sample_cuda
    spc_ops.unbatched_make_dual = _unbatched_make_dual
    spc_ops.unbatched_make_trinkets = _unbatched_make_trinkets
    spc_ops.unbatched_interpolate_trilinear = _unbatched_interpolate_trilinear
    spc_ops.coords_to_trilinear_coeffs = _coords_to_trilinear_coeffs
    wisp_C_render.find_depth_bound_cuda = _find_depth_bound_cuda
    wisp_C.render = wisp_C_render
    spc_render.sum_reduce = _sum_reduce
    spc_render.cumsum = _cumsum
    spc_render.exponential_integration = _exponential_integration
    wisp_C_ops.hashgrid_interpolate_cuda = _hashgrid_interpolate_cuda
    wisp_C_ops.hashgrid_interpolate_backward_cuda = _hashgrid_interpolate_backward_cuda
    wisp_C.ops = wisp_C_ops
    _installed = True
