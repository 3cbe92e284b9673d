"""install(): route an importable kaolin-wisp through this path by overriding methods on wisp's own classes
(INTEGRATION.md section 3).  Nothing is patched unless `wisp` imports; WISP_B200=0 disables the hook."""
from __future__ import annotations

import os

import torch

from . import ops


def _octree_tensors(blas) -> ops.OctreeTensors:
    t = getattr(blas, "_wb_tensors", None)
    if t is None or t.octree.data_ptr() != blas.octree.data_ptr():
        t = ops.OctreeTensors(blas.octree.contiguous(), blas.prefix.contiguous().int(), blas.points.contiguous(), blas.pyramid.cpu().int(), blas.max_level)
        blas._wb_tensors = t
    return t


def install() -> bool:
    if os.environ.get("WISP_B200", "1") == "0":
        return False
    try:
        import wisp.ops.grid as grid_ops
        from wisp.accelstructs import OctreeAS
        from wisp.accelstructs.base_as import ASQueryResults, ASRaymarchResults
        from wisp.core import RenderBuffer
        from wisp.models.grids import HashGrid
        from wisp.models.nefs import NeuralRadianceField
        from wisp.tracers import PackedRFTracer
    except Exception:          # wisp (or one of its dependencies) is not importable here
        return False

    grid_ops.hashgrid = ops.hashgrid

    def query(self, coords, level=None, with_parents=False):
        return ASQueryResults(pidx=ops.query(_octree_tensors(self), coords, self.max_level if level is None else level, with_parents))

    def _raymarch_ray(self, rays, num_samples, level=None):
        ms = ops.march_count(_octree_tensors(self), rays.origins, rays.dirs, rays.dist_min, rays.dist_max, num_samples,
                             self.max_level if level is None else level, seed=int(torch.randint(0, 2 ** 31 - 1, (1,))))
        ridx, samples, depth, deltas, boundary = ops.march_fill_reference_layout(ms, rays.origins.device)
        return ASRaymarchResults(ridx=ridx, samples=samples, depth_samples=depth, deltas=deltas, boundary=boundary, pack_info=None)

    OctreeAS.query = query
    OctreeAS._raymarch_ray = _raymarch_ray

    orig_trace = PackedRFTracer.trace

    def _spec(nef, lod_idx):
        g = nef.grid
        if not isinstance(nef, NeuralRadianceField) or not isinstance(g, HashGrid) or nef.activation_type != 'relu' or nef.layer_type not in ('linear', 'none'):
            return None
        from .nefs import NeuralRadianceField as Mirror
        pm, pf = Mirror._embed_mode(nef.pos_embedder_type, getattr(nef, "position_input", False) or nef.pos_embed_dim in (3, 3 + 6 * 10), 10)
        vm, vf = Mirror._embed_mode(nef.view_embedder_type, True, (nef.view_embed_dim - 3) // 6 if nef.view_embed_dim > 3 else 0)
        dims = lambda d: [d.input_dim] + [d.hidden_dim] * d.num_layers + [d.output_dim]
        return ops.NefSpec(resolutions=[int(r) for r in g.resolutions], begin_idxes=[int(b) for b in g.codebook.begin_idxes.tolist()],
                           codebook_size=g.codebook_size, feature_dim=g.feature_dim, multiscale=g.multiscale_type, lod_idx=int(lod_idx),
                           pos_mode=pm, pos_freq=pf, view_mode=vm, view_freq=vf, has_bias=bool(nef.bias),
                           dens_dims=dims(nef.decoder_density), col_dims=dims(nef.decoder_color))

    def _packed(dec):
        out = []
        for l in list(dec.layers) + [dec.lout]:
            out.append(l.weight)
            if l.bias is not None:
                out.append(l.bias)
        return out

    def trace(self, nef, rays, channels, extra_channels, lod_idx=None, raymarch_type='voxel', num_steps=64, step_size=1.0, bg_color='white'):
        if lod_idx is None:
            lod_idx = nef.grid.num_lods - 1
        spec = _spec(nef, lod_idx) if (raymarch_type == 'ray' and not extra_channels and nef.pos_embedder is None) else None
        if spec is None:
            return orig_trace(self, nef, rays, channels, extra_channels, lod_idx=lod_idx, raymarch_type=raymarch_type,
                              num_steps=num_steps, step_size=step_size, bg_color=bg_color)
        self.bg_color = self.bg_color.to(rays.origins.device)
        blas = nef.grid.blas
        ms = ops.march_count(_octree_tensors(blas), rays.origins, rays.dirs, rays.dist_min, rays.dist_max, num_steps, blas.max_level,
                             seed=int(torch.randint(0, 2 ** 31 - 1, (1,))))
        self.prev_num_samples = ms.total
        precision = 1 if torch.is_autocast_enabled() else 0
        rgb, depth, alpha, hit = ops.rf_trace(ms, spec, nef.grid.codebook.feats, _packed(nef.decoder_density), _packed(nef.decoder_color),
                                              self.bg_color, precision=precision)
        return RenderBuffer(depth=depth if "depth" in channels else None, hit=hit, rgb=rgb, alpha=alpha)

    PackedRFTracer.trace = trace
    return True
