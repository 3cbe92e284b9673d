"""PackedRFTracer: host-side mirror of wisp.tracers.PackedRFTracer (wisp/tracers/packed_rf_tracer.py:20-181) and
wisp.tracers.BaseTracer.forward (wisp/tracers/base_tracer.py:99-162).  trace() runs the fused native pipeline when
the neural field is a NeuralRadianceField(HashGrid) with the 'ray' sampler, and otherwise the unfused route
(native raymarch + nef forward + native compositing)."""
from __future__ import annotations

import inspect

import torch
import torch.nn as nn

from . import ops
from .core import RenderBuffer


class PackedRFTracer(nn.Module):
    def __init__(self, raymarch_type='ray', num_steps=1024, step_size=1.0, bg_color=(1.0, 1.0, 1.0)):
        super().__init__()
        self.raymarch_type, self.num_steps, self.step_size = raymarch_type, num_steps, step_size
        self.bg_color = torch.tensor(bg_color, dtype=torch.float32)
        self.prev_num_samples = None
        self.precision = 0          # 0: fp32 decoders (autocast off), 1: fp16 tensor-core decoders (autocast on)
        self.seed = 0               # base of the counter-based jitter stream; advanced once per trace() call
        self.jitter = None          # optional explicit [R, num_steps] jitter (parity tests)

    def get_prev_num_samples(self):
        return self.prev_num_samples

    def get_supported_channels(self):
        return {"depth", "hit", "rgb", "alpha"}

    def get_required_nef_channels(self):
        return {"rgb", "density"}

    def forward(self, nef, rays, channels=None, **kwargs):
        """base_tracer.py:99-162: channel negotiation, kwargs default to tracer attributes of the same name."""
        nef_channels = nef.get_supported_channels()
        unsupported_inputs = self.get_required_nef_channels() - nef_channels
        if unsupported_inputs:
            raise Exception(f"The neural field class {type(nef)} does not output the required channels {unsupported_inputs}.")
        requested = self.get_supported_channels() if channels is None else {channels} if isinstance(channels, str) else set(channels)
        extra = requested - self.get_supported_channels()
        unsupported_outputs = extra - nef_channels
        if unsupported_outputs:
            raise Exception(f"Channels {unsupported_outputs} are not supported in the tracer {type(self)} or neural field {type(nef)}.")
        input_args = {}
        for a in list(inspect.signature(self.trace).parameters)[4:]:
            if a in kwargs:
                input_args[a] = kwargs[a]
            else:
                d = getattr(self, a, None)
                if d is not None:
                    input_args[a] = d
        return self.trace(nef, rays, requested, extra, **input_args)

    def trace(self, nef, rays, channels, extra_channels, lod_idx=None, raymarch_type='voxel', num_steps=64, step_size=1.0, bg_color='white'):
        """packed_rf_tracer.py:84-181.  Like the reference, the body reads self.bg_color, not the bg_color argument."""
        assert nef.grid is not None and "this tracer requires a grid"
        N = rays.origins.shape[0]
        if lod_idx is None:
            lod_idx = nef.grid.num_lods - 1
        dev = rays.origins.device
        self.bg_color = self.bg_color.to(dev)
        jitter, seed = self.jitter, self.seed
        self.seed = (self.seed + 1) & 0x7FFFFFFF
        spec = nef.fused_spec(lod_idx) if hasattr(nef, "fused_spec") else None
        if raymarch_type not in ('ray', 'voxel', 'uniform'):
            raise TypeError(f"Raymarch sampler type: {raymarch_type} is not supported by OctreeAS.")       # octree_as.py:427
        if spec is not None and not extra_channels:
            blas = nef.grid.blas
            if raymarch_type == 'ray':
                ms = ops.march_count(blas.tensors(), rays.origins, rays.dirs, rays.dist_min, rays.dist_max, num_steps, blas.max_level,
                                     jitter=jitter, seed=seed)
            else:
                ms, _ = ops.march_nuggets(blas.tensors(), rays.origins, rays.dirs, blas.max_level, num_steps, raymarch_type,
                                          reference_layout=False, jitter=jitter, seed=seed)
            self.prev_num_samples = ms.total
            rgb, depth, alpha, hit = ops.rf_trace(ms, spec, nef.grid.codebook.feats, nef.decoder_density.packed_params(),
                                                  nef.decoder_color.packed_params(), self.bg_color, precision=self.precision)
            return RenderBuffer(depth=depth if "depth" in channels else None, hit=hit, rgb=rgb, alpha=alpha)

        # ---- unfused route: same operators, nef evaluated through its own forward() ----
        mr = nef.grid.raymarch(rays, level=nef.grid.active_lods[lod_idx], num_samples=num_steps, raymarch_type=raymarch_type,
                               jitter=jitter, seed=seed)
        ridx, samples, deltas, depths, boundary = mr.ridx, mr.samples, mr.deltas, mr.depth_samples, mr.boundary
        self.prev_num_samples = samples.shape[0]
        num_samples = samples.shape[0]
        hit_ray_d = rays.dirs.index_select(0, ridx)
        color, density = nef(coords=samples, ray_d=hit_ray_d, lod_idx=lod_idx, channels=["rgb", "density"])
        density = density.reshape(num_samples, 1)
        counts = torch.bincount(ridx, minlength=N)
        offsets = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(counts, 0)
        shaded = torch.cat([color.float(), density.float()], -1)
        rgb, depth, alpha, hit = ops.CompositeFn.apply(shaded, depths, deltas, offsets, self.bg_color)
        extra_outputs = {}
        for channel in extra_channels:
            feats = nef(coords=samples, ray_d=hit_ray_d, lod_idx=lod_idx, channels=channel)
            nc = feats.shape[-1]
            outs = []
            for c0 in range(0, nc, 3):                      # integrate 3 channels at a time through the same kernel
                chunk = feats[:, c0:c0 + 3].float()
                pad = torch.zeros(num_samples, 3 - chunk.shape[1], device=dev)
                f3, _, _, _ = ops.CompositeFn.apply(torch.cat([chunk, pad, density.float()], -1), depths, deltas, offsets, (0.0, 0.0, 0.0))
                outs.append(f3[:, :chunk.shape[1]])
            extra_outputs[channel] = alpha * torch.cat(outs, -1)   # packed_rf_tracer.py:176
        return RenderBuffer(depth=depth if "depth" in channels else None, hit=hit, rgb=rgb, alpha=alpha, **extra_outputs)
