"""MultiviewStep: the optimisation step of wisp.trainers.MultiviewTrainer (wisp/trainers/multiview_trainer.py:111-180) with the
optimiser set-up of BaseTrainer.init_optimizer (wisp/trainers/base_trainer.py:205-235), as ONE native sequence without autograd:

    march (count / scan / fill; pre-marched on a side stream when the next batch is known)
    shade forward (gather + decoders) -> composite forward
    composite backward WITH the image loss and its gradient inside (wb_composite_bwd_loss)       [torch: 6 launches + a [R,3] tensor]
    device loss scale -> decoder backward -> grid scatter, into persistent gradient buffers
    [N > 1: NCCL all-reduce(sum) of the gradient buffers; the 1/world is folded into the optimiser]
    Adam over grid + decoders in one launch that also clears the gradients (wb_adam_step)          [torch: zero_grad + fused Adam]

What the reference does around it and this keeps: parameter groups by name ('decoder' -> weight decay, 'grid' -> lr * grid_lr_weight),
rgb_loss_type l2 / l1 / huber, rgb_loss_denom rays / samples, tracer.prev_num_samples for the adaptive ray budget.  What it drops:
GradScaler (the fp16 decoder backward carries its own power-of-two loss scale on the device, no inf checks or skipped steps).
Fields outside the fused path (ops.nef_spec is None) fall back to autograd + the same native optimiser.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _cabi as A
from . import ops
from .core import Rays

LOSS_TYPES = {"l2": 0, "l1": 1, "huber": 2}


class NativeAdam:
    """torch.optim.Adam (amsgrad off) over a list of (tensor, lr, weight_decay) in one launch; see wb_adam_step."""

    def __init__(self, entries, betas=(0.9, 0.999), eps=1e-8):
        self.entries = [(p, float(lr), float(wd)) for p, lr, wd in entries]
        self.betas, self.eps, self.t = betas, float(eps), 0
        dev = self.entries[0][0].device
        self.exp_avg = [torch.zeros_like(p, dtype=torch.float32) for p, _, _ in self.entries]
        self.exp_avg_sq = [torch.zeros_like(p, dtype=torch.float32) for p, _, _ in self.entries]
        nb = int(A.lib().wb_adam_desc_bytes())
        self._dev = torch.empty(nb, dtype=torch.uint8, device=dev)
        self._pinned = [torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(4)]      # ring: a step's table is read by an async copy

    def step(self, grads, grad_scale: float = 1.0, zero_grad: bool = True):
        self.t += 1
        n = len(self.entries)
        segs = (A.AdamSegment * n)()
        for k, ((p, lr, wd), g) in enumerate(zip(self.entries, grads)):
            assert g.is_contiguous() and p.is_contiguous() and g.numel() == p.numel() and p.dtype == torch.float32 and g.dtype == torch.float32
            segs[k].param, segs[k].grad, segs[k].exp_avg, segs[k].exp_avg_sq = p.data_ptr(), g.data_ptr(), self.exp_avg[k].data_ptr(), self.exp_avg_sq[k].data_ptr()
            segs[k].numel, segs[k].lr, segs[k].weight_decay = p.numel(), lr, wd
        pin = self._pinned[self.t % len(self._pinned)]
        A.check(A.lib().wb_adam_step(segs, C.c_int32(n), C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps), C.c_int32(self.t),
                                     C.c_float(grad_scale), C.c_int32(int(zero_grad)), A.ptr(self._dev), C.c_void_p(pin.data_ptr()), A.stream()))


def _flatten_in_place(module_params):
    """Re-point the .data of `module_params` at consecutive views of one flat fp32 buffer (as DDP's buckets do): the packed decoder
    parameter vector the C ABI wants then exists without a per-step torch.cat, and one Adam segment covers the whole decoder."""
    flat = torch.cat([p.data.reshape(-1).float() for p in module_params]).contiguous()
    o = 0
    for p in module_params:
        n = p.numel()
        p.data = flat[o:o + n].view_as(p)
        o += n
    return flat


class MultiviewStep:
    def __init__(self, pipeline, lr: float = 1e-3, eps: float = 1e-15, weight_decay: float = 0.0, grid_lr_weight: float = 1.0, betas=(0.9, 0.999),
                 rgb_loss_type: str = "huber", rgb_loss_denom: str = "rays", precision: Optional[int] = None, group=None):
        if rgb_loss_type not in LOSS_TYPES or rgb_loss_denom not in ("rays", "samples"):
            raise NotImplementedError                                                        # multiview_trainer.py:147,157
        self.pipeline, self.nef, self.tracer = pipeline, pipeline.nef, pipeline.tracer
        self.loss_type, self.loss_denom, self.group = rgb_loss_type, rgb_loss_denom, group
        self.precision = precision
        nef = self.nef
        self.spec = ops.nef_spec(nef, None)
        self.fused = self.spec is not None
        if self.fused:
            self.dens_flat = _flatten_in_place(ops.decoder_params(nef.decoder_density))
            self.col_flat = _flatten_in_place(ops.decoder_params(nef.decoder_color))
            self.grid = ops.grid_tensors(nef, self.spec)             # every LOD: a step renders at the finest LOD (random_lod off, multiview_trainer.py:135-137)
            tensors = [(g.data, lr * grid_lr_weight, 0.0) for g in self.grid] + [(self.dens_flat, lr, weight_decay), (self.col_flat, lr, weight_decay)]
            rest = [p for n, p in nef.named_parameters() if p.requires_grad and "decoder" not in n and "grid" not in n]
            tensors += [(p.data, lr, 0.0) for p in rest]
            self.rest = rest
            self.g_grid = [torch.zeros_like(g.data, dtype=torch.float32) for g in self.grid]
            self.g_dens, self.g_col = torch.zeros_like(self.dens_flat), torch.zeros_like(self.col_flat)
            self.g_rest = [torch.zeros_like(p.data) for p in rest]
        else:
            named = list(nef.named_parameters())
            tensors = [(p.data, lr * grid_lr_weight if ("grid" in n and "decoder" not in n) else lr, weight_decay if "decoder" in n else 0.0)
                       for n, p in named if p.requires_grad]
            self.params = [p for _, p in named if p.requires_grad]
        self.opt = NativeAdam(tensors, betas=betas, eps=eps)
        dev = tensors[0][0].device
        self._scalars = torch.zeros(2, dtype=torch.float32, device=dev)       # loss accumulator | max |g_shaded|: cleared by ONE fill per step
        self.loss_buf, self.absmax = self._scalars[0:1], self._scalars[1:2]
        self.scale = torch.ones(1, dtype=torch.float32, device=dev)
        self.last_stage = {}
        self._cl = None                                                        # triplanar planes: channel-last copies + gradient accumulators

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _world(self) -> int:
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _march(self, rays: Rays, seed: int):
        nef, tr = self.nef, self.tracer
        blas = nef.grid.blas
        level = ops.raymarch_level(nef.grid, len(nef.grid.active_lods) - 1)
        if tr.raymarch_type == 'ray':
            pm = tr._pending.pop(tr._march_key(rays, seed, tr.num_steps, blas), None) if tr._pending else None
            if pm is not None:
                return pm.finalize()
            return ops.march_count(blas.tensors(), rays.origins, rays.dirs, rays.dist_min, rays.dist_max, tr.num_steps, level, seed=seed)
        ms, _ = ops.march_nuggets(blas.tensors(), rays.origins, rays.dirs, level, tr.num_steps, tr.raymarch_type, reference_layout=False, seed=seed)
        return ms

    def _precision(self) -> int:
        if self.precision is not None:
            return int(self.precision)
        if self.tracer.precision is not None:
            return int(self.tracer.precision)
        return 1 if (torch.is_autocast_enabled() and ops.precision_supported(self.spec, self.nef, 1, True)) else 0

    # ---- the step ------------------------------------------------------------------------------------------------------
    def step(self, rays: Rays, img_gts: torch.Tensor, seed: Optional[int] = None, next_rays: Optional[Rays] = None, next_seed: Optional[int] = None,
             next_ready=None, zero_grad: bool = True, local_only: bool = False, update: bool = True) -> torch.Tensor:
        """One optimisation step on (rays, img_gts [R,3]); returns the loss as a device scalar (no host sync).  `next_rays` (+ seed)
        lets the march of the following batch overlap this step (PackedRFTracer.premarch).  `update=False` stops after the backward
        (gradients left in g_grid / g_dens / g_col, parameters and optimiser state untouched: gradient checks, bench.py's parity leg)."""
        tr = self.tracer
        if seed is None:
            seed = tr.seed
            tr.seed = (tr.seed + 1) & 0x7FFFFFFF
        if next_rays is not None and tr.raymarch_type == 'ray':
            tr.premarch(self.nef, next_rays, tr.seed if next_seed is None else next_seed, ready=next_ready)
        world = 1 if local_only else self._world()      # local_only: no collective, loss normalised by this rank's rays (diagnostics)
        if not self.fused:
            return self._step_autograd(rays, img_gts, seed, world)
        L = A.lib()
        nef, spec = self.nef, self.spec
        dev = self.dens_flat.device
        ms = self._march(rays, seed)
        tr.prev_num_samples = ms.total
        S, R = ms.total, ms.rays.num_rays
        precision = self._precision()
        gt = [t.data for t in self.grid]
        g_used = self.g_grid
        layout = 1 if ops.triplane_wants_channel_last(spec) else 0
        if layout:          # channel-last copies of the planes for this step; their gradients are accumulated channel-last and converted back below
            if self._cl is None:
                self._cl = ([torch.empty((t.shape[2], t.shape[3], t.shape[1]), dtype=torch.float32, device=dev) for t in gt],
                            [torch.zeros((t.shape[2], t.shape[3], t.shape[1]), dtype=torch.float32, device=dev) for t in gt])
            gt = ops.triplane_relayout(gt, True, out=self._cl[0])
            g_used = self._cl[1]
        oct, trinkets = ops._grid_context(nef, spec)
        desc, keep = spec.desc(gt, self.dens_flat, self.col_flat, oct, trinkets, grads=g_used, layout=layout)
        blob = torch.empty(int(L.wb_rf_param_blob_floats(C.byref(desc), C.c_int32(precision))), dtype=torch.float32, device=dev)
        A.check(L.wb_rf_pack_params(C.byref(desc), C.c_int32(precision), A.ptr(blob), A.stream()))
        rec_t, rec_delta, rec_ray = ops.march_fill_records(ms, dev)
        Scap = ops._bucket(S)
        shaded = ops._empty_s(S, (4,), torch.float32, dev)
        g_sh = ops._empty_s(S, (4,), torch.float32, dev)
        wsb = int(L.wb_rf_workspace_bytes(C.byref(desc), C.c_int32(precision), C.c_int64(R), C.c_int64(Scap), C.c_int32(1)))
        fb = int(L.wb_rf_feat_bytes(C.byref(desc), C.c_int32(precision), C.c_int64(Scap)))
        if wsb < 0 or fb < 0:
            raise A.WispB200Error(L.wb_last_error().decode())
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb > 0 else None
        feat = torch.empty(fb, dtype=torch.uint8, device=dev) if fb > 0 else None
        rgb = torch.empty((R, 3), dtype=torch.float32, device=dev)
        alpha = torch.empty((R, 1), dtype=torch.float32, device=dev)
        hit = torch.empty(R, dtype=torch.bool, device=dev)
        tr.bg_color = tr.bg_color.to(dev)
        bgv = ops._bg3(tr.bg_color)
        tgt = A.f32c(img_gts)
        with ops._stage("shade_fwd"):
            A.check(L.wb_rf_shade_fwd(C.byref(desc), A.ptr(blob), C.c_int32(precision), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray), C.c_int64(S),
                                      A.ptr(shaded), A.ptr(feat), A.ptr(ws), A.stream()))
        with ops._stage("composite_fwd"):
            A.check(L.wb_composite_fwd(A.ptr(shaded), A.ptr(rec_t), A.ptr(rec_delta), A.ptr(ms.offsets), C.c_int64(R), bgv, A.ptr(rgb), None, A.ptr(alpha),
                                       A.ptr(hit), A.stream()))
        self._scalars.zero_()
        # rgb_loss.mean() over the GLOBAL batch: every rank contributes sum / (3 * R * world); 'samples' divides by the local count
        inv = 1.0 / (3.0 * R * world) if self.loss_denom == "rays" else 1.0 / (max(S, 1) * world)
        with ops._stage("composite_bwd"):
            A.check(L.wb_composite_bwd_loss(A.ptr(shaded), A.ptr(rec_t), A.ptr(rec_delta), A.ptr(ms.offsets), C.c_int64(R), bgv, A.ptr(rgb), A.ptr(tgt),
                                            C.c_int32(LOSS_TYPES[self.loss_type]), C.c_float(inv), A.ptr(g_sh), A.ptr(self.absmax), A.ptr(self.loss_buf), A.stream()))
        g_table = g_used[0] if spec.kind == "hash" else None
        if S > 0:
            if precision == 1:
                A.check(L.wb_rf_loss_scale(A.ptr(self.absmax), A.ptr(self.scale), A.stream()))
                L.wb_rf_workspace_holds_ray_rows(C.c_int32(1))      # same workspace, same rays as the forward above
            with ops._stage("shade_bwd"):      # precision 1: decoder backward + table scatter in one kernel where the shape allows
                A.check(L.wb_rf_shade_bwd(C.byref(desc), A.ptr(blob), C.c_int32(precision), C.byref(ms.rays), A.ptr(rec_t), A.ptr(rec_ray), C.c_int64(S), A.ptr(g_sh),
                                          A.ptr(self.scale) if precision == 1 else None, A.ptr(feat), A.ptr(ws), A.ptr(g_table), A.ptr(self.g_dens), A.ptr(self.g_col), A.stream()))
        if layout:          # -> self.g_grid (reference layout, overwritten with the accumulated channel-last gradients)
            ops.triplane_relayout(g_used, False, out=self.g_grid)
        self.last_rgb, self.last_alpha, self.last_hit = rgb, alpha, hit
        loss = self.loss_buf.clone()
        if world > 1:
            with ops._stage("all_reduce"):
                self._all_reduce(loss)
        if update:
            with ops._stage("adam"):
                self.opt.step(self.g_grid + [self.g_dens, self.g_col] + self.g_rest, grad_scale=1.0, zero_grad=zero_grad)
            if layout and zero_grad:
                torch._foreach_zero_(g_used)
        return loss[0]

    def zero_grads(self) -> None:
        """Clear every gradient accumulator (after a step(update=False) whose gradients were only inspected)."""
        for g in self.g_grid + [self.g_dens, self.g_col] + self.g_rest + (self._cl[1] if self._cl is not None else []):
            g.zero_()

    def _all_reduce(self, loss):
        """Gradient exchange of data-parallel training: sum over ranks (the mean's 1/world is already in the loss
        gradient: inv_count uses the global ray count).  Decoder gradients + the loss share one small collective."""
        small = torch.cat([self.g_dens, self.g_col, loss] + [g.reshape(-1) for g in self.g_rest])
        dist.all_reduce(small, op=dist.ReduceOp.SUM, group=self.group)
        for g in self.g_grid:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        o = 0
        for t in [self.g_dens, self.g_col, loss] + [g.reshape(-1) for g in self.g_rest]:
            t.copy_(small[o:o + t.numel()].view_as(t)); o += t.numel()

    def _step_autograd(self, rays, img_gts, seed, world):
        for p in self.params:
            p.grad = None
        self.tracer.seed = seed
        rb = self.pipeline(rays=rays, lod_idx=None, channels=["rgb"])
        d = rb.rgb - img_gts
        if self.loss_type == "l2":
            l = torch.nn.functional.mse_loss(rb.rgb, img_gts, reduction='none')
        elif self.loss_type == "l1":
            l = torch.abs(d)
        else:
            l = torch.nn.functional.smooth_l1_loss(rb.rgb, img_gts, reduction='none')
        loss = l.mean() if self.loss_denom == "rays" else l.sum() / max(self.tracer.prev_num_samples, 1)
        loss.backward()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        grads = [g.contiguous() for g in grads]
        if world > 1:
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step(grads, grad_scale=1.0 / world, zero_grad=False)
        return loss.detach()
