#!/usr/bin/env python
"""How far is fp16-tensor-core training (precision 1) from fp32, compared with torch's own AMP?  (GPU box only.)

For a TriplanarGrid NeRF ('sum' and 'cat') and the app/nerf HashGrid: gradients of one step computed by
  (a) the unfused route in fp32 (native grid kernel + torch nn.Linear): the comparison baseline,
  (b) the same route under torch.autocast(fp16) -- what the reference runs with enable_amp: True,
  (c) the fused native path at precision 1.
Prints max |error| / max |grad| per parameter for (b) and (c): the tolerance of tests/test_gpu_parity.py::test_fused_triplanar_octree_nerf
is justified if (c) is no worse than the reference's own AMP numerics (b)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wisp_b200 as W
from oracle import oracle as O


def grads(nef, tracer, rays, fused, precision, amp):
    for p in nef.parameters():
        p.grad = None
    tracer.precision = precision
    tracer.seed = 11
    if not fused:
        nef.fused_spec = lambda lod_idx=None: None
    try:
        with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
            rb = tracer(nef, rays=rays, channels=["rgb", "depth", "alpha", "hit"])
        tgt = torch.sigmoid(torch.randn(rays.origins.shape[0], 3, generator=torch.Generator().manual_seed(4))).cuda()
        torch.nn.functional.smooth_l1_loss(rb.rgb.float(), tgt).backward()
    finally:
        if not fused:
            del nef.fused_spec
    return {n: p.grad.detach().float().clone() for n, p in nef.named_parameters() if p.grad is not None}


def main():
    o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], 40, 40, 30.0)
    rays = W.Rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 0.0, 10.0)
    for kind in ("triplanar_sum", "triplanar_cat", "octree_sum", "hash_cat"):
        torch.manual_seed(2)
        ms = kind.split("_")[1]
        if kind.startswith("triplanar"):
            grid = W.TriplanarGrid(W.AxisAlignedBBoxAS(device="cuda"), feature_dim=4, log_base_resolution=6, num_lods=4, multiscale_type=ms, feature_std=0.3)
            tracer = W.PackedRFTracer('voxel', 48, bg_color=(1.0, 1.0, 1.0))
        elif kind.startswith("octree"):
            blas = W.OctreeAS.from_quantized_points(torch.from_numpy(O.lego_like_points(6)).cuda(), 6)
            grid = W.OctreeGrid(blas, feature_dim=8, num_lods=4, multiscale_type=ms, feature_std=0.3)
            tracer = W.PackedRFTracer('ray', 192, bg_color=(1.0, 1.0, 1.0))
        else:
            blas = W.OctreeAS.from_quantized_points(torch.from_numpy(O.lego_like_points(6)).cuda(), 6)
            grid = W.HashGrid.from_geometric(blas, feature_dim=2, num_lods=16, multiscale_type='cat', feature_std=0.3, codebook_bitwidth=16, min_grid_res=16, max_grid_res=256)
            tracer = W.PackedRFTracer('ray', 192, bg_color=(1.0, 1.0, 1.0))
        nef = W.NeuralRadianceField(grid, view_embedder='positional', view_multires=4, hidden_dim=64, num_layers=1, bias=True).cuda()
        ref = grads(nef, tracer, rays, fused=False, precision=0, amp=False)
        amp = grads(nef, tracer, rays, fused=False, precision=0, amp=True)
        p1 = grads(nef, tracer, rays, fused=True, precision=1, amp=False)
        worst = {"amp": 0.0, "p1": 0.0}
        for n, g in ref.items():
            sc = float(g.abs().max())
            for tag, other in (("amp", amp), ("p1", p1)):
                worst[tag] = max(worst[tag], float((other[n] - g).abs().max()) / max(sc, 1e-30))
        print(f"{kind:15s} worst max|err|/max|grad|:  torch-AMP unfused {worst['amp']:.4f}   native precision 1 {worst['p1']:.4f}", flush=True)


if __name__ == "__main__":
    main()
