#!/usr/bin/env python
"""Gradient error of precision 1 (fp16 tensor-core decoders) against the fp32 CPU restatement, by decoder width and bias (GPU box only).
Prints max |error| / max |grad| per gradient group: the numbers behind the tolerance of the hidden_dim = 128 cases of
tests/test_gpu_parity.py::test_trace_other_shapes_vs_oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wisp_b200 as W
from oracle import oracle as O
from gpu_util import nef_from_oracle, packed_grads


def main():
    spc = O.octree_to_spc(O.points_to_octree(O.lego_like_points(5), 5))
    o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], 24, 24, 30.0)
    tgt = torch.sigmoid(torch.randn(o.shape[0], 3, generator=torch.Generator().manual_seed(4)))
    for hidden in (64, 128):
        for bias in (True, False):
            for seed in (11, 12):
                shape = dict(num_lods=16, feature_dim=2, codebook_bitwidth=14, min_res=16, max_res=256, hidden_dim=hidden, multiscale="cat", view_freq=4, bias=bias)
                onef = O.make_nef(feature_std=0.3, seed=seed, **shape)
                st = O.rf_step(spc, onef, o, d, 0.0, 8.0, 256, tgt.numpy(), bg=(1, 1, 1), seed=5)
                row = []
                for precision in (0, 1):
                    nef, blas = nef_from_oracle(onef, spc)
                    tracer = W.PackedRFTracer('ray', 256, bg_color=(1.0, 1.0, 1.0)); tracer.seed = 5; tracer.precision = precision
                    rb = W.Pipeline(nef, tracer)(rays=W.Rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 0.0, 8.0), channels=["rgb"])
                    torch.nn.functional.smooth_l1_loss(rb.rgb, tgt.cuda()).backward()
                    gt, gd, gc = packed_grads(nef)
                    row.append(" ".join(f"{nm}={np.abs(got - ref).max() / np.abs(ref).max():.2e}/L2 {np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()):.2e}" for got, ref, nm in ((gt, st["table"], "table"), (gd, st["dens"], "dens"), (gc, st["col"], "col"))))
                print(f"hidden={hidden} bias={bias} seed={seed} max|table grad|={np.abs(st['table']).max():.2e}  p0: {row[0]}   p1: {row[1]}", flush=True)


if __name__ == "__main__":
    main()
