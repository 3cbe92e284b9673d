"""GPU tests (-m gpu) that restate the reference's OWN test files on the host mirror (they run last: the file name sorts after
test_gpu_parity.py).  tests/core/test_packed_rf_tracer.py::test_extra_channels asks the tracer for a channel it does not
composite itself ("density") and checks that the render buffer carries it, one row per ray."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def test_extra_channels():
    import wisp_b200 as W
    torch.manual_seed(0)
    blas = W.OctreeAS.make_dense(3, device="cuda")
    grid = W.HashGrid.from_geometric(blas, feature_dim=2, num_lods=1, multiscale_type='cat', feature_std=0.3, codebook_bitwidth=8,
                                     min_grid_res=2, max_grid_res=4)                     # init_from_geometric(2, 4, 1) in the reference test
    # bias=True: with a single LOD the reference's 'cat' quirk (hash_grid.py:226-229) zeroes the only feature level, so the biases carry the field
    nef = W.NeuralRadianceField(grid, view_embedder='positional', view_multires=4, hidden_dim=128, num_layers=1, bias=True).cuda()
    tracer = W.PackedRFTracer()                                                          # reference defaults: 'ray', 1024 steps
    pipeline = W.Pipeline(nef, tracer)
    o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], 16, 8, 30.0)                     # 128 rays (RandomViewDataset(num_rays=128))
    rays = W.Rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), dist_min=0.0, dist_max=6.0)
    rb = pipeline(rays=rays, channels=["rgb", "density"])
    assert hasattr(rb, "density")
    assert rb.rgb.shape[0] == rb.density.shape[0] == 128
    # the extra channel is alpha * (front-to-back integral of the channel), zero for rays without samples (packed_rf_tracer.py:167-179)
    miss = ~rb.hit
    assert torch.isfinite(rb.density).all() and (rb.density >= 0).all()
    if bool(miss.any()):
        assert float(rb.density[miss].abs().max()) == 0.0
    assert float(rb.density[rb.hit].sum()) > 0.0 or not bool(rb.hit.any())
    rb.density.sum().backward()                                                          # differentiable like every other channel
    gb = nef.decoder_density.lout.bias.grad
    assert gb is not None and torch.isfinite(gb).all()


@pytest.mark.gpu
@pytest.mark.parametrize("knob", ["WB_TC_FWD_TMEMA=0", "WB_TC_BWD_GROUPS=2"])
def test_kernel_variant_matches_default(knob):
    """Default kernels (TMEM-A forward, three-group decoder backward: validated and faster on B200 in round 2) against the
    round-1 kernels they replaced, which stay selectable through the knobs.  The knobs are read once per process, hence the
    subprocesses: same samples, rgb within fp16 round-off of each other, gradients within the precision-1 tolerance."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import os, sys
import numpy as np, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import wisp_b200 as W
from oracle import oracle as O
from gpu_util import nef_from_oracle, packed_grads
onef = O.make_nef(feature_std=0.2, seed=3)
spc = O.octree_to_spc(O.points_to_octree(O.lego_like_points(6), 6))
o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], 48, 48, 30.0)
nef, blas = nef_from_oracle(onef, spc)
tracer = W.PackedRFTracer('ray', 512, bg_color=(0.0, 0.0, 0.0)); tracer.seed = 9; tracer.precision = 1
rb = W.Pipeline(nef, tracer)(rays=W.Rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 0.0, 10.0), channels=["rgb"])
tgt = torch.sigmoid(torch.randn(o.shape[0], 3, generator=torch.Generator().manual_seed(2))).cuda()
torch.nn.functional.smooth_l1_loss(rb.rgb, tgt).backward()
gt, gd, gc = packed_grads(nef)
np.savez(OUT, rgb=rb.rgb.detach().cpu().numpy(), gt=gt, gd=gd, gc=gc, n=tracer.get_prev_num_samples())
'''
    name, value = knob.split("=")
    outs = []
    for on in (False, True):
        out = os.path.join(root, "gpurun_out", f"exp_{name}_{int(on)}.npz")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        env = dict(os.environ)
        if on:
            env[name] = value
        r = subprocess.run([sys.executable, "-c", code.replace("ROOT", repr(root)).replace("OUT", repr(out))], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-800:]
        outs.append(np.load(out))
    a, b = outs
    assert int(a["n"]) == int(b["n"]) > 1000
    np.testing.assert_allclose(b["rgb"], a["rgb"], atol=2e-3)
    for k in ("gt", "gd", "gc"):
        scale = np.abs(a[k]).max()
        assert np.abs(a[k] - b[k]).max() <= 3e-2 * scale, k
