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
