"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden

What is pinned:
  * golden/hashgrid_naive.npz  -- wisp.ops.grid.hashgrid_naive (ops/grid.py:16-75), the reference's own
    pure-torch statement of the hash index + trilinear blend.  Depends on NO oracle code (only a 3-line
    points_to_corners stub), so it pins oracle/wisp_oracle.c:wo_hashgrid_* independently.
  * golden/rf_trace_*.npz      -- wisp.models.Pipeline(NeuralRadianceField(HashGrid), PackedRFTracer)
    forward + backward, executed by the reference's own classes on CPU.  Kaolin / wisp._C native calls are
    answered by the oracle (oracle/ref_import.py), torch.rand is replaced by a recorded jitter tensor.
    Pins every piece of wisp-side glue: sample generation and culling (octree_as.py:247-309), the 'cat'
    zeroing / 'sum' reduction (hash_grid.py:224-233), decoder + embedder wiring (nerf.py:219-264),
    tau/integration/background blend and buffer scatter (packed_rf_tracer.py:130-165).
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from . import oracle as O
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _mlp_params(m):
    Ws = [l.weight.detach().numpy().copy() for l in m.layers] + [m.lout.weight.detach().numpy().copy()]
    if m.lout.bias is None:
        return Ws, None
    bs = [l.bias.detach().numpy().copy() for l in m.layers] + [m.lout.bias.detach().numpy().copy()]
    return Ws, bs


def _mlp_grads(m):
    parts = []
    for l in list(m.layers) + [m.lout]:
        parts.append(l.weight.grad.reshape(-1).numpy())
        if l.bias is not None:
            parts.append(l.bias.grad.reshape(-1).numpy())
    return np.concatenate(parts)


def gen_hashgrid_naive():
    from wisp.ops.grid import hashgrid_naive
    rng = np.random.default_rng(7)
    resolutions = [4, 7, 13, 24, 40]
    bw = 9                       # T=512: levels 4,7 dense (res^3<T), others hashed; no res^3 == T case
    T = 2 ** bw
    begin = O.table_layout(resolutions, bw)
    sizes = np.diff(begin)
    F = 2
    table = rng.standard_normal((int(begin[-1]), F)).astype(np.float32)
    coords = rng.uniform(-0.999, 0.999, (257, 3)).astype(np.float32)
    feats = hashgrid_naive(torch.from_numpy(coords), torch.tensor(resolutions), bw, len(resolutions) - 1,
                           torch.from_numpy(table), torch.from_numpy(sizes), torch.from_numpy(begin[:-1]))
    np.savez_compressed(os.path.join(OUT, "hashgrid_naive.npz"), coords=coords, table=table, resolutions=np.asarray(resolutions),
                        codebook_bitwidth=bw, feats=feats.numpy())
    print("hashgrid_naive", feats.shape, "T", T)


def gen_rf_trace(name, *, level, res, hw, n_steps, num_lods, bw, min_res, max_res, hidden, num_layers, bias, multiscale,
                 view_embedder, near, far, bg, feature_std=0.5, sparse=True, loss="huber", seed=0):
    from wisp.models import Pipeline
    from wisp.tracers import PackedRFTracer
    from wisp.accelstructs import OctreeAS
    from wisp.models.grids import HashGrid
    from wisp.models.nefs import NeuralRadianceField
    from wisp.core import Rays
    torch.manual_seed(seed)
    oct_np = O.points_to_octree(O.lego_like_points(level), level) if sparse else O.dense_octree(level)
    blas = OctreeAS(torch.from_numpy(oct_np))
    grid = HashGrid.from_geometric(blas, feature_dim=2, num_lods=num_lods, multiscale_type=multiscale, feature_std=feature_std,
                                   codebook_bitwidth=bw, min_grid_res=min_res, max_grid_res=max_res)
    nef = NeuralRadianceField(grid, view_embedder=view_embedder, view_multires=4, hidden_dim=hidden, num_layers=num_layers, bias=bias)
    tracer = PackedRFTracer(raymarch_type='ray', num_steps=n_steps, bg_color=bg)
    pipe = Pipeline(nef, tracer)
    o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], hw, hw, 30.0)
    R = o.shape[0]
    jit = np.random.default_rng(seed + 1).random((R, n_steps), dtype=np.float32)
    rays = Rays(torch.from_numpy(o), torch.from_numpy(d), dist_min=near, dist_max=far)
    orig = torch.rand
    torch.rand = lambda *a, **k: torch.from_numpy(jit)
    try:
        mr = grid.raymarch(rays, level=grid.active_lods[-1], num_samples=n_steps, raymarch_type='ray')
        rb = pipe(rays=rays, channels=["rgb", "depth", "alpha", "hit"])
    finally:
        torch.rand = orig
    target = torch.sigmoid(torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((R, 3)).astype(np.float32)))
    lossv = {"huber": torch.nn.functional.smooth_l1_loss(rb.rgb, target, reduction='none').mean(),
             "l2": torch.nn.functional.mse_loss(rb.rgb, target, reduction='none').mean(),
             "l1": torch.abs(rb.rgb - target).mean()}[loss]
    lossv.backward()
    dW, db = _mlp_params(nef.decoder_density)
    cW, cb = _mlp_params(nef.decoder_color)
    vm = {"positional": 3, "none": 1, "identity": 1}[view_embedder]   # nerf.py:105-106,116-119: include_input=True makes "none" an identity
    onef = O.Nef([int(r) for r in grid.resolutions], 2, bw, grid.codebook.feats.detach().numpy(), dW, db, cW, cb,
                 multiscale=multiscale, view_mode=vm, view_freq=4)
    icfg, resa, begin, table, pd, pc = onef.pack()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        octree=oct_np, level=level, origins=o, dirs=d, near=near, far=far, n_steps=n_steps, jitter=jit, bg=np.asarray(bg, np.float32),
        icfg=icfg, res=resa, begin=begin, table=table, dens_params=pd, col_params=pc, target=target.numpy(), loss_type=loss,
        # reference outputs
        mr_ridx=mr.ridx.numpy(), mr_samples=mr.samples.numpy(), mr_depth=mr.depth_samples.numpy(), mr_deltas=mr.deltas.numpy(),
        mr_boundary=mr.boundary.numpy(),
        rgb=rb.rgb.detach().numpy(), depth=rb.depth.detach().numpy(), alpha=rb.alpha.detach().numpy(), hit=rb.hit.numpy(),
        num_samples=tracer.get_prev_num_samples(), loss=float(lossv),
        g_table=grid.codebook.feats.grad.numpy(), g_dens=_mlp_grads(nef.decoder_density), g_col=_mlp_grads(nef.decoder_color))
    print(name, "R", R, "S", tracer.get_prev_num_samples(), "hit", int(rb.hit.sum()), "loss", float(lossv))


def gen_raymarch_nuggets():
    """OctreeAS._raymarch_voxel / _raymarch_uniform executed by the reference Python (octree_as.py:188-245, 311-374):
    pins sample_from_depth_intervals, expand_pack_boundary, the lattice scale, zero-count filtering and the deltas."""
    from wisp.accelstructs import OctreeAS
    from wisp.core import Rays
    level, n_v, n_u = 5, 6, 96
    oct_np = O.points_to_octree(O.lego_like_points(level), level)
    blas = OctreeAS(torch.from_numpy(oct_np))
    o, d = O.look_at_rays([-3.0, 0.65, -3.0], [0, 0, 0], 20, 20, 30.0)
    rays = Rays(torch.from_numpy(o), torch.from_numpy(d), dist_min=0.0, dist_max=10.0)
    rt = blas.raytrace(rays, level, with_exit=True)
    Ng = rt.ridx.shape[0]
    jit = np.random.default_rng(11).random((Ng, n_v), dtype=np.float32)
    orig = torch.rand_like
    torch.rand_like = lambda t, **k: torch.from_numpy(jit)
    try:
        mv = blas.raymarch(rays, 'voxel', n_v, level)
    finally:
        torch.rand_like = orig
    mu = blas.raymarch(rays, 'uniform', n_u, level)
    np.savez_compressed(os.path.join(OUT, "raymarch_nuggets.npz"), octree=oct_np, level=level, origins=o, dirs=d, n_voxel=n_v, n_uniform=n_u,
                        jitter=jit, nug_ridx=rt.ridx.numpy(), nug_pidx=rt.pidx.numpy(), nug_depth=rt.depth.numpy(),
                        v_ridx=mv.ridx.numpy(), v_samples=mv.samples.numpy(), v_depth=mv.depth_samples.numpy(), v_deltas=mv.deltas.numpy(),
                        v_boundary=mv.boundary.numpy(),
                        u_ridx=mu.ridx.numpy(), u_samples=mu.samples.numpy(), u_depth=mu.depth_samples.numpy(), u_deltas=mu.deltas.numpy(),
                        u_boundary=mu.boundary.numpy())
    print("raymarch_nuggets", Ng, mv.ridx.shape, mu.ridx.shape)


def gen_triplanar():
    """TriplanarGrid.interpolate (triplanar_grid.py:98-143, 205-223) forward + gradients, executed by the reference class
    (three F.grid_sample calls per LOD on CPU).  Coordinates reach outside [-1,1] to exercise the reflection padding."""
    from wisp.models.grids.triplanar_grid import TriplanarGrid
    torch.manual_seed(4)
    out = {}
    for ms in ("sum", "cat"):
        grid = TriplanarGrid(None, feature_dim=4, log_base_resolution=2, num_lods=3, multiscale_type=ms, feature_std=1.0)
        coords = torch.rand(301, 3) * 2.8 - 1.4
        coords[:8] = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [1.0, -1.0, 0.5], [0.25, 0.5, -0.75],
                                   [-1.4, 1.4, 0.0], [2.5, -2.5, 3.1], [0.999999, -0.999999, 1e-7]])
        feats = grid.interpolate(coords, 2)
        go = torch.randn_like(feats)
        feats.backward(go)
        planes = [getattr(f, n) for f in grid.features for n in ("fmx", "fmy", "fmz")]
        out[ms] = dict(coords=coords.numpy(), feats=feats.detach().numpy(), go=go.numpy(),
                       **{f"plane{i}": p.detach().numpy() for i, p in enumerate(planes)}, **{f"gplane{i}": p.grad.numpy() for i, p in enumerate(planes)})
        # a lower lod_idx uses only the first LODs
        out[ms]["feats_lod0"] = grid.interpolate(coords, 0).detach().numpy()
    np.savez_compressed(os.path.join(OUT, "triplanar.npz"), **{f"{ms}_{k}": v for ms, d in out.items() for k, v in d.items()})
    print("triplanar", out["sum"]["feats"].shape, out["cat"]["feats"].shape)


def octahedron_points(level: int, radius: float = 0.52, n: int = 400000, seed: int = 0) -> np.ndarray:
    """Quantised points on the surface |x|+|y|+|z| = radius (the zero set of the synthetic SDF below)."""
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n, 3))
    p = p / np.abs(p).sum(-1, keepdims=True) * radius
    q = np.floor(np.clip((2 ** level) * (p + 1.0) / 2.0, 0, 2 ** level - 1)).astype(np.int16)
    return np.unique(q, axis=0)


def gen_sdf():
    """app/nglod path in miniature (BASELINE config 3): OctreeGrid.interpolate (octree_grid.py:165-219), NeuralSDF.sdf
    (neural_sdf.py:120-155) and PackedSDFTracer.trace (packed_sdf_tracer.py:57-174) run by the reference classes."""
    from wisp.accelstructs import OctreeAS
    from wisp.models.grids import OctreeGrid
    from wisp.models.nefs import NeuralSDF
    from wisp.tracers import PackedSDFTracer
    from wisp.core import Rays
    torch.manual_seed(7)
    level = 5
    oct_np = O.points_to_octree(octahedron_points(level), level)
    blas = OctreeAS(torch.from_numpy(oct_np))
    out = dict(octree=oct_np, level=level)
    for ms in ("sum", "cat"):
        grid = OctreeGrid(blas, feature_dim=8, num_lods=3, interpolation_type='linear', multiscale_type=ms, feature_std=0.05)
        nef = NeuralSDF(grid, pos_embedder='none', position_input=True, hidden_dim=16, num_layers=1)
        with torch.no_grad():       # sdf ~ (|x|+|y|+|z|)/sqrt(3) - 0.3 + small learned perturbation
            W0 = nef.decoder.layers[0].weight; W0.mul_(0.05)
            W0[:6, :3] = torch.tensor([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1.0]])
            nef.decoder.layers[0].bias.zero_()
            nef.decoder.lout.weight.mul_(0.05); nef.decoder.lout.weight[0, :6] = 1.0 / np.sqrt(3.0)
            nef.decoder.lout.bias.fill_(-0.3)
        coords = torch.rand(500, 3) * 1.4 - 0.7
        feats = grid.interpolate(coords, 2)
        feats0 = grid.interpolate(coords, 0)
        sdf = nef(coords=coords, lod_idx=2, channels="sdf")
        (sdf.abs().sum()).backward()
        d = {f"{ms}_coords": coords.numpy(), f"{ms}_feats": feats.detach().numpy(), f"{ms}_feats_lod0": feats0.detach().numpy(), f"{ms}_sdf": sdf.detach().numpy()}
        for i, f in enumerate(grid.features):
            d[f"{ms}_feat{i}"] = f.detach().numpy(); d[f"{ms}_gfeat{i}"] = f.grad.numpy()
        d[f"{ms}_W0"] = nef.decoder.layers[0].weight.detach().numpy(); d[f"{ms}_b0"] = nef.decoder.layers[0].bias.detach().numpy()
        d[f"{ms}_W1"] = nef.decoder.lout.weight.detach().numpy(); d[f"{ms}_b1"] = nef.decoder.lout.bias.detach().numpy()
        d[f"{ms}_gW0"] = nef.decoder.layers[0].weight.grad.numpy()
        if ms == "sum":
            o, dd = O.look_at_rays([-2.0, 0.9, -1.6], [0, 0, 0], 20, 20, 40.0)
            tracer = PackedSDFTracer(num_steps=24, step_size=0.8, min_dis=1e-3)
            rb = tracer(nef, rays=Rays(torch.from_numpy(o), torch.from_numpy(dd), dist_min=0.0, dist_max=6.0), lod_idx=2,
                        channels=["rgb", "depth", "hit", "normal", "alpha", "xyz"])
            d.update(origins=o, dirs=dd, t_xyz=rb.xyz.detach().numpy(), t_depth=rb.depth.detach().numpy(), t_hit=rb.hit.numpy(),
                     t_normal=rb.normal.detach().numpy(), t_rgb=rb.rgb.detach().numpy(), t_alpha=rb.alpha.detach().numpy())
            print("sdf trace hits", int(rb.hit.sum()), "of", o.shape[0])
        out.update(d)
        out[f"{ms}_trinkets"] = grid.trinkets.numpy(); out[f"{ms}_pyramid_dual"] = grid.pyramid_dual.numpy()
    np.savez_compressed(os.path.join(OUT, "sdf_octree.npz"), **out)
    print("sdf_octree", out["sum_feats"].shape, out["cat_feats"].shape)


def gen_prune():
    """NeuralRadianceField.prune (nerf.py:175-212) executed by the reference class on CPU: occupancy decay, jittered density probe of
    every finest-level cell through the reference's HashGrid + decoders, threshold, octree rebuild (from_quantized_points).
    torch.rand is replaced by a recorded draw; `.cuda()` is the identity (no GPU in the build container)."""
    from wisp.accelstructs import OctreeAS
    from wisp.models.grids import HashGrid
    from wisp.models.nefs import NeuralRadianceField
    torch.manual_seed(5)
    level = 4
    oct_np = O.dense_octree(level)
    blas = OctreeAS(torch.from_numpy(oct_np))
    grid = HashGrid.from_geometric(blas, feature_dim=2, num_lods=6, multiscale_type='cat', feature_std=0.8, codebook_bitwidth=11,
                                   min_grid_res=4, max_grid_res=48)
    nef = NeuralRadianceField(grid, view_embedder='positional', view_multires=4, hidden_dim=32, num_layers=1, bias=True,
                              prune_density_decay=0.6, prune_min_density=0.0)
    N = grid.dense_points.shape[0]
    rng = np.random.default_rng(21)
    u = rng.random((N, 3), dtype=np.float32)
    occ0 = (rng.random(N, dtype=np.float32) * 1.5).astype(np.float32)
    grid.occupancy = torch.from_numpy(occ0.copy())
    # threshold at the median of what the update will produce, so that about half of the cells survive
    orig_rand, orig_cuda = torch.rand, torch.Tensor.cuda
    torch.rand = lambda *a, **k: torch.from_numpy(u)
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        with torch.no_grad():
            samples = (grid.dense_points.float() + torch.from_numpy(u)) / (2.0 ** level) * 2.0 - 1.0
            dens = nef(coords=samples, ray_d=torch.zeros_like(samples) + 0.5, channels="density")[:, 0]
        nef.prune_min_density = float(np.median(np.maximum(dens.numpy(), occ0 * 0.6)))
        dW, db = _mlp_params(nef.decoder_density)
        cW, cb = _mlp_params(nef.decoder_color)
        onef = O.Nef([int(r) for r in grid.resolutions], 2, 11, grid.codebook.feats.detach().numpy().copy(), dW, db, cW, cb, multiscale="cat", view_mode=3, view_freq=4)
        icfg, resa, begin, table, pd, pc = onef.pack()
        nef.prune()
    finally:
        torch.rand, torch.Tensor.cuda = orig_rand, orig_cuda
    new_occ = grid.occupancy.numpy()
    keep = new_occ > nef.prune_min_density
    np.savez_compressed(os.path.join(OUT, "prune.npz"), octree=oct_np, level=level, u=u, occupancy0=occ0, decay=0.6, min_density=nef.prune_min_density,
                        icfg=icfg, res=resa, begin=begin, table=table, dens_params=pd, col_params=pc,
                        density=dens.numpy(), occupancy1=new_occ, keep=keep, new_octree=grid.blas.octree.numpy(),
                        new_max_level=grid.blas.max_level)
    print("prune: cells", N, "kept", int(keep.sum()), "new octree bytes", grid.blas.octree.shape[0])


def gen_raygen():
    """_look_at / _generate_rays (wisp/trainers/tracker/offline_renderer.py:23-89) + normalized_grid (wisp/ops/geometric.py:65-99),
    executed from the UNMODIFIED reference source on CPU (the two functions are compiled from the file; the module itself cannot be
    imported because wisp.trainers pulls the GUI stack)."""
    import ast
    import torch.nn.functional as F
    from wisp.ops.geometric import normalized_grid
    path = ref_import.REF_ROOT + "/wisp/trainers/tracker/offline_renderer.py"
    tree = ast.parse(open(path).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_look_at", "_generate_rays")]
    ns = dict(torch=torch, F=F, np=np, normalized_grid=normalized_grid)
    exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
    out = {}
    for name, (f, t, h, w, mode, fov) in dict(square=([-3.0, 0.65, -3.0], [0, 0, 0], 48, 48, 'persp', 30.0),
                                              wide=([2.0, 1.5, -1.0], [0.1, -0.2, 0.3], 20, 36, 'persp', 55.0),
                                              tall=([0.5, 2.5, 3.0], [0, 0, 0], 31, 17, 'persp', 90.0),
                                              ortho=([-1.0, 1.0, 2.0], [0, 0, 0], 16, 24, 'ortho', 40.0)).items():
        o, d = ns["_look_at"](f, t, h, w, mode=mode, fov=fov, device='cpu')
        out[name + "_args"] = np.asarray(f + t + [h, w, fov], np.float64)
        out[name + "_mode"] = mode
        out[name + "_origins"] = o.numpy(); out[name + "_dirs"] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "raygen.npz"), **out)
    print("raygen", {k: v.shape for k, v in out.items() if k.endswith("_dirs")})


def gen_codebook():
    """CodebookOctreeGrid.interpolate (codebook_grid.py:103-172 over octree_grid.py:165-219) run by the reference class on CPU, in
    training mode (straight-through softmax selection, gradients to logits and dictionary) and in eval mode (argmax selection)."""
    from wisp.accelstructs import OctreeAS
    from wisp.models.grids import CodebookOctreeGrid
    torch.manual_seed(9)
    level = 4
    oct_np = O.points_to_octree(octahedron_points(level, radius=0.6, n=60000, seed=3), level)
    blas = OctreeAS(torch.from_numpy(oct_np))
    out = dict(octree=oct_np, level=level)
    for ms in ("sum", "cat"):
        grid = CodebookOctreeGrid(blas, feature_dim=4, num_lods=3, interpolation_type='linear', multiscale_type=ms, feature_std=1.0, codebook_bitwidth=4)
        pts = blas.points[blas.pyramid[1, level]: blas.pyramid[1, level] + blas.pyramid[0, level]].float()
        sel = torch.randint(0, pts.shape[0], (300,))
        coords = (((pts[sel] + torch.rand(300, 3)) / 2 ** level) * 2 - 1).float()
        coords = torch.cat([coords, torch.rand(60, 3) * 2 - 1])                       # some points outside the occupied cells
        grid.train()
        feats = grid.interpolate(coords, 2)
        go = torch.randn_like(feats)
        feats.backward(go)
        d = {f"{ms}_coords": coords.numpy(), f"{ms}_feats_train": feats.detach().numpy(), f"{ms}_go": go.numpy()}
        for i in range(3):
            d[f"{ms}_logits{i}"] = grid.features[i].detach().numpy().copy(); d[f"{ms}_dict{i}"] = grid.dictionary[i].detach().numpy().copy()
            d[f"{ms}_glogits{i}"] = grid.features[i].grad.numpy().copy(); d[f"{ms}_gdict{i}"] = grid.dictionary[i].grad.numpy().copy()
        grid.eval()
        with torch.no_grad():
            d[f"{ms}_feats_eval"] = grid.interpolate(coords, 2).numpy()
            d[f"{ms}_feats_eval_lod0"] = grid.interpolate(coords, 0).numpy()
        out.update(d)
    np.savez_compressed(os.path.join(OUT, "codebook.npz"), **out)
    print("codebook", out["sum_feats_train"].shape, out["cat_feats_train"].shape, float(np.abs(out["sum_glogits2"]).max()))


def main():
    warnings.filterwarnings("ignore")
    ref_import.install()
    os.makedirs(OUT, exist_ok=True)
    gen_hashgrid_naive()
    gen_raymarch_nuggets()
    gen_triplanar()
    gen_sdf()
    gen_prune()
    gen_raygen()
    gen_codebook()
    # A: miniature of BASELINE config 2 (cat, bias, positional view embedding, sparse lego-like octree)
    gen_rf_trace("rf_trace_cat", level=5, res=None, hw=24, n_steps=96, num_lods=6, bw=11, min_res=4, max_res=48, hidden=32,
                 num_layers=1, bias=True, multiscale="cat", view_embedder="positional", near=0.0, far=10.0, bg=(1.0, 1.0, 1.0))
    # B: 'sum' aggregation, no bias, 2 hidden layers, dense octree, black background, l2 loss, training-style near/far
    gen_rf_trace("rf_trace_sum", level=3, res=None, hw=16, n_steps=48, num_lods=4, bw=10, min_res=4, max_res=32, hidden=16,
                 num_layers=2, bias=False, multiscale="sum", view_embedder="positional", near=1.0, far=6.0, bg=(0.0, 0.0, 0.0),
                 sparse=False, loss="l2", seed=3)
    # C: view_embedder="none" (which nerf.py:105-106,116-119 turns into an identity embedding of ray_d), l1 loss
    gen_rf_trace("rf_trace_noview", level=4, res=None, hw=16, n_steps=64, num_lods=4, bw=10, min_res=4, max_res=32, hidden=16,
                 num_layers=1, bias=True, multiscale="cat", view_embedder="none", near=0.0, far=10.0, bg=(0.2, 0.5, 0.9),
                 loss="l1", seed=5)


if __name__ == "__main__":
    main()
