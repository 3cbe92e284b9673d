"""Build wisp_b200 objects (the product) from oracle-side descriptions, for the GPU parity tests."""
import numpy as np
import torch

import wisp_b200 as W


def nef_from_oracle(onef, spc_np, device="cuda"):
    """oracle.Nef + oracle.SPC -> (W.NeuralRadianceField, W.OctreeAS) with identical parameters."""
    blas = W.OctreeAS(torch.from_numpy(spc_np.octree).to(device))
    grid = W.HashGrid(blas, onef.feature_dim, onef.resolutions, multiscale_type=onef.multiscale, feature_std=0.0,
                      codebook_bitwidth=onef.codebook_bitwidth)
    view = {0: "none", 1: "identity", 3: "positional"}[onef.view_mode]
    pos = {0: "none", 1: "identity", 2: "positional", 3: "positional"}[onef.pos_mode]
    hidden = onef.dens_W[0].shape[0]
    nef = W.NeuralRadianceField(grid, pos_embedder=pos, view_embedder=view, pos_multires=max(onef.pos_freq, 1), view_multires=max(onef.view_freq, 1),
                                position_input=onef.pos_mode in (1, 3), hidden_dim=hidden, num_layers=len(onef.dens_W) - 1,
                                bias=onef.dens_b is not None)
    with torch.no_grad():
        grid.codebook.feats.copy_(torch.from_numpy(onef.table))
        for dec, Ws, bs in ((nef.decoder_density, onef.dens_W, onef.dens_b), (nef.decoder_color, onef.col_W, onef.col_b)):
            lin = list(dec.layers) + [dec.lout]
            for i, l in enumerate(lin):
                l.weight.copy_(torch.from_numpy(np.ascontiguousarray(Ws[i])))
                if bs is not None:
                    l.bias.copy_(torch.from_numpy(np.ascontiguousarray(bs[i])))
    return nef.to(device), blas


def packed_grads(nef):
    def flat(dec):
        parts = []
        for l in list(dec.layers) + [dec.lout]:
            parts.append(l.weight.grad.reshape(-1))
            if l.bias is not None:
                parts.append(l.bias.grad.reshape(-1))
        return torch.cat(parts).cpu().numpy()
    return nef.grid.codebook.feats.grad.cpu().numpy(), flat(nef.decoder_density), flat(nef.decoder_color)


def sdf_nef_from_case(case, device="cuda"):
    """oracle.octree_grid.make_sdf_case -> W.NeuralSDF(OctreeGrid) with identical octree, features and decoder."""
    blas = W.OctreeAS(torch.from_numpy(case["octree"]).to(device))
    grid = W.OctreeGrid(blas, feature_dim=case["feature_dim"], num_lods=len(case["active_lods"]), multiscale_type=case["multiscale"], feature_std=0.0)
    assert grid.active_lods == list(case["active_lods"])
    assert np.array_equal(grid.trinkets.cpu().numpy(), case["trinkets"])
    nef = W.NeuralSDF(grid, pos_embedder='none', position_input=True, hidden_dim=case["hidden_dim"], num_layers=len(case["W"]) - 1).to(device)
    with torch.no_grad():
        for f, ref in zip(grid.features, case["feats"]):
            f.copy_(torch.from_numpy(ref))
        lin = list(nef.decoder.layers) + [nef.decoder.lout]
        for l, Wm, b in zip(lin, case["W"], case["b"]):
            l.weight.copy_(torch.from_numpy(Wm)); l.bias.copy_(torch.from_numpy(b))
    return nef
