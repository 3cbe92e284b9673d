"""Helpers to rebuild oracle inputs from a tests/golden/rf_trace_*.npz fixture."""
import numpy as np

from oracle import oracle as O


def load_case(path):
    g = dict(np.load(path, allow_pickle=False))
    icfg = g["icfg"]
    L, F, T = int(icfg[0]), int(icfg[1]), int(icfg[2])
    has_bias = int(icfg[9])
    nd = int(icfg[10]); dd = [int(x) for x in icfg[11:11 + nd + 1]]
    nc = int(icfg[11 + nd + 1]); dc = [int(x) for x in icfg[11 + nd + 2: 11 + nd + 2 + nc + 1]]

    def split(flat, dims):
        Ws, bs, o = [], [], 0
        for i in range(len(dims) - 1):
            n = dims[i] * dims[i + 1]
            Ws.append(flat[o:o + n].reshape(dims[i + 1], dims[i])); o += n
            if has_bias:
                bs.append(flat[o:o + dims[i + 1]]); o += dims[i + 1]
        return Ws, (bs if has_bias else None)

    dW, db = split(g["dens_params"], dd)
    cW, cb = split(g["col_params"], dc)
    nef = O.Nef([int(r) for r in g["res"]], F, int(np.log2(T)), g["table"], dW, db, cW, cb,
                multiscale="cat" if int(icfg[3]) == 0 else "sum", lod_idx=int(icfg[4]),
                pos_mode=int(icfg[5]), pos_freq=int(icfg[6]), view_mode=int(icfg[7]), view_freq=int(icfg[8]))
    spc = O.octree_to_spc(g["octree"])
    return g, nef, spc
