// wb_octree_grid.cu -- OctreeGrid.interpolate and the SDF tracer's nugget cursor.
//   OctreeGrid.interpolate / _interpolate  (wisp/models/grids/octree_grid.py:130-219): the reference runs one
//     blas.query(with_parents=True) and then, per LOD, kaolin unbatched_interpolate_trilinear(feats.half()).float() and a
//     cat/sum.  Here one thread per sample descends the octree once, and for every active LOD on the way blends the 8
//     corner features found through the trinkets (level-local corner indices), in one launch for all LODs.
//     Numerics follow the call site: features are rounded to fp16 on load, blended in fp32, and each LOD's result is
//     rounded to fp16 before the float cat/sum (octree_grid.py:147-149).  [KAOLIN-EXT: parity unpinned, see DESIGN.md]
//   find_depth_bound  (wisp/ops/geometric.py:15-22 -> wisp/csrc/render/find_depth_bound_cuda.cu:16-45), quirks included.
#include "wb_common.cuh"

struct WbOctGrid {
    const int16_t* points; const int32_t* trinkets;     // [T,3], [T,8]
    const float* feats[WB_MAX_LODS]; float* gfeats[WB_MAX_LODS];
    int F, base_lod, nlods;                               // nlods = lod_idx + 1 active LODs used by this call
    int multiscale;                                       // 0 'cat', 1 'sum'
    int half_round;                                       // reproduce .half() on features and per-LOD outputs
};

__device__ __forceinline__ float wb_h(float v) { return __half2float(__float2half_rn(v)); }

template <bool BWD>
__global__ void __launch_bounds__(128)
wb_octree_interp_kernel(WbOct oc, WbOctGrid og, const float* __restrict__ coords, int64_t N, float* __restrict__ out, const float* __restrict__ gout)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float cx = __ldg(coords + 3 * i), cy = __ldg(coords + 3 * i + 1), cz = __ldg(coords + 3 * i + 2);
    const int L = oc.level, F = og.F;
    int qx, qy, qz;
    const bool in = wb_quantize(cx, oc.h, oc.inv_h, oc.maxq, qx) && wb_quantize(cy, oc.h, oc.inv_h, oc.maxq, qy) && wb_quantize(cz, oc.h, oc.inv_h, oc.maxq, qz);
    const int64_t orow = i * (og.multiscale ? F : og.nlods * F);
    if (!BWD) for (int f = 0; f < (og.multiscale ? F : og.nlods * F); ++f) out[orow + f] = 0.0f;
    if (!in) return;                                      // query == -1 on every level -> zeros (kaolin returns 0 for pidx == -1)
    int node = 0;
    for (int l = 0; l <= L; ++l) {
        if (l > 0) {                                      // one step of the descent
            const int d = L - l;
            const int ci = (((qx >> d) & 1) << 2) | (((qy >> d) & 1) << 1) | ((qz >> d) & 1);
            const uint32_t b = __ldg(oc.octree + node);
            if (!(b & (1u << ci))) return;                // unoccupied from here on: remaining LODs contribute zeros
            node = __ldg(oc.prefix + node) + __popc(b & ((2u << ci) - 1u));
        }
        const int k = l - og.base_lod;
        if (k < 0) continue;
        // trilinear coefficients of this cell: u = 2^l (c*0.5+0.5) - point  (coords_to_trilinear_coeffs)
        const float hl = ldexpf(1.0f, l - 1);
        const float ux = __fmaf_rn(cx, hl, hl) - (float)__ldg(og.points + 3 * (int64_t)node);
        const float uy = __fmaf_rn(cy, hl, hl) - (float)__ldg(og.points + 3 * (int64_t)node + 1);
        const float uz = __fmaf_rn(cz, hl, hl) - (float)__ldg(og.points + 3 * (int64_t)node + 2);
        const float ix = 1.0f - ux, iy = 1.0f - uy, iz = 1.0f - uz;
        float cf[8];
        cf[0] = (ix * iy) * iz; cf[1] = (ix * iy) * uz; cf[2] = (ix * uy) * iz; cf[3] = (ix * uy) * uz;
        cf[4] = (ux * iy) * iz; cf[5] = (ux * iy) * uz; cf[6] = (ux * uy) * iz; cf[7] = (ux * uy) * uz;
        int32_t tk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tk[j] = __ldg(og.trinkets + 8 * (int64_t)node + j);
        if (!BWD) {
            const float* ft = og.feats[k];
            for (int f = 0; f < F; ++f) {
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v = __ldg(ft + (int64_t)tk[j] * F + f);
                    if (og.half_round) v = wb_h(v);
                    acc = fmaf(v, cf[j], acc);
                }
                if (og.half_round) acc = wb_h(acc);
                if (og.multiscale) out[orow + f] += acc; else out[orow + k * F + f] = acc;
            }
        } else {
            float* gt = og.gfeats[k];
            for (int f = 0; f < F; ++f) {
                const float g = og.multiscale ? __ldg(gout + orow + f) : __ldg(gout + orow + k * F + f);
                if (g == 0.0f) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(gt + (int64_t)tk[j] * F + f, g * cf[j]);
            }
        }
        if (k == og.nlods - 1) return;
    }
}

static int wb_make_octgrid(const int16_t* points, const int32_t* trinkets, int32_t F, int32_t base_lod, int32_t nlods, int32_t multiscale,
                           int32_t half_round, const float* const* feats, float* const* gfeats, WbOctGrid* og)
{
    WB_CHECK_ARG(points && trinkets && feats, "null pointer");
    WB_CHECK_ARG(F >= 1 && nlods >= 1 && nlods <= WB_MAX_LODS && base_lod >= 0, "bad octree grid description");
    og->points = points; og->trinkets = trinkets; og->F = F; og->base_lod = base_lod; og->nlods = nlods; og->multiscale = multiscale; og->half_round = half_round;
    for (int k = 0; k < nlods; ++k) {
        WB_CHECK_ARG(feats[k] != nullptr, "null feature level");
        og->feats[k] = feats[k]; og->gfeats[k] = gfeats ? gfeats[k] : nullptr;
        if (gfeats) WB_CHECK_ARG(gfeats[k] != nullptr, "null gradient level");
    }
    return WB_OK;
}

extern "C" int wb_octree_interp_fwd(const wb_octree* oct, const int16_t* points, const int32_t* trinkets, const float* coords, int64_t N,
                                    int32_t feature_dim, int32_t base_lod, int32_t num_lods_used, int32_t multiscale, int32_t half_round,
                                    const float* const* feats, float* out, wb_stream s)
{
    if (N == 0) return WB_OK;
    WbOct oc; int rc = wb_make_oct(oct, base_lod + num_lods_used - 1, &oc); if (rc) return rc;
    WbOctGrid og; rc = wb_make_octgrid(points, trinkets, feature_dim, base_lod, num_lods_used, multiscale, half_round, feats, nullptr, &og); if (rc) return rc;
    WB_CHECK_ARG(coords && out, "null pointer");
    wb_octree_interp_kernel<false><<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)s>>>(oc, og, coords, N, out, nullptr);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
extern "C" int wb_octree_interp_bwd(const wb_octree* oct, const int16_t* points, const int32_t* trinkets, const float* coords, int64_t N,
                                    int32_t feature_dim, int32_t base_lod, int32_t num_lods_used, int32_t multiscale,
                                    const float* const* feats, const float* grad_out, float* const* grad_feats, wb_stream s)
{
    if (N == 0) return WB_OK;
    WbOct oc; int rc = wb_make_oct(oct, base_lod + num_lods_used - 1, &oc); if (rc) return rc;
    WbOctGrid og; rc = wb_make_octgrid(points, trinkets, feature_dim, base_lod, num_lods_used, multiscale, 0, feats, grad_feats, &og); if (rc) return rc;
    WB_CHECK_ARG(coords && grad_out && grad_feats, "null pointer");
    wb_octree_interp_kernel<true><<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)s>>>(oc, og, coords, N, nullptr, grad_out);
    WB_LAUNCH_CHECK();
    return WB_OK;
}

// ---- find_depth_bound --------------------------------------------------------------------------------------------------
__global__ void wb_find_depth_bound_kernel(int64_t P, int64_t num_nugs, const float* __restrict__ query, const int32_t* __restrict__ curr,
                                           int32_t* __restrict__ out, const float2* __restrict__ depth)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P) return;
    out[t] = -1;                                               // find_depth_bound.cpp: zeros - 1
    if (curr[t] <= -1) return;
    uint32_t i = (uint32_t)curr[t];
    const uint32_t mx = (t == P - 1) ? (uint32_t)P : (uint32_t)curr[t + 1];    // reference quirks kept (cu:28-29)
    const float q = query[t];
    while (i < mx && (int64_t)i < num_nugs) {
        const float2 d = __ldg(depth + i);
        if ((q >= d.x && q <= d.y) || q < d.x) { out[t] = (int32_t)i; return; }
        ++i;
    }
}
extern "C" int wb_find_depth_bound(const float* query, const int32_t* curr_idxes, const float* depth, int64_t num_packs, int64_t num_nugs,
                                   int32_t* out, wb_stream s)
{
    if (num_packs == 0) return WB_OK;
    WB_CHECK_ARG(query && curr_idxes && depth && out, "null pointer");
    wb_find_depth_bound_kernel<<<(unsigned)((num_packs + 255) / 256), 256, 0, (cudaStream_t)s>>>(num_packs, num_nugs, query, curr_idxes, out, reinterpret_cast<const float2*>(depth));
    WB_LAUNCH_CHECK();
    return WB_OK;
}
