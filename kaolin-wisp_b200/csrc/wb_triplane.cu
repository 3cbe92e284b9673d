// wb_triplane.cu -- TriplanarGrid.interpolate (wisp/models/grids/triplanar_grid.py:98-143, TriplanarFeatureVolume.forward :205-223).
// The reference issues, per LOD, three F.grid_sample(plane[1,C,res+1,res+1], coords[..., pair], align_corners=True,
// padding_mode='reflection') launches plus stack/permute/cat copies.  Here one launch covers all LODs and planes:
// one thread per (sample, LOD, plane) blends the 4 texels of every channel and writes its C outputs at
// [sample][lod][plane][channel] -- exactly the layout the reference's cat over LODs produces.
// Coordinate handling restates ATen's grid sampler (GridSampler.h: unnormalize, reflect_coordinates, clip, bilinear
// with in-bounds masking); plane pairing: x-plane <- (y,z), y-plane <- (x,z), z-plane <- (x,y) (:217-222).
#include "wb_common.cuh"
#include "wb_featx.cuh"          // wb_reflect / wb_tp_coord: ATen grid-sampler coordinate handling

struct WbTriplane {
    int num_lods, fdim;
    int res[WB_MAX_LODS];                 // plane side = res + 1 texels
    const float* planes[WB_MAX_LODS][3];  // fmx, fmy, fmz of every LOD, each [1, fdim, res+1, res+1]
    float* gplanes[WB_MAX_LODS][3];       // gradients (backward)
};

template <bool BWD>
__global__ void __launch_bounds__(256)
wb_triplane_kernel(WbTriplane tp, const float* __restrict__ coords, int64_t N, int nl, float* __restrict__ feats, const float* __restrict__ gfeats)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = nl * 3;
    const int64_t i = t / per; const int rem = (int)(t - i * per);
    if (i >= N) return;
    const int l = rem / 3, p = rem - l * 3;
    const float cx = __ldg(coords + 3 * i), cy = __ldg(coords + 3 * i + 1), cz = __ldg(coords + 3 * i + 2);
    const float gxc = p == 0 ? cy : cx;             // grid x -> W
    const float gyc = p == 2 ? cy : cz;             // grid y -> H
    const int size = tp.res[l] + 1;
    const float ix = wb_tp_coord(gxc, size), iy = wb_tp_coord(gyc, size);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float tx = ix - fx, ty = iy - fy;
    const float nw = (1.0f - tx) * (1.0f - ty), ne = tx * (1.0f - ty), sw = (1.0f - tx) * ty, se = tx * ty;
    const bool bx1 = x1 < size, by1 = y1 < size;     // x0,y0 are always in bounds after the clip
    const int C = tp.fdim; const int64_t hw = (int64_t)size * size;
    const int64_t out = ((i * nl + l) * 3 + p) * C;
    if (!BWD) {
        const float* pl = tp.planes[l][p];
        for (int c = 0; c < C; ++c) {
            const float* ch = pl + c * hw;
            float v = __ldg(ch + (int64_t)y0 * size + x0) * nw;
            if (bx1) v += __ldg(ch + (int64_t)y0 * size + x1) * ne;
            if (by1) v += __ldg(ch + (int64_t)y1 * size + x0) * sw;
            if (bx1 && by1) v += __ldg(ch + (int64_t)y1 * size + x1) * se;
            feats[out + c] = v;
        }
    } else {
        float* pl = tp.gplanes[l][p];
        for (int c = 0; c < C; ++c) {
            const float g = __ldg(gfeats + out + c);
            if (g == 0.0f) continue;
            float* ch = pl + c * hw;
            atomicAdd(ch + (int64_t)y0 * size + x0, g * nw);
            if (bx1) atomicAdd(ch + (int64_t)y0 * size + x1, g * ne);
            if (by1) atomicAdd(ch + (int64_t)y1 * size + x0, g * sw);
            if (bx1 && by1) atomicAdd(ch + (int64_t)y1 * size + x1, g * se);
        }
    }
}

static int wb_make_triplane(int32_t num_lods, int32_t fdim, const int32_t* res, const float* const* planes, float* const* gplanes, WbTriplane* tp)
{
    WB_CHECK_ARG(num_lods >= 1 && num_lods <= WB_MAX_LODS, "num_lods out of range");
    WB_CHECK_ARG(fdim >= 1 && res && planes, "bad triplane description");
    tp->num_lods = num_lods; tp->fdim = fdim;
    for (int l = 0; l < num_lods; ++l) {
        WB_CHECK_ARG(res[l] >= 1, "plane resolution must be >= 1");
        tp->res[l] = res[l];
        for (int p = 0; p < 3; ++p) {
            WB_CHECK_ARG(planes[l * 3 + p] != nullptr, "null plane");
            tp->planes[l][p] = planes[l * 3 + p];
            tp->gplanes[l][p] = gplanes ? gplanes[l * 3 + p] : nullptr;
        }
    }
    return WB_OK;
}

extern "C" int wb_triplane_fwd(const float* coords, int64_t N, int32_t num_lods, int32_t fdim, const int32_t* res,
                               const float* const* planes, float* feats, wb_stream s)
{
    if (N == 0) return WB_OK;
    WbTriplane tp; int rc = wb_make_triplane(num_lods, fdim, res, planes, nullptr, &tp); if (rc) return rc;
    WB_CHECK_ARG(coords && feats, "null pointer");
    const int64_t threads = N * num_lods * 3;
    wb_triplane_kernel<false><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)s>>>(tp, coords, N, num_lods, feats, nullptr);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
extern "C" int wb_triplane_bwd(const float* coords, int64_t N, int32_t num_lods, int32_t fdim, const int32_t* res,
                               const float* const* planes, const float* grad_feats, float* const* grad_planes, wb_stream s)
{
    if (N == 0) return WB_OK;
    WbTriplane tp; int rc = wb_make_triplane(num_lods, fdim, res, planes, grad_planes, &tp); if (rc) return rc;
    WB_CHECK_ARG(coords && grad_feats && grad_planes, "null pointer");
    for (int i = 0; i < num_lods * 3; ++i) WB_CHECK_ARG(grad_planes[i] != nullptr, "null gradient plane");
    const int64_t threads = N * num_lods * 3;
    wb_triplane_kernel<true><<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)s>>>(tp, coords, N, num_lods, nullptr, grad_feats);
    WB_LAUNCH_CHECK();
    return WB_OK;
}

// ---- plane layout conversion for the fused path (wb_nef_desc.grid_layout = 1) ------------------------------------
struct WbRelayout { const float* src[3 * WB_X_MAX_LODS]; float* dst[3 * WB_X_MAX_LODS]; int64_t start[3 * WB_X_MAX_LODS + 1]; int n, C, to_cl; };

__global__ void __launch_bounds__(256)
wb_triplane_relayout_kernel(WbRelayout r)
{
    const int64_t total = r.start[r.n];
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int i = 0;
        while (i + 1 < r.n && t >= r.start[i + 1]) ++i;           // <= 36 planes
        const int64_t e = t - r.start[i], hw = r.start[i + 1] - r.start[i];
        const float* s = r.src[i]; float* d = r.dst[i];
        if (r.C == 4) {
            if (r.to_cl) reinterpret_cast<float4*>(d)[e] = make_float4(__ldg(s + e), __ldg(s + hw + e), __ldg(s + 2 * hw + e), __ldg(s + 3 * hw + e));
            else { const float4 v = __ldg(reinterpret_cast<const float4*>(s) + e); d[e] = v.x; d[hw + e] = v.y; d[2 * hw + e] = v.z; d[3 * hw + e] = v.w; }
        } else {
            for (int c = 0; c < r.C; ++c) { if (r.to_cl) d[e * r.C + c] = __ldg(s + c * hw + e); else d[c * hw + e] = __ldg(s + e * r.C + c); }
        }
    }
}

extern "C" int wb_triplane_relayout(const float* const* src, float* const* dst, const int32_t* sizes, int32_t n_planes, int32_t fdim,
                                    int32_t to_channel_last, wb_stream s)
{
    WB_CHECK_ARG(src && dst && sizes && n_planes >= 1 && n_planes <= 3 * WB_X_MAX_LODS && fdim >= 1 && fdim <= WB_X_MAX_C, "bad plane list");
    WbRelayout r; memset(&r, 0, sizeof(r));
    r.n = n_planes; r.C = fdim; r.to_cl = to_channel_last ? 1 : 0;
    int64_t at = 0;
    for (int i = 0; i < n_planes; ++i) {
        WB_CHECK_ARG(src[i] && dst[i] && sizes[i] >= 1, "null plane / bad size");
        r.src[i] = src[i]; r.dst[i] = dst[i]; r.start[i] = at; at += (int64_t)sizes[i] * sizes[i];
    }
    r.start[n_planes] = at;
    int64_t bx = (at + 255) / 256; const int64_t cap = (int64_t)wb_num_sms() * 8; if (bx > cap) bx = cap;
    wb_triplane_relayout_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)s>>>(r);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
