// wb_hashgrid.cu -- HashGrid.interpolate kernels: all LODs in one launch.
// Replaces wisp._C.ops.hashgrid_interpolate_cuda / hashgrid_interpolate_backward_cuda
// (wisp/csrc/ops/hashgrid_interpolate.cpp:46-105; kernels hashgrid_interpolate_cuda.cu:19-81, 83-161), which
// launch one kernel per LOD from a host loop.  One thread per (coordinate, LOD): consecutive threads write
// consecutive LODs of one coordinate, so the [N, L*F] output rows are written fully coalesced.
#include "wb_common.cuh"

template <int F>
__global__ void __launch_bounds__(256)
wb_hashgrid_fwd_kernel(WbGrid g, const float* __restrict__ coords, int64_t N, float* __restrict__ feats)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / g.L; const int l = (int)(t - i * g.L);
    if (i >= N) return;
    const float cx = __ldg(coords + 3 * i), cy = __ldg(coords + 3 * i + 1), cz = __ldg(coords + 3 * i + 2);
    uint32_t idx[8]; float cf[8];
    wb_corner_setup(g, l, cx, cy, cz, idx, cf);
    const float* tb = g.table + g.begin[l] * g.F;
    if (F == 2) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(reinterpret_cast<const float2*>(tb) + idx[j]);
        float a0 = v[0].x * cf[0], a1 = v[0].y * cf[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) { a0 = fmaf(v[j].x, cf[j], a0); a1 = fmaf(v[j].y, cf[j], a1); }
        reinterpret_cast<float2*>(feats)[i * g.L + l] = make_float2(a0, a1);
    } else {
        const int Fr = g.F;
        for (int f = 0; f < Fr; ++f) {
            float a = __ldg(tb + (int64_t)idx[0] * Fr + f) * cf[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) a = fmaf(__ldg(tb + (int64_t)idx[j] * Fr + f), cf[j], a);
            feats[(i * g.L + l) * Fr + f] = a;
        }
    }
}

extern "C" int wb_hashgrid_fwd(const float* coords, int64_t N, const wb_nef_desc* grid, float* feats, wb_stream s)
{
    WbGrid g; int rc = wb_make_grid(grid, &g); if (rc) return rc;
    if (N == 0) return WB_OK;
    WB_CHECK_ARG(coords && feats, "null pointer");
    const int64_t threads = N * g.L;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (g.F == 2) wb_hashgrid_fwd_kernel<2><<<blocks, 256, 0, (cudaStream_t)s>>>(g, coords, N, feats);
    else wb_hashgrid_fwd_kernel<0><<<blocks, 256, 0, (cudaStream_t)s>>>(g, coords, N, feats);
    WB_LAUNCH_CHECK();
    return WB_OK;
}

// backward: grad_table[idx_j] += g * coef_j (cu:151-160).  fp32 accumulate always (the reference's AMP branch
// accumulates in __half2, cu:139-149 -- lossy and order dependent; we keep fp32 masters).
template <int F>
__global__ void __launch_bounds__(256)
wb_hashgrid_bwd_kernel(WbGrid g, const float* __restrict__ coords, int64_t N, const float* __restrict__ gfeats,
                       float* __restrict__ gtable)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / g.L; const int l = (int)(t - i * g.L);
    if (i >= N) return;
    float* tb = gtable + g.begin[l] * g.F;
    if (F == 2) {
        const float2 go = __ldg(reinterpret_cast<const float2*>(gfeats) + i * g.L + l);
        if (go.x == 0.0f && go.y == 0.0f) return;
        const float cx = __ldg(coords + 3 * i), cy = __ldg(coords + 3 * i + 1), cz = __ldg(coords + 3 * i + 2);
        uint32_t idx[8]; float cf[8];
        wb_corner_setup(g, l, cx, cy, cz, idx, cf);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            atomicAdd(reinterpret_cast<float2*>(tb) + idx[j], make_float2(go.x * cf[j], go.y * cf[j]));   // red.global.add.v2.f32
    } else {
        const int Fr = g.F;
        const float cx = __ldg(coords + 3 * i), cy = __ldg(coords + 3 * i + 1), cz = __ldg(coords + 3 * i + 2);
        uint32_t idx[8]; float cf[8];
        wb_corner_setup(g, l, cx, cy, cz, idx, cf);
        for (int f = 0; f < Fr; ++f) {
            const float go = __ldg(gfeats + (i * g.L + l) * Fr + f);
            if (go == 0.0f) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) atomicAdd(tb + (int64_t)idx[j] * Fr + f, go * cf[j]);
        }
    }
}

extern "C" int wb_hashgrid_bwd(const float* coords, int64_t N, const wb_nef_desc* grid, const float* grad_feats,
                               float* grad_table, wb_stream s)
{
    WbGrid g; int rc = wb_make_grid(grid, &g); if (rc) return rc;
    if (N == 0) return WB_OK;
    WB_CHECK_ARG(coords && grad_feats && grad_table, "null pointer");
    const int64_t threads = N * g.L;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (g.F == 2) wb_hashgrid_bwd_kernel<2><<<blocks, 256, 0, (cudaStream_t)s>>>(g, coords, N, grad_feats, grad_table);
    else wb_hashgrid_bwd_kernel<0><<<blocks, 256, 0, (cudaStream_t)s>>>(g, coords, N, grad_feats, grad_table);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
