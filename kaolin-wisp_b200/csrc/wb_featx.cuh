// wb_featx.cuh -- the feature grids other than the hash grid, as sources of the fused shade kernels' decoder input:
//   kind 1  TriplanarGrid.interpolate   (wisp/models/grids/triplanar_grid.py:98-143, TriplanarFeatureVolume.forward :205-223)
//   kind 2  OctreeGrid.interpolate      (wisp/models/grids/octree_grid.py:130-219)
// One device function per direction, parameterised by how a feature leaves / a gradient enters (shared-memory column of the
// fp32 SIMT kernels, fp16 slab tile of the tensor-core kernels, fp16 dL/dfeat planes of the scatter kernel), so that
// NeuralRadianceField(TriplanarGrid | OctreeGrid) runs through the same fused pipeline as the hash grid and never leaves the GPU
// kernels for torch's nn.Linear.  The standalone interpolate kernels (wb_triplane.cu, wb_octree_grid.cu) restate the same
// arithmetic for the unfused drop-ins of grid.interpolate().
#pragma once
#include "wb_common.cuh"

constexpr int WB_X_MAX_LODS = 12;
constexpr int WB_X_MAX_C = 8;             // triplanar channels per plane
constexpr int WB_X_MAX_F = 32;            // octree features per LOD

struct WbGridX {
    int kind;                              // 0 = hash grid (WbGrid), 1 = triplanar, 2 = octree
    int nl;                                // LODs used: 0 .. lod_idx (triplanar_grid.py:127, octree_grid.py:190-211)
    int sum;                               // multiscale_type == 'sum'
    int C;                                 // triplanar: channels per plane; octree: features per LOD
    const float* ptr[3 * WB_X_MAX_LODS];   // triplanar: fmx, fmy, fmz of LOD 0, LOD 1, ...; octree: features[k]
    float* gptr[3 * WB_X_MAX_LODS];        // gradients (backward), same shapes, accumulated into
    int res[WB_X_MAX_LODS];                // triplanar: plane side - 1
    // octree
    const uint8_t* octree; const int32_t* prefix; const int16_t* points; const int32_t* trinkets;
    int base_lod, half_round;
    int chlast;                            // triplanar, C == 4: planes and plane gradients are [H, W, 4] (one float4 per texel)
};

// ---- ATen grid sampler coordinate handling (GridSampler.h), align_corners=True, padding_mode='reflection' ----
__device__ __forceinline__ float wb_reflect(float in, float span)
{   // reflect_coordinates(in, 0, 2*span)
    if (span <= 0.0f) return 0.0f;
    in = fabsf(in);
    const float extra = fmodf(in, span);
    const int flips = (int)floorf(in / span);
    return (flips & 1) ? span - extra : extra;
}
__device__ __forceinline__ float wb_tp_coord(float c, int size)
{
    float x = ((c + 1.0f) * 0.5f) * (float)(size - 1);          // grid_sampler_unnormalize, align_corners=True
    x = wb_reflect(x, (float)(size - 1));
    return fminf((float)(size - 1), fmaxf(x, 0.0f));            // clip_coordinates
}
struct WbBilinear { int o00, o01, o10, o11; float nw, ne, sw, se; bool bx1, by1; };
// plane p of a sample at (cx, cy, cz): x-plane <- (y, z), y-plane <- (x, z), z-plane <- (x, y)  (triplanar_grid.py:217-222)
__device__ __forceinline__ WbBilinear wb_tp_setup(float cx, float cy, float cz, int p, int size)
{
    const float gxc = p == 0 ? cy : cx;             // grid x -> W
    const float gyc = p == 2 ? cy : cz;             // grid y -> H
    const float ix = wb_tp_coord(gxc, size), iy = wb_tp_coord(gyc, size);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float tx = ix - fx, ty = iy - fy;
    WbBilinear b;
    b.nw = (1.0f - tx) * (1.0f - ty); b.ne = tx * (1.0f - ty); b.sw = (1.0f - tx) * ty; b.se = tx * ty;
    b.bx1 = x0 + 1 < size; b.by1 = y0 + 1 < size;   // x0, y0 are always in bounds after the clip
    b.o00 = y0 * size + x0; b.o01 = b.o00 + 1; b.o10 = b.o00 + size; b.o11 = b.o10 + 1;
    return b;
}
__device__ __forceinline__ float wb_x_h(float v) { return __half2float(__float2half_rn(v)); }

// trilinear coefficients of cell `node` of level l (coords_to_trilinear_coeffs, z fastest) and its 8 corner-feature rows
__device__ __forceinline__ void wb_oct_cell(const WbGridX& x, int node, int l, float cx, float cy, float cz, float cf[8], int tk[8])
{
    const float hl = ldexpf(1.0f, l - 1);
    const float ux = __fmaf_rn(cx, hl, hl) - (float)__ldg(x.points + 3 * (int64_t)node);
    const float uy = __fmaf_rn(cy, hl, hl) - (float)__ldg(x.points + 3 * (int64_t)node + 1);
    const float uz = __fmaf_rn(cz, hl, hl) - (float)__ldg(x.points + 3 * (int64_t)node + 2);
    const float ix = 1.0f - ux, iy = 1.0f - uy, iz = 1.0f - uz;
    cf[0] = (ix * iy) * iz; cf[1] = (ix * iy) * uz; cf[2] = (ix * uy) * iz; cf[3] = (ix * uy) * uz;
    cf[4] = (ux * iy) * iz; cf[5] = (ux * iy) * uz; cf[6] = (ux * uy) * iz; cf[7] = (ux * uy) * uz;
    const int4 t0 = __ldg(reinterpret_cast<const int4*>(x.trinkets + 8 * (int64_t)node));
    const int4 t1 = __ldg(reinterpret_cast<const int4*>(x.trinkets + 8 * (int64_t)node) + 1);
    tk[0] = t0.x; tk[1] = t0.y; tk[2] = t0.z; tk[3] = t0.w; tk[4] = t1.x; tk[5] = t1.y; tk[6] = t1.z; tk[7] = t1.w;
}
// octree descent towards the level-L cell of the quantised point; calls visit(level, node) for every level >= base_lod reached
template <class Visit>
__device__ __forceinline__ void wb_oct_walk(const WbGridX& x, float cx, float cy, float cz, Visit visit)
{
    const int L = x.base_lod + x.nl - 1;
    const float h = ldexpf(1.0f, L - 1), inv_h = ldexpf(1.0f, -(L - 1)), maxq = (float)((1 << L) - 1);
    int qx, qy, qz;
    if (!(wb_quantize(cx, h, inv_h, maxq, qx) && wb_quantize(cy, h, inv_h, maxq, qy) && wb_quantize(cz, h, inv_h, maxq, qz))) return;
    int node = 0;
    for (int l = 0; l <= L; ++l) {
        if (l > 0) {
            const int d = L - l;
            const int ci = (((qx >> d) & 1) << 2) | (((qy >> d) & 1) << 1) | ((qz >> d) & 1);
            const uint32_t b = __ldg(x.octree + node);
            if (!(b & (1u << ci))) return;            // unoccupied from here on: the remaining LODs contribute zeros
            node = __ldg(x.prefix + node) + __popc(b & ((2u << ci) - 1u));
        }
        if (l >= x.base_lod) visit(l, node);
    }
}

// Features of one sample: emit(feature index in the decoder input, value) is called exactly once per feature.
// [k0, k1): the LODs this call evaluates (default: all).  A caller that splits the LODs of a sample over two threads gets, per call,
// the 'cat' features of its LODs, or for 'sum' grids the PARTIAL sums over its LODs (emit(f, partial)), which it adds up itself.
template <class Emit>
__device__ __forceinline__ void wb_featx_gather(const WbGridX& x, float cx, float cy, float cz, Emit emit, int k0 = 0, int k1 = 1 << 30)
{
    if (k1 > x.nl) k1 = x.nl;
    if (x.kind == 1) {
        const int C = x.C;
        float acc[3][WB_X_MAX_C];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int c = 0; c < WB_X_MAX_C; ++c) acc[p][c] = 0.0f;
        for (int l = k0; l < k1; ++l) {
            const int size = x.res[l] + 1; const int64_t hw = (int64_t)size * size;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const WbBilinear b = wb_tp_setup(cx, cy, cz, p, size);
                const float* pl = x.ptr[l * 3 + p];
                if (x.chlast) {                                  // one 16-byte load per texel
                    // The four loads are unconditional (out-of-range neighbours fold onto the cell's own texel, where their weight is
                    // exactly 0: x0 == size-1 only after the clip, i.e. tx == 0), so that the loads of all planes and LODs can be in
                    // flight together instead of waiting behind the bounds branches; v + a*0 == v for the finite plane values.
                    const float4* t4 = reinterpret_cast<const float4*>(pl);
                    const float4 a00 = __ldg(t4 + b.o00), a01 = __ldg(t4 + (b.bx1 ? b.o01 : b.o00));
                    const float4 a10 = __ldg(t4 + (b.by1 ? b.o10 : b.o00)), a11 = __ldg(t4 + ((b.bx1 && b.by1) ? b.o11 : b.o00));
                    const float ne = b.bx1 ? b.ne : 0.0f, sw = b.by1 ? b.sw : 0.0f, se = (b.bx1 && b.by1) ? b.se : 0.0f;
                    float v0 = a00.x * b.nw, v1 = a00.y * b.nw, v2 = a00.z * b.nw, v3 = a00.w * b.nw;
                    v0 += a01.x * ne; v1 += a01.y * ne; v2 += a01.z * ne; v3 += a01.w * ne;
                    v0 += a10.x * sw; v1 += a10.y * sw; v2 += a10.z * sw; v3 += a10.w * sw;
                    v0 += a11.x * se; v1 += a11.y * se; v2 += a11.z * se; v3 += a11.w * se;
                    if (x.sum) { acc[p][0] += v0; acc[p][1] += v1; acc[p][2] += v2; acc[p][3] += v3; }
                    else { const int f = (l * 3 + p) * 4; emit(f, v0); emit(f + 1, v1); emit(f + 2, v2); emit(f + 3, v3); }
                    continue;
                }
#pragma unroll
                for (int c = 0; c < WB_X_MAX_C; ++c) {
                    if (c < C) {
                        const float* ch = pl + c * hw;
                        float v = __ldg(ch + b.o00) * b.nw;
                        if (b.bx1) v += __ldg(ch + b.o01) * b.ne;
                        if (b.by1) v += __ldg(ch + b.o10) * b.sw;
                        if (b.bx1 && b.by1) v += __ldg(ch + b.o11) * b.se;
                        if (x.sum) acc[p][c] += v; else emit((l * 3 + p) * C + c, v);
                    }
                }
            }
        }
        if (x.sum) {
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int c = 0; c < WB_X_MAX_C; ++c) if (c < C) emit(p * C + c, acc[p][c]);
        }
    } else {
        const int F = x.C;
        const bool sum = x.sum && x.nl > 1;           // lod_idx == 0 is a single LOD either way (octree_grid.py:190-198)
        float acc[WB_X_MAX_F];
#pragma unroll
        for (int f = 0; f < WB_X_MAX_F; ++f) acc[f] = 0.0f;
        int reached = x.base_lod;                     // 'cat': LODs below `reached` have been emitted
        wb_oct_walk(x, cx, cy, cz, [&](int l, int node) {
            const int k = l - x.base_lod;
            reached = l + 1;
            if (k < k0 || k >= k1) return;
            float cf[8]; int tk[8];
            wb_oct_cell(x, node, l, cx, cy, cz, cf, tk);
            const float* ft = x.ptr[k];
#pragma unroll
            for (int f = 0; f < WB_X_MAX_F; ++f) {
                if (f < F) {
                    float a = 0.0f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = __ldg(ft + (int64_t)tk[j] * F + f);
                        if (x.half_round) v = wb_x_h(v);  // feats.half() (octree_grid.py:147)
                        a = fmaf(v, cf[j], a);
                    }
                    if (x.half_round) a = wb_x_h(a);      // per-LOD result in fp16, then .float() (:148-149)
                    if (sum) acc[f] += a; else emit(k * F + f, a);
                }
            }
        });
        if (sum) {
#pragma unroll
            for (int f = 0; f < WB_X_MAX_F; ++f) if (f < F) emit(f, acc[f]);
        } else {
            for (int k = max(reached - x.base_lod, k0); k < k1; ++k)       // LODs the point never reached: zeros (pidx == -1)
                for (int f = 0; f < F; ++f) emit(k * F + f, 0.0f);
        }
    }
}

// Gradients of one sample: grad(feature index) -> dL/dfeat; accumulated into x.gptr with atomics
template <class Grad>
__device__ __forceinline__ void wb_featx_scatter(const WbGridX& x, float cx, float cy, float cz, Grad grad)
{
    if (x.kind == 1) {
        const int C = x.C;
        for (int l = 0; l < x.nl; ++l) {
            const int size = x.res[l] + 1; const int64_t hw = (int64_t)size * size;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const WbBilinear b = wb_tp_setup(cx, cy, cz, p, size);
                float* pl = x.gptr[l * 3 + p];
                if (x.chlast) {                                  // one 16-byte reduction per texel
                    const int f = x.sum ? p * 4 : (l * 3 + p) * 4;
                    const float g0 = grad(f), g1 = grad(f + 1), g2 = grad(f + 2), g3 = grad(f + 3);
                    if (g0 == 0.0f && g1 == 0.0f && g2 == 0.0f && g3 == 0.0f) continue;
                    float4* t4 = reinterpret_cast<float4*>(pl);
                    atomicAdd(t4 + b.o00, make_float4(g0 * b.nw, g1 * b.nw, g2 * b.nw, g3 * b.nw));
                    if (b.bx1) atomicAdd(t4 + b.o01, make_float4(g0 * b.ne, g1 * b.ne, g2 * b.ne, g3 * b.ne));
                    if (b.by1) atomicAdd(t4 + b.o10, make_float4(g0 * b.sw, g1 * b.sw, g2 * b.sw, g3 * b.sw));
                    if (b.bx1 && b.by1) atomicAdd(t4 + b.o11, make_float4(g0 * b.se, g1 * b.se, g2 * b.se, g3 * b.se));
                    continue;
                }
                for (int c = 0; c < C; ++c) {
                    const float g = grad(x.sum ? p * C + c : (l * 3 + p) * C + c);
                    if (g == 0.0f) continue;
                    float* ch = pl + c * hw;
                    atomicAdd(ch + b.o00, g * b.nw);
                    if (b.bx1) atomicAdd(ch + b.o01, g * b.ne);
                    if (b.by1) atomicAdd(ch + b.o10, g * b.sw);
                    if (b.bx1 && b.by1) atomicAdd(ch + b.o11, g * b.se);
                }
            }
        }
    } else {
        const int F = x.C;
        const bool sum = x.sum && x.nl > 1;
        wb_oct_walk(x, cx, cy, cz, [&](int l, int node) {
            float cf[8]; int tk[8];
            wb_oct_cell(x, node, l, cx, cy, cz, cf, tk);
            const int k = l - x.base_lod;
            float* gt = x.gptr[k];
            for (int f = 0; f < F; ++f) {
                const float g = grad(sum ? f : k * F + f);
                if (g == 0.0f) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) atomicAdd(gt + (int64_t)tk[j] * F + f, g * cf[j]);
            }
        });
    }
}

int wb_make_gridx(const wb_nef_desc* d, bool backward, WbGridX* x);      // host: validate + copy (wb_core.cu)
