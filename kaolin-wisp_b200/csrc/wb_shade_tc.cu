// wb_shade_tc.cu -- fused "shade" stage with tensor-core decoders (precision 1): tcgen05.mma + TMEM + TMA-staged weights.
//
// Numerics = the reference under torch.cuda.amp.autocast (nerf_hash.yaml:76 enable_amp: True): decoder operands in
// fp16, fp32 accumulation; unlike the reference the hash table is read as fp32 master (no per-call .half() copy,
// ops/grid.py:88-89), features are blended in fp32 and table gradients accumulate in fp32.
//
// CTA = 256 threads = two 128-sample sub-tiles; thread r of a sub-tile owns sample row r == TMEM lane r.
// Per layer:  every thread writes its row of the fp16 operand tile (slab layout, wb_tc.cuh) -> fence -> CTA barrier ->
// ONE thread issues the UMMAs of both sub-tiles (A = sample tile, B = TMA-staged weight pack, D in TMEM) and commits
// to an mbarrier -> all threads wait, tcgen05.ld their accumulator row, apply bias/relu in fp32, write the next tile.
//
// Forward  (wb_shade_fwd_tc_kernel): gather 15 LODs x 8 corners (fp32 blend) -> decoders -> (r,g,b,sigma).  It also saves
//   the gathered feature rows (fp16, chunk-major [Kp0/8][S] x 16 B: coalesced both ways) for the backward.
// Backward (wb_mlp_bwd_tc_kernel): reloads those rows -- it touches neither the hash table nor the octree -- recomputes the
//   decoders (tiles stay in shared memory) and per layer issues
//     weight grad   acc_l[in, out] += X_l^T . dY_l   (both operands MN-major straight from the sample tiles; accumulators
//                                                     stay in TMEM for the whole kernel; a constant-one slab behind every
//                                                     X_l tile makes row `Kp_l` the bias gradient)
//     data grad     dX_l = dY_l . W_l                (weight pack read MN-major: no transposed copy)
//   and writes dL/dfeat as fp16 level-major planes [L][S][F].
// Scatter  (wb_table_scatter_kernel, SIMT): lanes = consecutive samples of ONE level; runs of lanes that fall into the same
//   cell are summed with a segmented warp scan and only the last lane of a run issues the 8 red.global.add.v2.f32.  The
//   per-SM atomic issue rate bounds the backward, and neighbouring samples of a ray share cells on all but the finest LODs.
// The per-ray view embedding (positional_embedder.py:51-66) is evaluated once per ray (wb_ray_embed_kernel), not per sample.
// Gradients are carried in fp16 under a power-of-two loss scale supplied on the device (no host sync) and unscaled in fp32
// at the two exits (table scatter, weight-gradient flush).
#include "wb_common.cuh"
#include "wb_tc.cuh"
#include <math.h>

#define TC_ML 16
constexpr int TC_THREADS = 256;

struct WbTc {
    int nl_d, nl_c;
    int I[TC_ML], O[TC_ML], Kp[TC_ML], Np[TC_ML];
    int w_off[TC_ML], b_off[TC_ML];            // byte offsets in the parameter blob
    int src_w[TC_ML], src_b[TC_ML];
    int blob_bytes;
    int tile_off[TC_ML];                       // byte offset of layer l's INPUT tile inside a sub-tile region
    int dy_off[2];                             // byte offset (from smem base) of the dY tile of sub-tile 0/1
    int sub_off[2];                            // byte offset of the sub-tile regions
    int w_smem_off;                            // byte offset of the staged parameter blob
    int smem_bytes;
    int acc_col[TC_ML];                        // TMEM column of the weight-grad accumulator of layer l
    int work_col[2];                           // TMEM working accumulator of sub-tile 0/1
    int tmem_cols;
    int feat_dim, pos_dim, view_dim, pos_mode, pos_freq, view_mode, view_freq;
};

static int tc_round_up(int v, int m) { return (v + m - 1) / m * m; }
static int tc_embed_dim(int mode, int freq) { return mode == 0 ? 0 : mode == 1 ? 3 : mode == 2 ? 6 * freq : 3 + 6 * freq; }

// returns WB_OK, or WB_ERR_INVALID with a message when the configuration does not fit the tensor-core path
int wb_tc_make(const wb_nef_desc* d, bool backward, WbTc* m)
{
    WB_CHECK_ARG(d->dens_layers >= 1 && d->col_layers >= 1 && d->dens_layers + d->col_layers <= TC_ML, "unsupported decoder depth");
    WB_CHECK_ARG(d->dens_params && d->col_params, "null decoder parameters");
    memset(m, 0, sizeof(*m));
    m->nl_d = d->dens_layers; m->nl_c = d->col_layers;
    m->feat_dim = d->multiscale == 0 ? d->num_lods * d->feature_dim : d->feature_dim;
    m->pos_mode = d->pos_mode; m->pos_freq = d->pos_freq; m->view_mode = d->view_mode; m->view_freq = d->view_freq;
    m->pos_dim = tc_embed_dim(d->pos_mode, d->pos_freq); m->view_dim = tc_embed_dim(d->view_mode, d->view_freq);
    WB_CHECK_ARG(d->dens_dims[0] == m->feat_dim + m->pos_dim, "decoder_density input width != grid features + position embedding");
    const int dout = d->dens_dims[d->dens_layers];
    WB_CHECK_ARG(dout >= 2 && dout <= 16, "tensor-core path: decoder_density output must be 2..16 wide");
    WB_CHECK_ARG(d->col_dims[0] == dout - 1 + m->view_dim, "decoder_color input width != density feats - 1 + view embedding");
    WB_CHECK_ARG(d->col_dims[d->col_layers] == 3, "decoder_color output must be 3 wide");
    const int nl = m->nl_d + m->nl_c;
    int off = 0, srcd = 0, srcc = 0, maxw = 0;
    for (int l = 0; l < nl; ++l) {
        const bool dens = l < m->nl_d;
        const int I = dens ? d->dens_dims[l] : d->col_dims[l - m->nl_d];
        const int O = dens ? d->dens_dims[l + 1] : d->col_dims[l - m->nl_d + 1];
        WB_CHECK_ARG(I >= 1 && I <= 128 && O >= 1 && O <= 128, "tensor-core path: layer widths must be <= 128");
        m->I[l] = I; m->O[l] = O; m->Kp[l] = tc_round_up(I, 16); m->Np[l] = tc_round_up(O, 16);
        m->w_off[l] = off; off += m->Kp[l] * m->Np[l] * 2;
        m->b_off[l] = off; off += m->Np[l] * 4;
        int& src = dens ? srcd : srcc;
        m->src_w[l] = src; src += I * O;
        if (d->has_bias) { m->src_b[l] = src; src += O; } else m->src_b[l] = -1;
        maxw = max(maxw, max(m->Kp[l], m->Np[l]));
    }
    m->blob_bytes = tc_round_up(off, 16);
    // shared memory map: [sub0 tiles][sub1 tiles][dY0][dY1][params][pad]
    int sub_bytes = 0;
    if (backward) {
        for (int l = 0; l < nl; ++l) { m->tile_off[l] = sub_bytes; sub_bytes += (m->Kp[l] / 8 + 1) * 2048; }   // + constant-one slab
    } else {
        for (int l = 0; l < nl; ++l) m->tile_off[l] = (l & 1) * (maxw / 8) * 2048;                              // ping-pong
        sub_bytes = 2 * (maxw / 8) * 2048;
    }
    m->sub_off[0] = 0; m->sub_off[1] = sub_bytes;
    int p = 2 * sub_bytes;
    if (backward) { m->dy_off[0] = p; p += (maxw / 8) * 2048; m->dy_off[1] = p; p += (maxw / 8) * 2048; }
    m->w_smem_off = p; p += m->blob_bytes;
    if (backward) {   // weight-grad MMAs read 16 slabs (M = 128 feature rows) from every X tile: keep that window inside the allocation
        const int need = m->sub_off[1] + m->tile_off[nl - 1] + 16 * 2048;
        if (p < need) p = need;
    }
    m->smem_bytes = p + 64;
    WB_CHECK_ARG(m->smem_bytes <= 227 * 1024, "tensor-core path: decoder does not fit in shared memory (use precision 0)");
    int col = 0;
    m->work_col[0] = col; col += maxw; m->work_col[1] = col; col += maxw;
    if (backward) for (int l = 0; l < nl; ++l) { m->acc_col[l] = col; col += m->Np[l]; }
    WB_CHECK_ARG(col <= 512, "tensor-core path: accumulators do not fit in TMEM (use precision 0)");
    int alloc = 32; while (alloc < col) alloc <<= 1;
    m->tmem_cols = alloc;
    return WB_OK;
}

// number of dL/dfeat planes and halfs per (plane, sample): 'cat' -> one plane per live LOD, 'sum' -> a single plane
static void tc_dfeat_shape(const wb_nef_desc* d, int* planes, int* width)
{
    *width = d->feature_dim;
    *planes = d->multiscale == 0 ? (d->lod_idx < d->num_lods ? d->lod_idx : d->num_lods) : 1;
    if (*planes < 0) *planes = 0;
}
static int64_t tc_align256(int64_t b) { return (b + 255) / 256 * 256; }

// workspace layout (bytes): [ray_embed: R * Kc * 2][dfeat: planes * S * F * 2 (backward only)]
int64_t wb_tc_workspace_bytes(const wb_nef_desc* nef, int64_t R, int64_t S, int backward)
{
    WbTc m; if (wb_tc_make(nef, false, &m)) return -1;
    int planes, width; tc_dfeat_shape(nef, &planes, &width);
    int64_t b = tc_align256(R * m.Kp[m.nl_d] * 2);
    if (backward) b += tc_align256((int64_t)planes * S * width * 2);
    return b + 256;
}
int64_t wb_tc_feat_bytes(const wb_nef_desc* nef, int64_t S)
{
    WbTc m; if (wb_tc_make(nef, false, &m)) return -1;
    return (int64_t)m.Kp[0] * 2 * S + 256;
}

// ---- parameter blob: fp16 weight packs (wb_tc.cuh layout) + fp32 biases -------------------------------------------
__global__ void wb_tc_pack_kernel(WbTc m, const float* __restrict__ dens, const float* __restrict__ col, uint8_t* __restrict__ blob)
{
    const int nl = m.nl_d + m.nl_c;
    for (int l = 0; l < nl; ++l) {
        const float* src = l < m.nl_d ? dens : col;
        const int I = m.I[l], O = m.O[l], Kp = m.Kp[l], Np = m.Np[l];
        __half* w = reinterpret_cast<__half*>(blob + m.w_off[l]);
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Kp * Np; e += gridDim.x * blockDim.x) {
            const int kc = e / (Np * 8), n = (e / 8) % Np, k = kc * 8 + (e & 7);      // element (n,k) at (k/8)*(Np*8) + n*8 + k%8 halves
            w[e] = __float2half_rn((n < O && k < I) ? src[m.src_w[l] + n * I + k] : 0.0f);
        }
        float* b = reinterpret_cast<float*>(blob + m.b_off[l]);
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Np; e += gridDim.x * blockDim.x)
            b[e] = (e < O && m.src_b[l] >= 0) ? src[m.src_b[l] + e] : 0.0f;
    }
}

int wb_tc_blob_floats(const wb_nef_desc* nef) { WbTc m; if (wb_tc_make(nef, false, &m)) return -1; return m.blob_bytes / 4; }

int wb_tc_pack(const wb_nef_desc* nef, float* blob, cudaStream_t st)
{
    WbTc m; int rc = wb_tc_make(nef, false, &m); if (rc) return rc;
    wb_tc_pack_kernel<<<16, 256, 0, st>>>(m, nef->dens_params, nef->col_params, reinterpret_cast<uint8_t*>(blob));
    WB_LAUNCH_CHECK();
    return WB_OK;
}

// ---- per-ray colour-input rows: zeros with the view embedding at features [dout-1, dout-1+view_dim) -----------------
__global__ void __launch_bounds__(128)
wb_ray_embed_kernel(WbTc m, const float* __restrict__ dirs, int64_t R, uint4* __restrict__ out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int Kc = m.Kp[m.nl_d], f0 = m.O[m.nl_d - 1] - 1;
    __align__(16) __half row[128];
    for (int i = 0; i < Kc; ++i) row[i] = __float2half_rn(0.0f);
    const float x = dirs[3 * r], y = dirs[3 * r + 1], z = dirs[3 * r + 2];
    int o = f0;
    if (m.view_mode == 1 || m.view_mode == 3) { row[o] = __float2half_rn(x); row[o + 1] = __float2half_rn(y); row[o + 2] = __float2half_rn(z); o += 3; }
    if (m.view_mode >= 2) {
        float band = 1.0f;
        for (int f = 0; f < m.view_freq; ++f) {
            const float w3[3] = { x * band, y * band, z * band };
            for (int c = 0; c < 3; ++c) {
                row[o + f * 3 + c] = __float2half_rn(sinf(w3[c]));
                row[o + 3 * m.view_freq + f * 3 + c] = __float2half_rn(cosf(w3[c]));
            }
            band *= 2.0f;
        }
    }
    const uint4* rv = reinterpret_cast<const uint4*>(row);
    for (int c = 0; c < Kc / 8; ++c) out[r * (Kc / 8) + c] = rv[c];
}

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
struct TcIn {
    const float* origins; const float* dirs; const float* rec_t; const int32_t* rec_ray; int64_t S;
    const uint4* ray_embed;      // [R][Kc/8] rows prepared by wb_ray_embed_kernel
    uint4* x0_save;              // forward: optional [Kp0/8][S] copy of the density-decoder input rows
    const uint4* x0_saved;       // backward: the same buffer
};

__device__ __forceinline__ void tile_store1(uint8_t* tile, int r, int f, float v)
{
    *reinterpret_cast<__half*>(tile + tc_slab_off(r, f)) = __float2half_rn(v);
}
__device__ __forceinline__ void tile_store8(uint8_t* tile, int r, int slab, const float v[8])
{
    uint4 q; q.x = tc_pack2(v[0], v[1]); q.y = tc_pack2(v[2], v[3]); q.z = tc_pack2(v[4], v[5]); q.w = tc_pack2(v[6], v[7]);
    *reinterpret_cast<uint4*>(tile + slab * 2048 + r * 16) = q;
}
// embedding (positional_embedder.py:51-66) written into tile features [f0, f0+dim)
__device__ __forceinline__ void tile_embed(uint8_t* tile, int r, int f0, int mode, int freq, float x, float y, float z)
{
    if (mode == 0) return;
    int o = f0;
    if (mode == 1 || mode == 3) { tile_store1(tile, r, o, x); tile_store1(tile, r, o + 1, y); tile_store1(tile, r, o + 2, z); o += 3; }
    if (mode == 1) return;
    float band = 1.0f;
    for (int f = 0; f < freq; ++f) {
        const float wx = x * band, wy = y * band, wz = z * band;
        tile_store1(tile, r, o + f * 3 + 0, sinf(wx)); tile_store1(tile, r, o + f * 3 + 1, sinf(wy)); tile_store1(tile, r, o + f * 3 + 2, sinf(wz));
        tile_store1(tile, r, o + 3 * freq + f * 3 + 0, cosf(wx)); tile_store1(tile, r, o + 3 * freq + f * 3 + 1, cosf(wy)); tile_store1(tile, r, o + 3 * freq + f * 3 + 2, cosf(wz));
        band *= 2.0f;
    }
}
// hash-grid gather of one sample -> features [0, feat_dim) of the X0 tile (fp32 blend, fp16 store)
__device__ __forceinline__ void tile_gather(const WbGrid& g, uint8_t* tile, int r, float px, float py, float pz)
{
    const int L = g.L, F = g.F;
    if (g.multiscale == 0 && F == 2) {
        for (int l0 = 0; l0 < L; l0 += 4) {                     // 4 levels = 8 features = one slab row (16 B store)
            float v[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int l = l0 + q;
                if (l >= L || l >= g.lod_idx) { v[2 * q] = 0.0f; v[2 * q + 1] = 0.0f; continue; }       // hash_grid.py:226-229
                uint32_t idx[8]; float cf[8];
                wb_corner_setup(g, l, px, py, pz, idx, cf);
                const float2* tb = reinterpret_cast<const float2*>(g.table + g.begin[l] * 2);
                float2 c[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) c[j] = __ldg(tb + idx[j]);
                float a0 = c[0].x * cf[0], a1 = c[0].y * cf[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) { a0 = fmaf(c[j].x, cf[j], a0); a1 = fmaf(c[j].y, cf[j], a1); }
                v[2 * q] = a0; v[2 * q + 1] = a1;
            }
            tile_store8(tile, r, l0 >> 2, v);
        }
    } else if (g.multiscale == 0) {
        for (int l = 0; l < L; ++l) {
            if (l >= g.lod_idx) { for (int f = 0; f < F; ++f) tile_store1(tile, r, l * F + f, 0.0f); continue; }
            uint32_t idx[8]; float cf[8];
            wb_corner_setup(g, l, px, py, pz, idx, cf);
            const float* tb = g.table + g.begin[l] * F;
            for (int f = 0; f < F; ++f) {
                float a = __ldg(tb + (int64_t)idx[0] * F + f) * cf[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) a = fmaf(__ldg(tb + (int64_t)idx[j] * F + f), cf[j], a);
                tile_store1(tile, r, l * F + f, a);
            }
        }
    } else {
        float s[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) s[f] = 0.0f;
        for (int l = 0; l < L; ++l) {
            uint32_t idx[8]; float cf[8];
            wb_corner_setup(g, l, px, py, pz, idx, cf);
            const float* tb = g.table + g.begin[l] * F;
#pragma unroll
            for (int f = 0; f < 8; ++f) if (f < F) {
                float a = __ldg(tb + (int64_t)idx[0] * F + f) * cf[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) a = fmaf(__ldg(tb + (int64_t)idx[j] * F + f), cf[j], a);
                s[f] += a;
            }
        }
#pragma unroll
        for (int f = 0; f < 8; ++f) if (f < F) tile_store1(tile, r, f, s[f]);
    }
}
// zero features [f0, f1) of this thread's row
__device__ __forceinline__ void tile_zero(uint8_t* tile, int r, int f0, int f1) { for (int f = f0; f < f1; ++f) tile_store1(tile, r, f, 0.0f); }

struct TcCtx {
    uint8_t* smem; uint64_t* bar; uint32_t tmem; uint32_t phase;
    int sub, r, laneq;          // sub-tile, row in sub-tile, 32*(warp%4)
};

// CTA-wide: operand tiles written -> one thread issues `issue()` -> everybody waits for completion
template <class IssueFn>
__device__ __forceinline__ void tc_round(TcCtx& c, IssueFn issue)
{
    tc_fence_smem_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) { tc_fence_after(); issue(); tc_commit(c.bar); }
    tc_mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    tc_fence_after();
}

// forward UMMAs of layer l for both sub-tiles: D_work[sub] = X_l . W_l^T
__device__ __forceinline__ void tc_issue_fwd(const WbTc& m, const TcCtx& c, int l)
{
    const uint32_t base = tc_smem_u32(c.smem);
    const uint32_t id = tc_idesc(128, m.Np[l], 0, 0);
    const uint32_t b0 = base + m.w_smem_off + m.w_off[l];
    for (int sub = 0; sub < 2; ++sub) {
        const uint32_t a0 = base + m.sub_off[sub] + m.tile_off[l];
        for (int kb = 0; kb < m.Kp[l] / 16; ++kb)
            tc_mma(c.tmem + m.work_col[sub], tc_desc(a0 + kb * 4096, 2048, 128), tc_desc(b0 + kb * 2 * m.Np[l] * 16, m.Np[l] * 16, 128), id, kb > 0);
    }
}
// backward UMMAs of layer l: weight grad into acc_l (persistent), data grad into D_work[sub] (N = Kp_l)
__device__ __forceinline__ void tc_issue_bwd(const WbTc& m, const TcCtx& c, int l, bool first_tile)
{
    const uint32_t base = tc_smem_u32(c.smem);
    const uint32_t idw = tc_idesc(128, m.Np[l], 1, 1);
    const uint32_t idd = tc_idesc(128, m.Kp[l], 0, 1);
    const uint32_t w0 = base + m.w_smem_off + m.w_off[l];
    for (int sub = 0; sub < 2; ++sub) {
        const uint32_t x0 = base + m.sub_off[sub] + m.tile_off[l];
        const uint32_t y0 = base + m.dy_off[sub];
        for (int kb = 0; kb < 8; ++kb)            // K = 128 samples
            tc_mma(c.tmem + m.acc_col[l], tc_desc(x0 + kb * 256, 128, 2048), tc_desc(y0 + kb * 256, 128, 2048), idw, !(first_tile && sub == 0 && kb == 0));
        for (int kb = 0; kb < m.Np[l] / 16; ++kb) // K = out features
            tc_mma(c.tmem + m.work_col[sub], tc_desc(y0 + kb * 4096, 2048, 128), tc_desc(w0 + kb * 256, 128, m.Np[l] * 16), idd, kb > 0);
    }
}

// Decoders of one 256-sample tile, starting from an X0 tile that the caller has already written.
// Returns (in registers) the density-decoder output df[16] and the colour pre-activations c3[3].
__device__ __forceinline__ void tc_decoders(const WbTc& m, TcCtx& c, const TcIn& in, int64_t ray, float df[16], float c3[3])
{
    uint8_t* sub = c.smem + m.sub_off[c.sub];
    const float* P = reinterpret_cast<const float*>(c.smem + m.w_smem_off);
    const int nl = m.nl_d + m.nl_c;
    for (int l = 0; l < nl; ++l) {
        tc_round(c, [&]() { tc_issue_fwd(m, c, l); });
        const float* bias = P + m.b_off[l] / 4;
        const uint32_t trow = c.tmem + ((uint32_t)c.laneq << 16) + m.work_col[c.sub];
        const bool last_d = (l == m.nl_d - 1), last_c = (l == nl - 1);
        if (last_d) {
            float v[16]; tc_ld16(trow, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) df[j] = v[j] + bias[j];
            // colour input = [df[1:], embed(ray_d)], zero padded (nerf.py:248-259): the per-ray row already holds the
            // embedding and the zero padding, only the first dout-1 (<= 15) features are per-sample
            uint8_t* tcol = sub + m.tile_off[l + 1];
            const int nd = m.O[l] - 1, nch = m.Kp[l + 1] / 8;
            const uint4* re = in.ray_embed + ray * nch;
            uint4 q0 = __ldg(re), q1 = __ldg(re + 1);
            {
                __half* h0 = reinterpret_cast<__half*>(&q0); __half* h1 = reinterpret_cast<__half*>(&q1);
#pragma unroll
                for (int j = 0; j < 8; ++j) { if (j < nd) h0[j] = __float2half_rn(df[j + 1]); }
#pragma unroll
                for (int j = 0; j < 7; ++j) { if (8 + j < nd) h1[j] = __float2half_rn(df[9 + j]); }
            }
            *reinterpret_cast<uint4*>(tcol + c.r * 16) = q0;
            *reinterpret_cast<uint4*>(tcol + 2048 + c.r * 16) = q1;
            for (int ch = 2; ch < nch; ++ch) *reinterpret_cast<uint4*>(tcol + ch * 2048 + c.r * 16) = __ldg(re + ch);
        } else if (last_c) {
            float v[16]; tc_ld16(trow, v);
            c3[0] = v[0] + bias[0]; c3[1] = v[1] + bias[1]; c3[2] = v[2] + bias[2];
        } else {
            uint8_t* tn = sub + m.tile_off[l + 1];
            for (int cc = 0; cc < m.Np[l]; cc += 16) {
                float v[16]; tc_ld16(trow + cc, v);
                float h[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = fmaxf(v[j] + bias[cc + j], 0.0f);          // relu
                tile_store8(tn, c.r, (cc >> 3), h); tile_store8(tn, c.r, (cc >> 3) + 1, h + 8);
            }
            // Np[l] == Kp[l+1] (both round_up(hidden,16)); padded outputs are relu(0 + 0) = 0
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS)
wb_shade_fwd_tc_kernel(WbGrid g, WbTc m, const uint8_t* __restrict__ blob, TcIn in, float4* __restrict__ shaded)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_s;
    if (threadIdx.x == 0) {
        tc_mbar_init(&bars[0], 1); tc_mbar_init(&bars[1], 1); tc_mbar_init_fence();
        tc_mbar_expect_tx(&bars[1], (uint32_t)m.blob_bytes);
        tc_bulk_g2s(smem + m.w_smem_off, blob, (uint32_t)m.blob_bytes, &bars[1]);      // TMA: parameters -> shared memory
    }
    if (threadIdx.x < 32) tc_tmem_alloc(&tmem_s, (uint32_t)m.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    tc_mbar_wait(&bars[1], 0);
    TcCtx c; c.smem = smem; c.bar = &bars[0]; c.tmem = tmem_s; c.phase = 0;
    c.sub = threadIdx.x >> 7; c.r = threadIdx.x & 127; c.laneq = ((threadIdx.x >> 5) & 3) * 32;
    uint8_t* t0 = smem + m.sub_off[c.sub] + m.tile_off[0];
    const int nch0 = m.Kp[0] / 8;
    const int64_t ntiles = (in.S + TC_THREADS - 1) / TC_THREADS;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t s = tile * TC_THREADS + threadIdx.x;
        const bool valid = s < in.S;
        if (!valid) s = in.S - 1;
        const int64_t ray = __ldg(in.rec_ray + s);
        const float t = __ldg(in.rec_t + s);
        const float px = wb_addcmul(__ldg(in.origins + 3 * ray), __ldg(in.dirs + 3 * ray), t);
        const float py = wb_addcmul(__ldg(in.origins + 3 * ray + 1), __ldg(in.dirs + 3 * ray + 1), t);
        const float pz = wb_addcmul(__ldg(in.origins + 3 * ray + 2), __ldg(in.dirs + 3 * ray + 2), t);
        // density-decoder input row: grid features (+ position embedding), zero padded to Kp
        tile_gather(g, t0, c.r, px, py, pz);
        tile_embed(t0, c.r, m.feat_dim, m.pos_mode, m.pos_freq, px, py, pz);
        tile_zero(t0, c.r, m.I[0], m.Kp[0]);
        if (in.x0_save && valid)
            for (int ch = 0; ch < nch0; ++ch) in.x0_save[(int64_t)ch * in.S + s] = *reinterpret_cast<const uint4*>(t0 + ch * 2048 + c.r * 16);
        float df[16], c3[3];
        tc_decoders(m, c, in, ray, df, c3);
        if (valid) {
            const float r = 1.0f / (1.0f + expf(-c3[0])), gg = 1.0f / (1.0f + expf(-c3[1])), b = 1.0f / (1.0f + expf(-c3[2]));
            shaded[s] = make_float4(r, gg, b, fmaxf(df[0], 0.0f));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc_tmem_dealloc(c.tmem, (uint32_t)m.tmem_cols);
}

static int tc_launch_ray_embed(const WbTc& m, const wb_rays* rays, void* workspace, cudaStream_t st)
{
    const int64_t R = rays->num_rays;
    if (R == 0) return WB_OK;
    wb_ray_embed_kernel<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(m, rays->dirs, R, reinterpret_cast<uint4*>(workspace));
    WB_LAUNCH_CHECK();
    return WB_OK;
}

int wb_tc_shade_fwd(const wb_nef_desc* nef, const float* blob, const wb_rays* rays, const float* rec_t, const int32_t* rec_ray,
                    int64_t S, float* shaded, void* feat_save, void* workspace, cudaStream_t st)
{
    WbGrid g; int rc = wb_make_grid(nef, &g); if (rc) return rc;
    WbTc m; rc = wb_tc_make(nef, false, &m); if (rc) return rc;
    WB_CHECK_ARG(workspace != nullptr, "precision 1 needs the workspace (wb_rf_workspace_bytes)");
    rc = tc_launch_ray_embed(m, rays, workspace, st); if (rc) return rc;
    TcIn in = { rays->origins, rays->dirs, rec_t, rec_ray, S, reinterpret_cast<const uint4*>(workspace), reinterpret_cast<uint4*>(feat_save), nullptr };
    WB_CUDA(cudaFuncSetAttribute(wb_shade_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m.smem_bytes));
    const int64_t ntiles = (S + TC_THREADS - 1) / TC_THREADS;
    int per_sm = (227 * 1024) / (m.smem_bytes + 2048); per_sm = max(1, min(per_sm, 512 / m.tmem_cols)); per_sm = min(per_sm, 4);
    int64_t grid = (int64_t)wb_num_sms() * per_sm; if (grid > ntiles) grid = ntiles;
    wb_shade_fwd_tc_kernel<<<(unsigned)grid, TC_THREADS, m.smem_bytes, st>>>(g, m, reinterpret_cast<const uint8_t*>(blob), in, reinterpret_cast<float4*>(shaded));
    WB_LAUNCH_CHECK();
    return WB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// decoder backward kernel
// ---------------------------------------------------------------------------------------------------------------
struct TcGrads { float* gdens; float* gcol; const float* scale; __half* dfeat; int planes, width; };

__global__ void __launch_bounds__(TC_THREADS, 1)
wb_mlp_bwd_tc_kernel(WbTc m, const uint8_t* __restrict__ blob, TcIn in, const float4* __restrict__ g_shaded, TcGrads G)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_s;
    const int nl = m.nl_d + m.nl_c;
    if (threadIdx.x == 0) {
        tc_mbar_init(&bars[0], 1); tc_mbar_init(&bars[1], 1); tc_mbar_init_fence();
        tc_mbar_expect_tx(&bars[1], (uint32_t)m.blob_bytes);
        tc_bulk_g2s(smem + m.w_smem_off, blob, (uint32_t)m.blob_bytes, &bars[1]);
    }
    if (threadIdx.x < 32) tc_tmem_alloc(&tmem_s, (uint32_t)m.tmem_cols);
    TcCtx c; c.smem = smem; c.bar = &bars[0]; c.phase = 0;
    c.sub = threadIdx.x >> 7; c.r = threadIdx.x & 127; c.laneq = ((threadIdx.x >> 5) & 3) * 32;
    uint8_t* sub = smem + m.sub_off[c.sub];
    // constant-one slab behind every input tile: feature 0 = 1, features 1..7 = 0  (bias gradient row of the weight grad)
    for (int l = 0; l < nl; ++l) {
        uint4 one; one.x = 0x00003C00u; one.y = 0; one.z = 0; one.w = 0;           // fp16 1.0 in the low half
        *reinterpret_cast<uint4*>(sub + m.tile_off[l] + (m.Kp[l] / 8) * 2048 + c.r * 16) = one;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = tmem_s;
    tc_mbar_wait(&bars[1], 0);
    const float scale = __ldg(G.scale), inv_scale = 1.0f / scale;
    uint8_t* dyt = smem + m.dy_off[c.sub];
    uint8_t* t0 = sub + m.tile_off[0];
    const int nch0 = m.Kp[0] / 8;
    const int64_t ntiles = (in.S + TC_THREADS - 1) / TC_THREADS;
    bool first = true;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int64_t s = tile * TC_THREADS + threadIdx.x;
        const bool valid = s < in.S;
        if (!valid) s = in.S - 1;
        const int64_t ray = __ldg(in.rec_ray + s);
        for (int ch = 0; ch < nch0; ++ch)                          // saved density-decoder input row (coalesced 16 B per lane)
            *reinterpret_cast<uint4*>(t0 + ch * 2048 + c.r * 16) = __ldg(in.x0_saved + (int64_t)ch * in.S + s);
        float df[16], c3[3];
        tc_decoders(m, c, in, ray, df, c3);
        float4 go = valid ? __ldg(g_shaded + s) : make_float4(0, 0, 0, 0);
        // ---- colour decoder, last layer: dY = dL/d(pre-sigmoid), zero padded ----
        {
            const float r = 1.0f / (1.0f + expf(-c3[0])), gg = 1.0f / (1.0f + expf(-c3[1])), b = 1.0f / (1.0f + expf(-c3[2]));
            float v[8] = { go.x * r * (1.0f - r) * scale, go.y * gg * (1.0f - gg) * scale, go.z * b * (1.0f - b) * scale, 0, 0, 0, 0, 0 };
            float z[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
            tile_store8(dyt, c.r, 0, v);
            for (int sl = 1; sl < m.Np[nl - 1] / 8; ++sl) tile_store8(dyt, c.r, sl, z);
        }
        const uint32_t trow = c.tmem + ((uint32_t)c.laneq << 16) + m.work_col[c.sub];
        for (int l = nl - 1; l >= 0; --l) {
            tc_round(c, [&]() { tc_issue_bwd(m, c, l, first); });
            // D_work row = dL/d(input of layer l), Kp[l] wide
            if (l == m.nl_d) {
                // first colour layer: inputs [df[1:dout], embed(ray_d)]; only the first dout-1 carry gradient (nerf.py:259)
                float v[16]; tc_ld16(trow, v);
                const int dout = m.O[m.nl_d - 1];
                float gdf[16];
                gdf[0] = (df[0] > 0.0f) ? go.w * scale : 0.0f;   // relu' of density (nerf.py:263)
#pragma unroll
                for (int j = 1; j < 16; ++j) gdf[j] = (j < dout) ? v[j - 1] : 0.0f;
                tile_store8(dyt, c.r, 0, gdf); tile_store8(dyt, c.r, 1, gdf + 8);
                float z[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
                for (int sl = 2; sl < m.Np[l - 1] / 8; ++sl) tile_store8(dyt, c.r, sl, z);
            } else if (l == 0) {
                // dL/d(grid features) -> fp16 planes [plane][S][width] (still loss-scaled); tcgen05.ld is warp-collective
                const int W = G.width, nfe = G.planes * W;
                for (int f0 = 0; f0 < nfe; f0 += 16) {
                    float v[16]; tc_ld16(trow + f0, v);
                    if (!valid) continue;
                    if (W == 2) {
#pragma unroll
                        for (int qq = 0; qq < 8; ++qq) {
                            const int pl = (f0 >> 1) + qq;
                            if (pl < G.planes) reinterpret_cast<__half2*>(G.dfeat)[(int64_t)pl * in.S + s] = __floats2half2_rn(v[2 * qq], v[2 * qq + 1]);
                        }
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            const int fe = f0 + jj;
                            if (fe < nfe) G.dfeat[((int64_t)(fe / W) * in.S + s) * W + (fe % W)] = __float2half_rn(v[jj]);
                        }
                    }
                }
            } else {
                // hidden layer input: apply relu' from the retained activation tile, write the next dY (Np[l-1] == Kp[l])
                const uint8_t* xt = sub + m.tile_off[l];
                for (int cc = 0; cc < m.Kp[l]; cc += 16) {
                    float v[16]; tc_ld16(trow + cc, v);
#pragma unroll
                    for (int hsl = 0; hsl < 2; ++hsl) {
                        const uint4 a = *reinterpret_cast<const uint4*>(xt + ((cc >> 3) + hsl) * 2048 + c.r * 16);
                        const __half2* ah = reinterpret_cast<const __half2*>(&a);
                        float o8[8];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float2 af = __half22float2(ah[p]);
                            o8[2 * p] = af.x > 0.0f ? v[hsl * 8 + 2 * p] : 0.0f;
                            o8[2 * p + 1] = af.y > 0.0f ? v[hsl * 8 + 2 * p + 1] : 0.0f;
                        }
                        tile_store8(dyt, c.r, (cc >> 3) + hsl, o8);
                    }
                }
            }
        }
        first = false;
    }
    // ---- flush weight / bias gradient accumulators (TMEM rows = input feature, row Kp = bias) ----
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 128) {
        const int row = threadIdx.x;
        const uint32_t tr = c.tmem + ((uint32_t)c.laneq << 16);
        for (int l = 0; l < nl; ++l) {
            float* gbase = l < m.nl_d ? G.gdens : G.gcol;
            const int I = m.I[l], O = m.O[l];
            // all 32 lanes of a warp must execute tcgen05.ld: decide per warp, predicate the stores per lane
            const int wrow0 = row & ~31;
            if (wrow0 > m.Kp[l]) continue;
            for (int cc = 0; cc < m.Np[l]; cc += 16) {
                float v[16]; tc_ld16(tr + m.acc_col[l] + cc, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int o = cc + j;
                    if (o >= O) continue;
                    const float val = v[j] * inv_scale;
                    if (row < I) { if (val != 0.0f) atomicAdd(gbase + m.src_w[l] + o * I + row, val); }
                    else if (row == m.Kp[l] && m.src_b[l] >= 0) atomicAdd(gbase + m.src_b[l] + o, val);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc_tmem_dealloc(c.tmem, (uint32_t)m.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// table scatter: dL/dfeat planes -> hash table (hashgrid_interpolate_cuda.cu:151-160), with warp-level run merging
// ---------------------------------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(256)
wb_table_scatter_kernel(WbGrid g, TcIn in, const __half* __restrict__ dfeat, int planes, const float* __restrict__ scale_p, float* __restrict__ gtable)
{
    const int l = blockIdx.y;                                   // level
    const int lane = threadIdx.x & 31;
    const float inv_scale = 1.0f / __ldg(scale_p);
    const int Fr = F > 0 ? F : g.F;
    const int pl = g.multiscale == 0 ? l : 0;
    float* tb = gtable + g.begin[l] * Fr;
    const int64_t nwork = (in.S + 31) & ~(int64_t)31;            // whole warps
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nwork; s += (int64_t)gridDim.x * blockDim.x) {
        const bool valid = s < in.S;
        uint32_t idx[8]; float cf[8]; uint64_t key = ~0ull - (uint64_t)lane;    // invalid lanes never merge
        float gv[F > 0 ? F : 8];
#pragma unroll
        for (int f = 0; f < (F > 0 ? F : 8); ++f) gv[f] = 0.0f;
        if (valid) {
            const int64_t ray = __ldg(in.rec_ray + s);
            const float t = __ldg(in.rec_t + s);
            const float px = wb_addcmul(__ldg(in.origins + 3 * ray), __ldg(in.dirs + 3 * ray), t);
            const float py = wb_addcmul(__ldg(in.origins + 3 * ray + 1), __ldg(in.dirs + 3 * ray + 1), t);
            const float pz = wb_addcmul(__ldg(in.origins + 3 * ray + 2), __ldg(in.dirs + 3 * ray + 2), t);
            int ix, iy, iz; float wx, wy, wz, jx, jy, jz;
            wb_cell(px, g.hres[l], g.hi[l], ix, wx, jx); wb_cell(py, g.hres[l], g.hi[l], iy, wy, jy); wb_cell(pz, g.hres[l], g.hi[l], iz, wz, jz);
            key = (uint64_t)ix | ((uint64_t)iy << 20) | ((uint64_t)iz << 40);
            const float xy00 = jx * jy, xy01 = jx * wy, xy10 = wx * jy, xy11 = wx * wy;
            cf[0] = xy00 * jz; cf[1] = xy00 * wz; cf[2] = xy01 * jz; cf[3] = xy01 * wz;
            cf[4] = xy10 * jz; cf[5] = xy10 * wz; cf[6] = xy11 * jz; cf[7] = xy11 * wz;
#pragma unroll
            for (int j = 0; j < 8; ++j) idx[j] = wb_hash_idx(ix + ((j & 4) >> 2), iy + ((j & 2) >> 1), iz + (j & 1), g.res[l], g.Tmask, g.dense[l]);
            if (F == 2) {
                const float2 gg = __half22float2(reinterpret_cast<const __half2*>(dfeat)[(int64_t)pl * in.S + s]);
                gv[0] = gg.x * inv_scale; gv[1] = gg.y * inv_scale;
            } else {
                for (int f = 0; f < Fr; ++f) gv[f] = __half2float(dfeat[((int64_t)pl * in.S + s) * Fr + f]) * inv_scale;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { idx[j] = 0; cf[j] = 0.0f; }
        }
        // runs of consecutive lanes with the same cell
        const uint64_t kprev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (kprev != key);
        const uint32_t heads = __ballot_sync(0xffffffffu, head);
        const int run_head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        const int dist = lane - run_head;
        const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
        int maxd = dist;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxd = max(maxd, __shfl_xor_sync(0xffffffffu, maxd, o));
        if (F == 2) {
            float v0[8], v1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { v0[j] = gv[0] * cf[j]; v1[j] = gv[1] * cf[j]; }
            for (int o = 1; o <= maxd; o <<= 1) {               // segmented inclusive scan (warp-uniform trip count)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float a = __shfl_up_sync(0xffffffffu, v0[j], o), b = __shfl_up_sync(0xffffffffu, v1[j], o);
                    if (dist >= o) { v0[j] += a; v1[j] += b; }
                }
            }
            if (tail && valid) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (v0[j] != 0.0f || v1[j] != 0.0f) atomicAdd(reinterpret_cast<float2*>(tb) + idx[j], make_float2(v0[j], v1[j]));
            }
        } else {
            for (int f = 0; f < Fr; ++f) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gv[f] * cf[j];
                for (int o = 1; o <= maxd; o <<= 1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float a = __shfl_up_sync(0xffffffffu, v[j], o); if (dist >= o) v[j] += a; }
                }
                if (tail && valid) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) if (v[j] != 0.0f) atomicAdd(tb + (int64_t)idx[j] * Fr + f, v[j]);
                }
            }
        }
    }
}

// decoder backward only: dL/d(shaded) -> weight gradients + dL/dfeat planes in the workspace
int wb_tc_decoder_bwd(const wb_nef_desc* nef, const float* blob, const wb_rays* rays, const float* rec_t, const int32_t* rec_ray,
                      int64_t S, const float* g_shaded, const float* scale, const void* feat_saved, void* workspace,
                      float* grad_dens, float* grad_col, cudaStream_t st)
{
    WbTc m; int rc = wb_tc_make(nef, true, &m); if (rc) return rc;
    WB_CHECK_ARG(scale != nullptr, "precision 1 needs the device loss-scale pointer");
    WB_CHECK_ARG(feat_saved != nullptr && workspace != nullptr, "precision 1 backward needs the saved features and the workspace");
    rc = tc_launch_ray_embed(m, rays, workspace, st); if (rc) return rc;
    int planes, width; tc_dfeat_shape(nef, &planes, &width);
    const int64_t R = rays->num_rays;
    __half* dfeat = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(workspace) + tc_align256(R * m.Kp[m.nl_d] * 2));
    TcIn in = { rays->origins, rays->dirs, rec_t, rec_ray, S, reinterpret_cast<const uint4*>(workspace), nullptr, reinterpret_cast<const uint4*>(feat_saved) };
    TcGrads G = { grad_dens, grad_col, scale, dfeat, planes, width };
    WB_CUDA(cudaFuncSetAttribute(wb_mlp_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m.smem_bytes));
    const int64_t ntiles = (S + TC_THREADS - 1) / TC_THREADS;
    int64_t grid = (int64_t)wb_num_sms(); if (grid > ntiles) grid = ntiles;       // 1 CTA / SM: TMEM holds the weight-grad accumulators
    wb_mlp_bwd_tc_kernel<<<(unsigned)grid, TC_THREADS, m.smem_bytes, st>>>(m, reinterpret_cast<const uint8_t*>(blob), in,
                                                                           reinterpret_cast<const float4*>(g_shaded), G);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
// table scatter only: dL/dfeat planes (written by wb_tc_decoder_bwd into the same workspace) -> grad_table
int wb_tc_table_scatter(const wb_nef_desc* nef, const wb_rays* rays, const float* rec_t, const int32_t* rec_ray, int64_t S,
                        const float* scale, void* workspace, float* grad_table, cudaStream_t st)
{
    WbGrid g; int rc = wb_make_grid(nef, &g); if (rc) return rc;
    WbTc m; rc = wb_tc_make(nef, true, &m); if (rc) return rc;
    WB_CHECK_ARG(scale != nullptr && workspace != nullptr && grad_table != nullptr, "null pointer");
    int planes, width; tc_dfeat_shape(nef, &planes, &width);
    const int64_t R = rays->num_rays;
    const __half* dfeat = reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(workspace) + tc_align256(R * m.Kp[m.nl_d] * 2));
    TcIn in = { rays->origins, rays->dirs, rec_t, rec_ray, S, nullptr, nullptr, nullptr };
    const int levels = g.multiscale == 0 ? planes : g.L;
    if (levels > 0) {
        int64_t bx = (S + 255) / 256; const int64_t cap = (int64_t)wb_num_sms() * 8; if (bx > cap) bx = cap;
        dim3 grid2((unsigned)bx, (unsigned)levels);
        if (g.F == 2) wb_table_scatter_kernel<2><<<grid2, 256, 0, st>>>(g, in, dfeat, planes, scale, grad_table);
        else wb_table_scatter_kernel<0><<<grid2, 256, 0, st>>>(g, in, dfeat, planes, scale, grad_table);
        WB_LAUNCH_CHECK();
    }
    return WB_OK;
}

int wb_tc_shade_bwd(const wb_nef_desc* nef, const float* blob, const wb_rays* rays, const float* rec_t, const int32_t* rec_ray,
                    int64_t S, const float* g_shaded, const float* scale, const void* feat_saved, void* workspace,
                    float* grad_table, float* grad_dens, float* grad_col, cudaStream_t st)
{
    int rc = wb_tc_decoder_bwd(nef, blob, rays, rec_t, rec_ray, S, g_shaded, scale, feat_saved, workspace, grad_dens, grad_col, st);
    if (rc) return rc;
    return wb_tc_table_scatter(nef, rays, rec_t, rec_ray, S, scale, workspace, grad_table, st);
}
