// wb_core.cu -- error channel, launch accounting and descriptor validation of libwispb200.
#include "wb_common.cuh"
#include <stdarg.h>
#include <atomic>
#include <math.h>

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void wb_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
void wb_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int wb_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0; cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

extern "C" const char* wb_last_error(void) { return g_err; }
extern "C" int wb_version(void) { return 100; }
extern "C" int64_t wb_launch_count(void) { return g_launches.load(); }

extern "C" int wb_device_check(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { wb_set_error("wb_device_check: no CUDA device (there is no CPU fallback)"); return WB_ERR_NODEVICE; }
    WB_CHECK_ARG(device >= 0 && device < n, "device index out of range");
    int major = 0, minor = 0;
    WB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    WB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    if (major != 10) { wb_set_error("wb_device_check: device %d is sm_%d%d; libwispb200 is built for sm_100a only", device, major, minor); return WB_ERR_NODEVICE; }
    return WB_OK;
}

int wb_make_grid(const wb_nef_desc* d, WbGrid* g) {
    WB_CHECK_ARG(d != nullptr, "null descriptor");
    WB_CHECK_ARG(d->num_lods >= 1 && d->num_lods <= WB_MAX_LODS, "num_lods out of range");
    WB_CHECK_ARG(d->feature_dim >= 1 && d->feature_dim <= 8, "feature_dim must be in [1,8]");
    WB_CHECK_ARG(d->codebook_size > 0 && (d->codebook_size & (d->codebook_size - 1)) == 0, "codebook_size must be a power of two");
    WB_CHECK_ARG(d->table != nullptr, "null table");
    WB_CHECK_ARG(d->multiscale == 0 || d->multiscale == 1, "multiscale must be 0 ('cat') or 1 ('sum')");
    g->table = d->table; g->L = d->num_lods; g->F = d->feature_dim; g->Tmask = (uint32_t)d->codebook_size - 1u;
    g->multiscale = d->multiscale; g->lod_idx = d->lod_idx;
    const int64_t T = d->codebook_size;
    for (int l = 0; l < d->num_lods; ++l) {
        int res = d->resolutions[l];
        WB_CHECK_ARG(res >= 2 && res < (1 << 20), "resolution out of range");
        int64_t r2 = (int64_t)res * res, r3 = r2 * res;
        g->res[l] = res; g->hres[l] = 0.5f * (float)res;
        g->hi[l] = (float)((double)(res - 1) - 1e-5);          // clamp upper bound, cu:40
        g->dense[l] = (res < T && r2 < T && r3 < T) ? 1 : 0;   // hash_utils.cuh:27-29
        g->begin[l] = d->begin_idxes[l];
    }
    g->begin[d->num_lods] = d->begin_idxes[d->num_lods];
    return WB_OK;
}

int wb_make_march(const wb_rays* rays, int n, const float* jitter, uint32_t seed, WbMarch* m) {
    WB_CHECK_ARG(rays != nullptr && (rays->num_rays == 0 || (rays->origins && rays->dirs)), "null rays");
    WB_CHECK_ARG(rays->num_rays >= 0 && rays->num_rays < ((int64_t)1 << 31), "num_rays out of range");
    WB_CHECK_ARG(n >= 1 && n <= (1 << 20), "num_samples out of range");
    WB_CHECK_ARG((rays->near_v == nullptr) == (rays->far_v == nullptr), "near_v and far_v must both be given or both be NULL");
    m->origins = rays->origins; m->dirs = rays->dirs; m->near_v = rays->near_v; m->far_v = rays->far_v;
    m->jitter = jitter; m->near_s = rays->dist_min;
    m->range_s = (float)((double)rays->dist_max - (double)rays->dist_min);
    m->step = n > 1 ? 1.0f / (float)(n - 1) : 0.0f;
    m->n_pow2 = (n & (n - 1)) == 0; m->inv_n = 1.0f / (float)n;
    m->n = n; m->seed = seed; m->R = rays->num_rays;
    return WB_OK;
}

int wb_make_oct(const wb_octree* o, int level, WbOct* out) {
    WB_CHECK_ARG(o != nullptr && o->octree && o->prefix, "null octree");
    WB_CHECK_ARG(level >= 0 && level <= o->max_level && level <= 15, "level out of range");
    out->octree = o->octree; out->prefix = o->prefix; out->bits = o->bits; out->level = level;
    out->use_bits = (o->bits != nullptr && o->bits_level == level && level <= 10) ? 1 : 0;
    out->h = ldexpf(1.0f, level - 1); out->inv_h = ldexpf(1.0f, -(level - 1)); out->maxq = (float)((1 << level) - 1);
    out->has_bbox = (o->has_bbox && o->bits_level == level) ? 1 : 0;
    for (int a = 0; a < 3; ++a) {   // widen by 1e-4: far above the fp32 error of p = fma(d,t,o) near the unit cube
        out->blo[a] = fmaxf(o->bbox_lo[a], -1.0f) - 1e-4f; out->bhi[a] = fminf(o->bbox_hi[a], 1.0f) + 1e-4f;
    }
    const bool coarse_ok = out->has_bbox && o->coarse_bits != nullptr && o->coarse_level >= 1 && o->coarse_level < level;
    out->coarse = coarse_ok ? o->coarse_bits : nullptr; out->clevel = coarse_ok ? o->coarse_level : 0;
    out->ch = ldexpf(1.0f, out->clevel - 1); out->cmax = (float)((1 << out->clevel) - 1);
    return WB_OK;
}
