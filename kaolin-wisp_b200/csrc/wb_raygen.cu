// wb_raygen.cu -- camera -> rays on the device.
//   wb_raygen_lookat   _look_at / _generate_rays of the offline renderer (wisp/trainers/tracker/offline_renderer.py:23-89) over
//                      normalized_grid (wisp/ops/geometric.py:65-99): the reference builds the window with two linspace
//                      launches, a meshgrid, ~10 broadcast elementwise kernels and a normalize; a render call then needs 64 bytes
//                      of camera instead of 24 bytes per ray from the host.
//   wb_raygen_pinhole  generate_pinhole_rays (wisp/ops/raygen/raygen.py:40-85) for a pinhole camera given as plain numbers
//                      (Kaolin's Camera class is not a dependency of this library): pixel centres of generate_centered_pixel_coords
//                      (:24-31), principal point, NDC, tan(fov/2) scaling, camera-to-world rotation, normalisation.
// One thread per pixel, 24 bytes written per ray, nothing read.
#include "wb_common.cuh"

struct WbLookAt { float o[3], view[3], right[3], up[3]; float tanf; int H, W; float ax, ay, sx, sy; int ortho; };

// torch.linspace(start, end, steps)[i] (ATen: start + step*i below steps/2, end - step*(steps-1-i) above)
__device__ __forceinline__ float wb_linspace(float start, float end, int steps, int i)
{
    if (steps == 1) return start;
    const float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? __fmaf_rn(step, (float)i, start) : __fmaf_rn(-step, (float)(steps - 1 - i), end);
}

__global__ void __launch_bounds__(256)
wb_raygen_lookat_kernel(WbLookAt c, float* __restrict__ origins, float* __restrict__ dirs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)c.H * c.W) return;
    const int y = (int)(i / c.W), x = (int)(i - (int64_t)y * c.W);
    const float gx = wb_linspace(-1.0f, 1.0f, c.W, x) * c.ax;          // window_x (* width/height when wider than tall)
    const float gy = wb_linspace(1.0f, -1.0f, c.H, y) * c.ay;
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)        // ((right*gx*tan + up*gy*tan) + origin) + view, left to right as the reference's tensor expression
        p[a] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(c.right[a], gx), c.tanf), __fmul_rn(__fmul_rn(c.up[a], gy), c.tanf)), c.o[a]), c.view[a]);
    float d[3], o[3];
    if (c.ortho) { d[0] = c.view[0]; d[1] = c.view[1]; d[2] = c.view[2]; o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; }
    else { d[0] = p[0] - c.o[0]; d[1] = p[1] - c.o[1]; d[2] = p[2] - c.o[2]; o[0] = c.o[0]; o[1] = c.o[1]; o[2] = c.o[2]; }
    const float n = fmaxf(sqrtf(__fadd_rn(__fadd_rn(d[0] * d[0], d[1] * d[1]), d[2] * d[2])), 1e-12f);     // F.normalize(dim=-1)
#pragma unroll
    for (int a = 0; a < 3; ++a) { origins[3 * i + a] = o[a]; dirs[3 * i + a] = __fdiv_rn(d[a], n); }
}

extern "C" int wb_raygen_lookat(const float* origin, const float* view, const float* right, const float* up, float tan_half_fov,
                                int32_t height, int32_t width, int32_t ortho, float* origins, float* dirs, wb_stream s)
{
    WB_CHECK_ARG(origin && view && right && up && origins && dirs, "null pointer");
    WB_CHECK_ARG(height >= 1 && width >= 1 && (int64_t)height * width < ((int64_t)1 << 31), "image size out of range");
    WbLookAt c;
    for (int a = 0; a < 3; ++a) { c.o[a] = origin[a]; c.view[a] = view[a]; c.right[a] = right[a]; c.up[a] = up[a]; }
    c.tanf = tan_half_fov; c.H = height; c.W = width; c.ortho = ortho;
    c.ax = width > height ? (float)((double)width / (double)height) : 1.0f;       // use_aspect (geometric.py:92-96)
    c.ay = height > width ? (float)((double)height / (double)width) : 1.0f;
    c.sx = c.sy = 0.0f;
    const int64_t n = (int64_t)height * width;
    wb_raygen_lookat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(c, origins, dirs);
    WB_LAUNCH_CHECK();
    return WB_OK;
}

struct WbPinhole { float cam_pos[3]; float R[9]; float x0, y0, tanh, tanv; int H, W, res_x, res_y; };

__global__ void __launch_bounds__(256)
wb_raygen_pinhole_kernel(WbPinhole c, float* __restrict__ origins, float* __restrict__ dirs)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)c.res_y * c.res_x) return;
    const int iy = (int)(i / c.res_x), ix = (int)(i - (int64_t)iy * c.res_x);
    // generate_centered_pixel_coords (raygen.py:24-31): pixel centre, scaled when the ray grid is coarser than the image
    float px = (float)ix * ((float)c.W / (float)c.res_x) + 0.5f, py = (float)iy * ((float)c.H / (float)c.res_y) + 0.5f;
    px = px - c.x0; py = py + c.y0;                                            // principal point (:66-67)
    px = 2.0f * (px / (float)c.W) - 1.0f; py = 2.0f * (py / (float)c.H) - 1.0f;      // _to_ndc_coords (:35-38)
    const float dc[3] = { px * c.tanh, -py * c.tanv, -1.0f };                       // (:72-74)
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) d[a] = c.R[a * 3] * dc[0] + c.R[a * 3 + 1] * dc[1] + c.R[a * 3 + 2] * dc[2];    // camera -> world rotation
    const float n = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);                 // ray_dir /= norm (:81)
#pragma unroll
    for (int a = 0; a < 3; ++a) { origins[3 * i + a] = c.cam_pos[a]; dirs[3 * i + a] = d[a] / n; }
}

extern "C" int wb_raygen_pinhole(const float* cam_pos, const float* cam_to_world_rot, float x0, float y0, float tan_half_fov_h, float tan_half_fov_v,
                                 int32_t img_height, int32_t img_width, int32_t res_y, int32_t res_x, float* origins, float* dirs, wb_stream s)
{
    WB_CHECK_ARG(cam_pos && cam_to_world_rot && origins && dirs, "null pointer");
    WB_CHECK_ARG(img_height >= 1 && img_width >= 1 && res_x >= 1 && res_y >= 1 && (int64_t)res_x * res_y < ((int64_t)1 << 31), "image size out of range");
    WbPinhole c;
    for (int a = 0; a < 3; ++a) c.cam_pos[a] = cam_pos[a];
    for (int a = 0; a < 9; ++a) c.R[a] = cam_to_world_rot[a];
    c.x0 = x0; c.y0 = y0; c.tanh = tan_half_fov_h; c.tanv = tan_half_fov_v; c.H = img_height; c.W = img_width; c.res_x = res_x; c.res_y = res_y;
    const int64_t n = (int64_t)res_x * res_y;
    wb_raygen_pinhole_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(c, origins, dirs);
    WB_LAUNCH_CHECK();
    return WB_OK;
}
