// Scene terms of the PROX stage (reference temp_prox/fitting_temp_slide.py:573-739, config 4 of BASELINE.json):
//   * PerspectiveCamera projection of the 3-D joints (temp_prox/camera.py:93-116) and its adjoint w.r.t. the points,
//   * rigid camera->world transform (fitting_temp_slide.py:673-679),
//   * signed-distance lookup  F.grid_sample(sdf, norm_vertices[:, :, [2,1,0]], padding_mode='border')  (:684-687), trilinear,
//     align_corners=False (the torch>=1.3 default the reference ran with), and its adjoint w.r.t. the query points.
// The SDF volume is SHARED by the batch (the reference replicates it B times: 6.4 GB at B=100, fit_temp_loadprox_slide.py:299).
#include "common.cuh"
#include "prox_common.cuh"
#include "../../include/lemo_b200.h"

namespace lemo {

__global__ void k_cam_project(const float* __restrict__ p, Cam c, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const float X = c.R[0] * x + c.R[1] * y + c.R[2] * z + c.t[0];
    const float Y = c.R[3] * x + c.R[4] * y + c.R[5] * z + c.t[1];
    const float Z = c.R[6] * x + c.R[7] * y + c.R[8] * z + c.t[2];
    out[i * 2] = c.fx * (X / Z) + c.cx;
    out[i * 2 + 1] = c.fy * (Y / Z) + c.cy;
}
__global__ void k_cam_project_bwd(const float* __restrict__ p, Cam c, int n, const float* __restrict__ g, float* __restrict__ dp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    const float X = c.R[0] * x + c.R[1] * y + c.R[2] * z + c.t[0];
    const float Y = c.R[3] * x + c.R[4] * y + c.R[5] * z + c.t[1];
    const float Z = c.R[6] * x + c.R[7] * y + c.R[8] * z + c.t[2];
    const float gu = g[i * 2] * c.fx, gv = g[i * 2 + 1] * c.fy;
    const float dX = gu / Z, dY = gv / Z, dZ = -(gu * X + gv * Y) / (Z * Z);
    dp[i * 3] = c.R[0] * dX + c.R[3] * dY + c.R[6] * dZ;
    dp[i * 3 + 1] = c.R[1] * dX + c.R[4] * dY + c.R[7] * dZ;
    dp[i * 3 + 2] = c.R[2] * dX + c.R[5] * dY + c.R[8] * dZ;
}
// out = R p + t   (transpose = 1: out = R^T g, the adjoint, no translation)
__global__ void k_rigid(const float* __restrict__ p, Cam c, int n, int transpose, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
    if (!transpose) {
        out[i * 3] = c.R[0] * x + c.R[1] * y + c.R[2] * z + c.t[0];
        out[i * 3 + 1] = c.R[3] * x + c.R[4] * y + c.R[5] * z + c.t[1];
        out[i * 3 + 2] = c.R[6] * x + c.R[7] * y + c.R[8] * z + c.t[2];
    } else {
        out[i * 3] = c.R[0] * x + c.R[3] * y + c.R[6] * z;
        out[i * 3 + 1] = c.R[1] * x + c.R[4] * y + c.R[7] * z;
        out[i * 3 + 2] = c.R[2] * x + c.R[5] * y + c.R[8] * z;
    }
}

template <bool BWD>
__global__ void k_sdf_sample(const float* __restrict__ pts, const float* __restrict__ sdf, Grid g, long long n, float* __restrict__ val,
                             const float* __restrict__ gval, float* __restrict__ dpts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float p[3] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
    float d[3];
    const float v = sdf_eval<BWD>(sdf, g, p, d);
    if (!BWD) val[i] = v;
    else {
        const float gv = gval[i];
#pragma unroll
        for (int a = 0; a < 3; ++a) dpts[i * 3 + a] = gv * d[a];
    }
}

}  // namespace lemo

using namespace lemo;
static Cam make_cam(const float* h_R, const float* h_t, float fx, float fy, float cx, float cy) {
    Cam c;
    for (int i = 0; i < 9; ++i) c.R[i] = h_R ? h_R[i] : (i % 4 == 0 ? 1.f : 0.f);
    for (int i = 0; i < 3; ++i) c.t[i] = h_t ? h_t[i] : 0.f;
    c.fx = fx; c.fy = fy; c.cx = cx; c.cy = cy;
    return c;
}
static Grid make_grid(const float* h_min, const float* h_max, int dim) {
    Grid g;
    for (int i = 0; i < 3; ++i) { g.gmin[i] = h_min[i]; g.gmax[i] = h_max[i]; }
    g.dim = dim;
    return g;
}

extern "C" {
int lemo_camera_project(const float* points, int64_t n, const float* h_R, const float* h_t, float fx, float fy, float cx, float cy,
                        float* out, void* stream) {
    LEMO_CHECK(points && out && n >= 0, "bad arguments");
    if (n) k_cam_project<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(points, make_cam(h_R, h_t, fx, fy, cx, cy), (int)n, out);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int lemo_camera_project_backward(const float* points, int64_t n, const float* h_R, const float* h_t, float fx, float fy, float cx, float cy,
                                 const float* d_out, float* d_points, void* stream) {
    LEMO_CHECK(points && d_out && d_points && n >= 0, "bad arguments");
    if (n) k_cam_project_bwd<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(points, make_cam(h_R, h_t, fx, fy, cx, cy), (int)n, d_out, d_points);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int lemo_rigid_transform(const float* points, int64_t n, const float* h_R, const float* h_t, int32_t adjoint, float* out, void* stream) {
    LEMO_CHECK(points && out && n >= 0 && h_R, "bad arguments");
    if (n) k_rigid<<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(points, make_cam(h_R, h_t, 1, 1, 0, 0), (int)n, adjoint, out);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int lemo_sdf_sample(const float* points, int64_t n, const float* sdf, int32_t dim, const float* h_grid_min, const float* h_grid_max,
                    float* values, void* stream) {
    LEMO_CHECK(points && sdf && values && h_grid_min && h_grid_max && dim > 1 && n >= 0, "bad arguments");
    if (n) k_sdf_sample<false><<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(points, sdf, make_grid(h_grid_min, h_grid_max, dim), n, values, nullptr, nullptr);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int lemo_sdf_sample_backward(const float* points, int64_t n, const float* sdf, int32_t dim, const float* h_grid_min, const float* h_grid_max,
                             const float* d_values, float* d_points, void* stream) {
    LEMO_CHECK(points && sdf && d_values && d_points && h_grid_min && h_grid_max && dim > 1 && n >= 0, "bad arguments");
    if (n) k_sdf_sample<true><<<cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(points, sdf, make_grid(h_grid_min, h_grid_max, dim), n, nullptr, d_values, d_points);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
}
