// Scene-term device helpers shared by prox.cu (operator-level C ABI) and fit_prox.cu (fused PROX stage-2 driver).
#pragma once
#include "common.cuh"

namespace lemo {

struct Cam { float R[9]; float t[3]; float fx, fy, cx, cy; };


__device__ __forceinline__ void cam_apply(const Cam& c, const float* p, float* o) {      // o = R p + t
    o[0] = c.R[0] * p[0] + c.R[1] * p[1] + c.R[2] * p[2] + c.t[0];
    o[1] = c.R[3] * p[0] + c.R[4] * p[1] + c.R[5] * p[2] + c.t[1];
    o[2] = c.R[6] * p[0] + c.R[7] * p[1] + c.R[8] * p[2] + c.t[2];
}
__device__ __forceinline__ void cam_apply_t(const Cam& c, const float* g, float* o) {    // o = R^T g
    o[0] = c.R[0] * g[0] + c.R[3] * g[1] + c.R[6] * g[2];
    o[1] = c.R[1] * g[0] + c.R[4] * g[1] + c.R[7] * g[2];
    o[2] = c.R[2] * g[0] + c.R[5] * g[1] + c.R[8] * g[2];
}

struct Grid { float gmin[3], gmax[3]; int dim; };

// grid_sample semantics for one axis: normalise to [-1,1], un-normalise with align_corners=False, clamp to [0, dim-1] (padding 'border')
__device__ __forceinline__ void axis_coord(float p, float gmin, float gmax, int dim, float& ic, float& scale) {
    const float nrm = (p - gmin) / (gmax - gmin) * 2.f - 1.f;
    float i = ((nrm + 1.f) * (float)dim - 1.f) * 0.5f;
    scale = (float)dim / (gmax - gmin);                      // d i / d p
    if (i < 0.f) { i = 0.f; scale = 0.f; }                   // clip_coordinates_set_grad: zero gradient where clipped
    else if (i > (float)(dim - 1)) { i = (float)(dim - 1); scale = 0.f; }
    ic = i;
}

// trilinear lookup of the [dim][dim][dim] volume (indexed [x][y][z]: the reference feeds (z,y,x) as grid_sample's (W,H,D) coordinates) at
// one world point; returns the value and, if GRAD, d value / d point (zero along an axis where the coordinate was clipped)
template <bool GRAD>
__device__ __forceinline__ float sdf_eval(const float* __restrict__ sdf, const Grid& g, const float* p, float* dp) {
    float ic[3], sc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) axis_coord(p[a], g.gmin[a], g.gmax[a], g.dim, ic[a], sc[a]);
    int i0[3];
    float f[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { i0[a] = (int)floorf(ic[a]); f[a] = ic[a] - (float)i0[a]; }
    const int D = g.dim;
    float acc = 0.f, d[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int cx = 0; cx < 2; ++cx)
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
            for (int cz = 0; cz < 2; ++cz) {
                const int x = i0[0] + cx, y = i0[1] + cy, z = i0[2] + cz;
                if (x >= D || y >= D || z >= D) continue;                 // zero-weight corner past the border
                const float v = __ldg(sdf + ((size_t)x * D + y) * D + z);
                const float wx = cx ? f[0] : 1.f - f[0], wy = cy ? f[1] : 1.f - f[1], wz = cz ? f[2] : 1.f - f[2];
                acc = fmaf(v, wx * wy * wz, acc);
                if (GRAD) {
                    d[0] += v * (cx ? 1.f : -1.f) * wy * wz;
                    d[1] += v * wx * (cy ? 1.f : -1.f) * wz;
                    d[2] += v * wx * wy * (cz ? 1.f : -1.f);
                }
            }
    if (GRAD) {
#pragma unroll
        for (int a = 0; a < 3; ++a) dp[a] = d[a] * sc[a];
    }
    return acc;
}

}  // namespace lemo
