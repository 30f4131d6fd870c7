// lemo_b200 -- shared device/host helpers (sm_100a only).
// Small 3x3 / 3x4 algebra and the closed-form forward/backward of the rotation maps used on the
// fitting hot path.  Every function is __host__ __device__ so tests/cpu_math_check.cpp can exercise the
// derivative code on the CPU against finite differences (no GPU in the build container).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <nvtx3/nvToolsExt.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#define HD __host__ __device__ __forceinline__

namespace lemo {

constexpr int NJ = 55;          // SMPL-X LBS joints
constexpr int NBETA = 20;       // 10 betas + 10 expression
constexpr int NPF = 486;        // pose-blend features (54*9)
constexpr int XK = 512;         // padded K of the fused blend GEMM: [486 pose feat | 20 betas | 6 zero]
constexpr int NBODY = 21;       // VPoser body joints

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string& s);
#define LEMO_CUDA(expr)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            lemo::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ +  \
                            ":" + std::to_string(__LINE__));                                        \
            return 1;                                                                               \
        }                                                                                           \
    } while (0)
#define LEMO_CHECK(cond, msg)                                                                       \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            lemo::set_error(std::string(msg) + " (" #cond ") @" + __FILE__ + ":" +                  \
                            std::to_string(__LINE__));                                              \
            return 2;                                                                               \
        }                                                                                           \
    } while (0)
#define LEMO_TRY(expr)                                                                              \
    do {                                                                                            \
        int _r = (expr);                                                                            \
        if (_r) return _r;                                                                          \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// NVTX range over a C-ABI entry point (header-only nvtx3: a no-op unless a tool such as nsys / ncu --nvtx is attached), so a timeline
// shows the reference-level operations (lemo_fit_run, lemo_fit_prox_run, lemo_ae_finetune_run, ...) around the kernels they enqueue.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define LEMO_NVTX(name) lemo::NvtxRange _lemo_nvtx_range(name)

// ---------------------------------------------------------------- 3x3 helpers (row-major float[9])
HD void m3_mul(const float* a, const float* b, float* c) {          // c = a b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
HD void m3_mul_bt(const float* a, const float* b, float* c) {       // c = a b^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j * 3] + a[i * 3 + 1] * b[j * 3 + 1] + a[i * 3 + 2] * b[j * 3 + 2];
}
HD void m3_mul_at(const float* a, const float* b, float* c) {       // c = a^T b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
HD void m3_vec(const float* a, const float* v, float* o) {          // o = a v
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = a[i * 3] * v[0] + a[i * 3 + 1] * v[1] + a[i * 3 + 2] * v[2];
}
HD void m3t_vec(const float* a, const float* v, float* o) {         // o = a^T v
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = a[i] * v[0] + a[3 + i] * v[1] + a[6 + i] * v[2];
}

// ---------------------------------------------------------------- Rodrigues (reference lbs.py:166-193)
// theta = || aa + 1e-8 ||, u = aa / theta, R = I + sin K + (1-cos) K^2.
HD void rodrigues_fwd(const float* aa, float* R) {
    const float ex = aa[0] + 1e-8f, ey = aa[1] + 1e-8f, ez = aa[2] + 1e-8f;
    const float th = sqrtf(ex * ex + ey * ey + ez * ez);
    const float ux = aa[0] / th, uy = aa[1] / th, uz = aa[2] / th;
    const float s = sinf(th), c1 = 1.f - cosf(th);
    const float K[9] = {0.f, -uz, uy, uz, 0.f, -ux, -uy, ux, 0.f};
    float K2[9];
    m3_mul(K, K, K2);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = s * K[i] + c1 * K2[i];
    R[0] += 1.f; R[4] += 1.f; R[8] += 1.f;
}
// d aa from dR (adjoint of rodrigues_fwd).
HD void rodrigues_bwd(const float* aa, const float* dR, float* daa) {
    const float e[3] = {aa[0] + 1e-8f, aa[1] + 1e-8f, aa[2] + 1e-8f};
    const float th = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const float ux = aa[0] / th, uy = aa[1] / th, uz = aa[2] / th;
    const float s = sinf(th), c = cosf(th), c1 = 1.f - c;
    const float K[9] = {0.f, -uz, uy, uz, 0.f, -ux, -uy, ux, 0.f};
    float K2[9];
    m3_mul(K, K, K2);
    float dK_dot = 0.f, dK2_dot = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { dK_dot += dR[i] * K[i]; dK2_dot += dR[i] * K2[i]; }
    const float dth = c * dK_dot + s * dK2_dot;
    float t1[9], t2[9], dK[9];
    m3_mul_bt(dR, K, t1);                 // dR K^T
    m3_mul_at(K, dR, t2);                 // K^T dR
#pragma unroll
    for (int i = 0; i < 9; ++i) dK[i] = s * dR[i] + c1 * (t1[i] + t2[i]);
    const float du[3] = {dK[7] - dK[5], dK[2] - dK[6], dK[3] - dK[1]};
    const float du_aa = du[0] * aa[0] + du[1] * aa[1] + du[2] * aa[2];
    const float coef = (dth - du_aa / (th * th)) / th;
#pragma unroll
    for (int i = 0; i < 3; ++i) daa[i] = du[i] / th + coef * e[i];
}

// ---------------------------------------------------------------- 6D Gram-Schmidt (reference utils/utils.py:64-70)
// x6 viewed (3,2): a = (x0,x2,x4), b = (x1,x3,x5); R columns = b1,b2,b3; F.normalize eps 1e-12.
HD void gs6d_fwd(const float* x, float* R) {
    const float a[3] = {x[0], x[2], x[4]}, b[3] = {x[1], x[3], x[5]};
    const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
    const float b1[3] = {a[0] / na, a[1] / na, a[2] / na};
    const float d = b1[0] * b[0] + b1[1] * b[1] + b1[2] * b[2];
    const float w[3] = {b[0] - d * b1[0], b[1] - d * b1[1], b[2] - d * b1[2]};
    const float nw = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
    const float b2[3] = {w[0] / nw, w[1] / nw, w[2] / nw};
    const float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
#pragma unroll
    for (int i = 0; i < 3; ++i) { R[i * 3] = b1[i]; R[i * 3 + 1] = b2[i]; R[i * 3 + 2] = b3[i]; }
}
HD void gs6d_bwd(const float* x, const float* dR, float* dx) {
    const float a[3] = {x[0], x[2], x[4]}, b[3] = {x[1], x[3], x[5]};
    const float na = fmaxf(sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]), 1e-12f);
    const float b1[3] = {a[0] / na, a[1] / na, a[2] / na};
    const float d = b1[0] * b[0] + b1[1] * b[1] + b1[2] * b[2];
    const float w[3] = {b[0] - d * b1[0], b[1] - d * b1[1], b[2] - d * b1[2]};
    const float nw = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 1e-12f);
    const float b2[3] = {w[0] / nw, w[1] / nw, w[2] / nw};
    float g1[3] = {dR[0], dR[3], dR[6]}, g2[3] = {dR[1], dR[4], dR[7]};
    const float g3[3] = {dR[2], dR[5], dR[8]};
    // b3 = b1 x b2 :  g1 += b2 x g3 ; g2 += g3 x b1
    g1[0] += b2[1] * g3[2] - b2[2] * g3[1]; g1[1] += b2[2] * g3[0] - b2[0] * g3[2]; g1[2] += b2[0] * g3[1] - b2[1] * g3[0];
    g2[0] += g3[1] * b1[2] - g3[2] * b1[1]; g2[1] += g3[2] * b1[0] - g3[0] * b1[2]; g2[2] += g3[0] * b1[1] - g3[1] * b1[0];
    // b2 = w/|w|
    const float p2 = b2[0] * g2[0] + b2[1] * g2[1] + b2[2] * g2[2];
    const float dw[3] = {(g2[0] - b2[0] * p2) / nw, (g2[1] - b2[1] * p2) / nw, (g2[2] - b2[2] * p2) / nw};
    // w = b - (b1.b) b1
    const float q = dw[0] * b1[0] + dw[1] * b1[1] + dw[2] * b1[2];
    const float db[3] = {dw[0] - q * b1[0], dw[1] - q * b1[1], dw[2] - q * b1[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i) g1[i] += -d * dw[i] - q * b[i];
    // b1 = a/|a|
    const float p1 = b1[0] * g1[0] + b1[1] * g1[1] + b1[2] * g1[2];
    const float da[3] = {(g1[0] - b1[0] * p1) / na, (g1[1] - b1[1] * p1) / na, (g1[2] - b1[2] * p1) / na};
    dx[0] = da[0]; dx[2] = da[1]; dx[4] = da[2];
    dx[1] = db[0]; dx[3] = db[1]; dx[5] = db[2];
}

// ---------------------------------------------------------------- R -> axis-angle (torchgeometry 0.1.2 algorithm)
// rotation_matrix_to_angle_axis on the 3x4-padded matrix: quaternion from the TRANSPOSED matrix with the
// four-case selection (eps 1e-6), then quaternion_to_angle_axis.  Forward only (used for [T,72] outputs).
HD void rotmat_to_aa_tgm(const float* R, float* aa) {
    // t[i][j] = R[j][i]
    const float t00 = R[0], t01 = R[3], t02 = R[6], t10 = R[1], t11 = R[4], t12 = R[7], t20 = R[2], t21 = R[5], t22 = R[8];
    float q[4], tt;
    if (t22 < 1e-6f) {
        if (t00 > t11) { tt = 1.f + t00 - t11 - t22; q[0] = t12 - t21; q[1] = tt; q[2] = t01 + t10; q[3] = t20 + t02; }
        else           { tt = 1.f - t00 + t11 - t22; q[0] = t20 - t02; q[1] = t01 + t10; q[2] = tt; q[3] = t12 + t21; }
    } else {
        if (t00 < -t11) { tt = 1.f - t00 - t11 + t22; q[0] = t01 - t10; q[1] = t20 + t02; q[2] = t12 + t21; q[3] = tt; }
        else            { tt = 1.f + t00 + t11 + t22; q[0] = tt; q[1] = t12 - t21; q[2] = t20 - t02; q[3] = t01 - t10; }
    }
    const float sc = 0.5f / sqrtf(tt);
    const float w = q[0] * sc, x = q[1] * sc, y = q[2] * sc, z = q[3] * sc;
    const float s2 = x * x + y * y + z * z, s = sqrtf(s2);
    const float two_theta = 2.f * (w < 0.f ? atan2f(-s, -w) : atan2f(s, w));
    const float k = s2 > 0.f ? two_theta / s : 2.f;
    aa[0] = x * k; aa[1] = y * k; aa[2] = z * k;
}

// adjoint of rotmat_to_aa_tgm: dR[9] from daa[3], differentiating the SELECTED quaternion branch exactly as autograd does through
// torchgeometry's mask-multiplied branches (the masks carry no gradient).  This is what lets vposer.decode(Z,'aa') and
// convert_to_3D_rot carry gradient like the reference's own graph (utils/utils.py:148, opt_amass_temp.py:356-357,
// fitting_temp_slide.py:243).  At exactly zero rotation (s2 == 0) torch's where() backward yields NaN; here k = 2 is a constant.
HD void rotmat_to_aa_tgm_bwd(const float* R, const float* daa, float* dR) {
    const float t00 = R[0], t01 = R[3], t02 = R[6], t10 = R[1], t11 = R[4], t12 = R[7], t20 = R[2], t21 = R[5], t22 = R[8];
    float q[4], tt;
    int br;
    if (t22 < 1e-6f) {
        if (t00 > t11) { br = 0; tt = 1.f + t00 - t11 - t22; q[0] = t12 - t21; q[1] = tt; q[2] = t01 + t10; q[3] = t20 + t02; }
        else           { br = 1; tt = 1.f - t00 + t11 - t22; q[0] = t20 - t02; q[1] = t01 + t10; q[2] = tt; q[3] = t12 + t21; }
    } else {
        if (t00 < -t11) { br = 2; tt = 1.f - t00 - t11 + t22; q[0] = t01 - t10; q[1] = t20 + t02; q[2] = t12 + t21; q[3] = tt; }
        else            { br = 3; tt = 1.f + t00 + t11 + t22; q[0] = tt; q[1] = t12 - t21; q[2] = t20 - t02; q[3] = t01 - t10; }
    }
    const float sc = 0.5f / sqrtf(tt);
    const float w = q[0] * sc, x = q[1] * sc, y = q[2] * sc, z = q[3] * sc;
    const float s2 = x * x + y * y + z * z, s = sqrtf(s2);
    const float two_theta = 2.f * (w < 0.f ? atan2f(-s, -w) : atan2f(s, w));
    const float k = s2 > 0.f ? two_theta / s : 2.f;
    float dv[4] = {0.f, daa[0] * k, daa[1] * k, daa[2] * k};          // d(w,x,y,z)
    if (s2 > 0.f) {
        const float dk = daa[0] * x + daa[1] * y + daa[2] * z;
        const float dtwo = dk / s;
        const float den = s2 + w * w;
        const float ds = -dk * two_theta / s2 + dtwo * 2.f * w / den;
        dv[0] = dtwo * 2.f * (-s) / den;
        const float ds2 = ds / (2.f * s);
        dv[1] += 2.f * x * ds2; dv[2] += 2.f * y * ds2; dv[3] += 2.f * z * ds2;
    }
    float dq[4];
    float dsc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) { dq[i] = dv[i] * sc; dsc += dv[i] * q[i]; }
    const float dtt = -dsc * sc / (2.f * tt);
    float g[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // gradient on t[i][j], row-major
    if (br == 0) {
        const float a = dtt + dq[1];
        g[0] += a; g[4] -= a; g[8] -= a; g[5] += dq[0]; g[7] -= dq[0]; g[1] += dq[2]; g[3] += dq[2]; g[6] += dq[3]; g[2] += dq[3];
    } else if (br == 1) {
        const float a = dtt + dq[2];
        g[0] -= a; g[4] += a; g[8] -= a; g[6] += dq[0]; g[2] -= dq[0]; g[1] += dq[1]; g[3] += dq[1]; g[5] += dq[3]; g[7] += dq[3];
    } else if (br == 2) {
        const float a = dtt + dq[3];
        g[0] -= a; g[4] -= a; g[8] += a; g[1] += dq[0]; g[3] -= dq[0]; g[6] += dq[1]; g[2] += dq[1]; g[5] += dq[2]; g[7] += dq[2];
    } else {
        const float a = dtt + dq[0];
        g[0] += a; g[4] += a; g[8] += a; g[5] += dq[1]; g[7] -= dq[1]; g[6] += dq[2]; g[2] -= dq[2]; g[1] += dq[3]; g[3] -= dq[3];
    }
    // t = R^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) dR[j * 3 + i] = g[i * 3 + j];
}

// aa -> R, torchgeometry 0.1.2 angle_axis_to_rotation_matrix (init conversion, utils/utils.py:84-90)
HD void aa_to_rotmat_tgm(const float* aa, float* R) {
    const float th2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
    if (th2 > 1e-6f) {
        const float th = sqrtf(th2);
        const float wx = aa[0] / (th + 1e-6f), wy = aa[1] / (th + 1e-6f), wz = aa[2] / (th + 1e-6f);
        const float c = cosf(th), s = sinf(th), c1 = 1.f - c;
        R[0] = c + wx * wx * c1;      R[1] = wx * wy * c1 - wz * s; R[2] = wy * s + wx * wz * c1;
        R[3] = wz * s + wx * wy * c1; R[4] = c + wy * wy * c1;      R[5] = -wx * s + wy * wz * c1;
        R[6] = -wy * s + wx * wz * c1; R[7] = wx * s + wy * wz * c1; R[8] = c + wz * wz * c1;
    } else {
        R[0] = 1.f; R[1] = -aa[2]; R[2] = aa[1];
        R[3] = aa[2]; R[4] = 1.f; R[5] = -aa[0];
        R[6] = -aa[1]; R[7] = aa[0]; R[8] = 1.f;
    }
}

#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum (blockDim.x <= 1024, multiple of 32); result valid in thread 0
__device__ __forceinline__ float block_sum(float v, float* sh /*>=32 floats*/) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
    if (w == 0) v = warp_sum(v);
    __syncthreads();
    return v;
}
__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.2f * x; }
#endif

}  // namespace lemo
