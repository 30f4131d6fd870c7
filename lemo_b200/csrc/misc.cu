// Error plumbing, element-wise rotation conversions, Adam, host-math test hooks.
#include "common.cuh"
#include "../../include/lemo_b200.h"

namespace lemo {
static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }

__global__ void k_gs6d(const float* __restrict__ x, int n, float* __restrict__ R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float xi[6], r[9];
    for (int k = 0; k < 6; ++k) xi[k] = x[i * 6 + k];
    gs6d_fwd(xi, r);
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = r[k];
}
__global__ void k_gs6d_bwd(const float* __restrict__ x, const float* __restrict__ dR, int n, float* __restrict__ dx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float xi[6], g[9], d[6];
    for (int k = 0; k < 6; ++k) xi[k] = x[i * 6 + k];
    for (int k = 0; k < 9; ++k) g[k] = dR[i * 9 + k];
    gs6d_bwd(xi, g, d);
    for (int k = 0; k < 6; ++k) dx[i * 6 + k] = d[k];
}
__global__ void k_rotmat_to_aa(const float* __restrict__ R, int n, float* __restrict__ aa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[9], a[3];
    for (int k = 0; k < 9; ++k) r[k] = R[i * 9 + k];
    rotmat_to_aa_tgm(r, a);
    for (int k = 0; k < 3; ++k) aa[i * 3 + k] = a[k];
}
__global__ void k_rotmat_to_aa_bwd(const float* __restrict__ R, const float* __restrict__ daa, int n, float* __restrict__ dR) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[9], a[3], g[9];
    for (int k = 0; k < 9; ++k) r[k] = R[i * 9 + k];
    for (int k = 0; k < 3; ++k) a[k] = daa[i * 3 + k];
    rotmat_to_aa_tgm_bwd(r, a, g);
    for (int k = 0; k < 9; ++k) dR[i * 9 + k] = g[k];
}
__global__ void k_aa_to_rot6d(const float* __restrict__ aa, int n, float* __restrict__ x6) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]}, r[9];
    aa_to_rotmat_tgm(a, r);
    // R[:, :, :2].reshape(6) -> r00 r01 r10 r11 r20 r21   (utils/utils.py:129)
    x6[i * 6 + 0] = r[0]; x6[i * 6 + 1] = r[1]; x6[i * 6 + 2] = r[3];
    x6[i * 6 + 3] = r[4]; x6[i * 6 + 4] = r[6]; x6[i * 6 + 5] = r[7];
}
__global__ void k_rodrigues(const float* __restrict__ aa, int n, float* __restrict__ R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]}, r[9];
    rodrigues_fwd(a, r);
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = r[k];
}
__global__ void k_rodrigues_bwd(const float* __restrict__ aa, const float* __restrict__ dR, int n, float* __restrict__ daa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[3] = {aa[i * 3], aa[i * 3 + 1], aa[i * 3 + 2]}, g[9], d[3];
    for (int k = 0; k < 9; ++k) g[k] = dR[i * 9 + k];
    rodrigues_bwd(a, g, d);
    for (int k = 0; k < 3; ++k) daa[i * 3 + k] = d[k];
}

// torch.optim.Adam.step (no weight decay, no amsgrad):  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                       float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
}
}  // namespace lemo

using namespace lemo;
#define ST(s) ((cudaStream_t)(s))
#define EW(kern, n, ...)                                                           \
    do {                                                                           \
        if ((n) > 0) kern<<<cdiv((n), 128), 128, 0, ST(stream)>>>(__VA_ARGS__);    \
        LEMO_CUDA(cudaGetLastError());                                             \
        return 0;                                                                  \
    } while (0)

extern "C" {
const char* lemo_last_error(void) { return lemo::g_err.c_str(); }
int lemo_version(void) { return 100; }
// debugging aid (LEMO_DEBUG_CHECK=1 in the Python loader): synchronise and report-and-clear the runtime's sticky error state, so a
// failing asynchronous launch or an unchecked API call is attributed to the C-ABI call that caused it
int lemo_debug_check(void) {
    const cudaError_t a = cudaDeviceSynchronize();
    const cudaError_t b = cudaGetLastError();
    const cudaError_t e = a != cudaSuccess ? a : b;
    if (e != cudaSuccess) lemo::set_error(std::string("CUDA error state: ") + cudaGetErrorString(e));
    return (int)e;
}

int lemo_rot6d_to_rotmat(const float* x6, int32_t n, float* R, void* stream) { EW(k_gs6d, n, x6, n, R); }
int lemo_rot6d_to_rotmat_backward(const float* x6, const float* dR, int32_t n, float* dx6, void* stream) { EW(k_gs6d_bwd, n, x6, dR, n, dx6); }
int lemo_rotmat_to_aa(const float* R, int32_t n, float* aa, void* stream) { EW(k_rotmat_to_aa, n, R, n, aa); }
int lemo_rotmat_to_aa_backward(const float* R, const float* daa, int32_t n, float* dR, void* stream) { EW(k_rotmat_to_aa_bwd, n, R, daa, n, dR); }
int lemo_aa_to_rot6d(const float* aa, int32_t n, float* x6, void* stream) { EW(k_aa_to_rot6d, n, aa, n, x6); }
int lemo_rodrigues(const float* aa, int32_t n, float* R, void* stream) { EW(k_rodrigues, n, aa, n, R); }
int lemo_rodrigues_backward(const float* aa, const float* dR, int32_t n, float* daa, void* stream) { EW(k_rodrigues_bwd, n, aa, dR, n, daa); }

int lemo_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                   int32_t t, void* stream) {
    LEMO_CHECK(p && g && m && v && t >= 1, "bad arguments");
    // bias corrections in double, as torch.optim.Adam computes them in Python floats
    const float bc1 = (float)(1.0 - pow(beta1, (double)t)), bc2s = (float)sqrt(1.0 - pow(beta2, (double)t));
    if (n > 0)
        k_adam<<<cdiv(n, 256), 256, 0, ST(stream)>>>(p, g, m, v, n, (float)lr, (float)beta1, (float)beta2, (float)eps, bc1, bc2s);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// host-side hooks so the derivative math can be unit-tested without a GPU
void lemo_host_rodrigues(const float* aa, float* R) { rodrigues_fwd(aa, R); }
void lemo_host_rodrigues_bwd(const float* aa, const float* dR, float* daa) { rodrigues_bwd(aa, dR, daa); }
void lemo_host_gs6d(const float* x6, float* R) { gs6d_fwd(x6, R); }
void lemo_host_gs6d_bwd(const float* x6, const float* dR, float* dx6) { gs6d_bwd(x6, dR, dx6); }
void lemo_host_rotmat_to_aa(const float* R, float* aa) { rotmat_to_aa_tgm(R, aa); }
void lemo_host_rotmat_to_aa_bwd(const float* R, const float* daa, float* dR) { rotmat_to_aa_tgm_bwd(R, daa, dR); }
void lemo_host_aa_to_rotmat_tgm(const float* aa, float* R) { aa_to_rotmat_tgm(aa, R); }
}
