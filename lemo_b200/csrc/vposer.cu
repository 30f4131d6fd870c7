// VPoser decoder (reference human_body_prior/train/vposer_smpl.py:107-121, eval mode: dropout = identity):
//   z[B,32] -> LeakyReLU(fc1) -> LeakyReLU(fc2) -> out[B,126] -> per-joint 6D Gram-Schmidt -> R[B,21,3,3]
// The reference then converts R -> axis-angle (torchgeometry) and the body model converts back with
// Rodrigues; the fused fit path consumes R directly (gradient-equivalent), the aa
// output exists for the `decode(Z, 'aa')` API and the [T,72] result vectors.
#include "common.cuh"
#include "gemm.cuh"
#include "vposer.cuh"
#include "../../include/lemo_b200.h"

namespace lemo {

__global__ void k_vp_gs(const float* __restrict__ o, int n, float* __restrict__ R, float* __restrict__ aa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[6], r[9];
    for (int k = 0; k < 6; ++k) x[k] = o[i * 6 + k];
    gs6d_fwd(x, r);
    for (int k = 0; k < 9; ++k) R[i * 9 + k] = r[k];
    if (aa) {
        float a[3];
        rotmat_to_aa_tgm(r, a);
        for (int k = 0; k < 3; ++k) aa[i * 3 + k] = a[k];
    }
}
__global__ void k_vp_gs_bwd(const float* __restrict__ o, const float* __restrict__ dR, int n, float* __restrict__ d_o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[6], g[9], d[6];
    for (int k = 0; k < 6; ++k) x[k] = o[i * 6 + k];
    for (int k = 0; k < 9; ++k) g[k] = dR[i * 9 + k];
    gs6d_bwd(x, g, d);
    for (int k = 0; k < 6; ++k) d_o[i * 6 + k] = d[k];
}

template <typename T>
static int up(T** dst, const T* src, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    if (src) LEMO_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    else LEMO_CUDA(cudaMemset(*dst, 0, n * sizeof(T)));
    return 0;
}

int vposer_create(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                  int maxB, int device, VPoser** out) {
    LEMO_CHECK(w1 && b1 && w2 && b2 && w3 && b3 && out && maxB > 0, "bad arguments");
    LEMO_CUDA(cudaSetDevice(device));
    VPoser* v = new VPoser();
    v->device = device; v->maxB = maxB;
    LEMO_TRY(up(&v->W1, w1, 512 * 32)); LEMO_TRY(up(&v->b1, b1, 512));
    LEMO_TRY(up(&v->W2, w2, 512 * 512)); LEMO_TRY(up(&v->b2, b2, 512));
    LEMO_TRY(up(&v->W3, w3, 126 * 512)); LEMO_TRY(up(&v->b3, b3, 126));
    const size_t B = maxB;
    LEMO_TRY(up<float>(&v->h1, nullptr, B * 512)); LEMO_TRY(up<float>(&v->h2, nullptr, B * 512));
    LEMO_TRY(up<float>(&v->o, nullptr, B * 126)); LEMO_TRY(up<float>(&v->d_o, nullptr, B * 126));
    LEMO_TRY(up<float>(&v->dh2, nullptr, B * 512)); LEMO_TRY(up<float>(&v->dh1, nullptr, B * 512));
    *out = v;
    return 0;
}
void vposer_free(VPoser* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    float* ps[] = {v->W1, v->b1, v->W2, v->b2, v->W3, v->b3, v->h1, v->h2, v->o, v->d_o, v->dh2, v->dh1};
    for (float* p : ps) cudaFree(p);
    delete v;
}

int vposer_decode(VPoser* v, const float* z, int B, float* R_body, float* aa, cudaStream_t st) {
    LEMO_CHECK(v && z && R_body && B > 0 && B <= v->maxB, "bad arguments / batch exceeds handle size");
    GemmP g = gemm_rowmajor(z, v->W1, v->h1, B, 512, 32, true);  g.bias = v->b1; g.act = 1; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->h1, v->W2, v->h2, B, 512, 512, true);    g.bias = v->b2; g.act = 1; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->h2, v->W3, v->o, B, 126, 512, true);     g.bias = v->b3; g.act = 0; LEMO_TRY(gemm_launch(g, st));
    k_vp_gs<<<cdiv(B * NBODY, 128), 128, 0, st>>>(v->o, B * NBODY, R_body, aa);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int vposer_decode_backward(VPoser* v, const float* z, int B, const float* dR_body, float* dz, cudaStream_t st) {
    LEMO_CHECK(v && dR_body && dz && B > 0 && B <= v->maxB, "bad arguments / batch exceeds handle size");
    (void)z;
    k_vp_gs_bwd<<<cdiv(B * NBODY, 128), 128, 0, st>>>(v->o, dR_body, B * NBODY, v->d_o);
    LEMO_CUDA(cudaGetLastError());
    // dh2 = (d_o . W3) * lrelu'(h2) ;  W3 is [126,512] row-major = B operand [K=126, N=512]
    GemmP g = gemm_rowmajor(v->d_o, v->W3, v->dh2, B, 512, 126, false); g.act = 2; g.mask_src = v->h2; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->dh2, v->W2, v->dh1, B, 512, 512, false);        g.act = 2; g.mask_src = v->h1; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->dh1, v->W1, dz, B, 32, 512, false);             g.act = 0; LEMO_TRY(gemm_launch(g, st));
    return 0;
}

}  // namespace lemo
