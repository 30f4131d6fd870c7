// VPoser decoder (reference human_body_prior/train/vposer_smpl.py:107-121, eval mode: dropout = identity):
//   z[B,32] -> LeakyReLU(fc1) -> LeakyReLU(fc2) -> out[B,126] -> per-joint 6D Gram-Schmidt -> R[B,21,3,3]
// The reference then converts R -> axis-angle (torchgeometry) and the body model converts back with
// Rodrigues; the fused fit path consumes R directly (gradient-equivalent), the aa
// output exists for the `decode(Z, 'aa')` API and the [T,72] result vectors.
#include "common.cuh"
#include "gemm.cuh"
#include "vposer.cuh"
#include "body.cuh"
#include <cstdlib>
#include <cstring>
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
__global__ void k_vp_gs_bwd(const float* __restrict__ o, const float* __restrict__ dR, int n, float* __restrict__ d_o, float* __restrict__ dos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x[6], g[9], d[6];
    for (int k = 0; k < 6; ++k) x[k] = o[i * 6 + k];
    for (int k = 0; k < 9; ++k) g[k] = dR[i * 9 + k];
    gs6d_bwd(x, g, d);
    for (int k = 0; k < 6; ++k) d_o[i * 6 + k] = d[k];
    if (dos) {                                 // (hi|lo) split, rows of 2 x 128 (126 padded): A operand of the adjoint GEMM
        const int b = i / NBODY, c0 = (i - b * NBODY) * 6;
        for (int k = 0; k < 6; ++k) {
            const float hi = __uint_as_float(__float_as_uint(d[k]) & 0xFFFFE000u);
            dos[(size_t)b * 256 + c0 + k] = hi;
            dos[(size_t)b * 256 + 128 + c0 + k] = d[k] - hi;
        }
    }
}
__global__ void k_vp_split(const float* __restrict__ src, int M, int K, int Kp, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * K) return;
    const int r = i / K, c = i - r * K;
    const float v = src[i], hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    dst[(size_t)r * 2 * Kp + c] = hi;
    dst[(size_t)r * 2 * Kp + Kp + c] = v - hi;
}

template <typename T>
static int up(T** dst, const T* src, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    if (src) LEMO_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    else LEMO_CUDA(cudaMemset(*dst, 0, n * sizeof(T)));
    return 0;
}

static int vposer_tc_mode();

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
    // ---- tensor-core path
    LEMO_TRY(up<float>(&v->W1r, nullptr, 512 * 32)); LEMO_TRY(up<float>(&v->W2r, nullptr, 512 * 512)); LEMO_TRY(up<float>(&v->W3r, nullptr, 126 * 512));
    LEMO_TRY(up<float>(&v->W1t, nullptr, 32 * 512)); LEMO_TRY(up<float>(&v->W2t, nullptr, 512 * 512)); LEMO_TRY(up<float>(&v->W3t, nullptr, 512 * 128));
    LEMO_TRY(up<float>(&v->L1r, nullptr, 512 * 32)); LEMO_TRY(up<float>(&v->L2r, nullptr, 512 * 512)); LEMO_TRY(up<float>(&v->L3r, nullptr, 126 * 512));
    LEMO_TRY(up<float>(&v->L1t, nullptr, 32 * 512)); LEMO_TRY(up<float>(&v->L2t, nullptr, 512 * 512)); LEMO_TRY(up<float>(&v->L3t, nullptr, 512 * 128));
    LEMO_TRY(tc_prep_b(v->W1, 512, 32, 0, 32, v->W1r, 0, v->L1r)); LEMO_TRY(tc_prep_b(v->W2, 512, 512, 0, 512, v->W2r, 0, v->L2r));
    LEMO_TRY(tc_prep_b(v->W3, 126, 512, 0, 512, v->W3r, 0, v->L3r));
    LEMO_TRY(tc_prep_b(v->W1, 512, 32, 1, 512, v->W1t, 0, v->L1t)); LEMO_TRY(tc_prep_b(v->W2, 512, 512, 1, 512, v->W2t, 0, v->L2t));
    LEMO_TRY(tc_prep_b(v->W3, 126, 512, 1, 128, v->W3t, 0, v->L3t));
    LEMO_TRY(tc_map_b(v->l_w1, v->L1r, 512, 32)); LEMO_TRY(tc_map_b(v->l_w2, v->L2r, 512, 512)); LEMO_TRY(tc_map_b(v->l_w3, v->L3r, 126, 512));
    LEMO_TRY(tc_map_b(v->l_w1t, v->L1t, 32, 512)); LEMO_TRY(tc_map_b(v->l_w2t, v->L2t, 512, 512)); LEMO_TRY(tc_map_b(v->l_w3t, v->L3t, 512, 128));
    LEMO_TRY(up<float>(&v->zs, nullptr, B * 64)); LEMO_TRY(up<float>(&v->h1s, nullptr, B * 1024)); LEMO_TRY(up<float>(&v->h2s, nullptr, B * 1024));
    LEMO_TRY(up<float>(&v->dos, nullptr, B * 256)); LEMO_TRY(up<float>(&v->dh2s, nullptr, B * 1024)); LEMO_TRY(up<float>(&v->dh1s, nullptr, B * 1024));
    LEMO_TRY(tc_map_b(v->m_w1, v->W1r, 512, 32)); LEMO_TRY(tc_map_b(v->m_w2, v->W2r, 512, 512)); LEMO_TRY(tc_map_b(v->m_w3, v->W3r, 126, 512));
    LEMO_TRY(tc_map_b(v->m_w1t, v->W1t, 32, 512)); LEMO_TRY(tc_map_b(v->m_w2t, v->W2t, 512, 512)); LEMO_TRY(tc_map_b(v->m_w3t, v->W3t, 512, 128));
    LEMO_TRY(tc_map_a(v->m_zs, v->zs, maxB, 64)); LEMO_TRY(tc_map_a(v->m_h1s, v->h1s, maxB, 1024)); LEMO_TRY(tc_map_a(v->m_h2s, v->h2s, maxB, 1024));
    LEMO_TRY(tc_map_a(v->m_dos, v->dos, maxB, 256)); LEMO_TRY(tc_map_a(v->m_dh2s, v->dh2s, maxB, 1024)); LEMO_TRY(tc_map_a(v->m_dh1s, v->dh1s, maxB, 1024));
    LEMO_CUDA(cudaDeviceSynchronize());
    v->has_tc = true;
    *out = v;
    return 0;
}
void vposer_free(VPoser* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    float* ps[] = {v->W1, v->b1, v->W2, v->b2, v->W3, v->b3, v->h1, v->h2, v->o, v->d_o, v->dh2, v->dh1, v->W1r, v->W2r, v->W3r, v->W1t, v->W2t,
                   v->W3t, v->zs, v->h1s, v->h2s, v->dos, v->dh2s, v->dh1s, v->L1r, v->L2r, v->L3r, v->L1t, v->L2t, v->L3t};
    for (float* p : ps) cudaFree(p);
    delete v;
}

// The TF32 tensor-core path is OFF by default (LEMO_VPOSER=tc enables it): measured on B200 it is both slower at the fit's shapes
// (M = 960, N = 512 gives 24 CTAs of 128x224 against 120 CTAs for the CUDA-core GEMM: 2.24 vs 2.03 ms per fitting step) and less
// accurate (R_body 4e-5 vs 1e-6 of max even with the exact 3-term split -- the tensor core's fp32 accumulation itself is only good to
// ~1e-5 over K = 512), and the rotations it produces are NOT diluted by a larger term the way blend-shape offsets are.
// A K-chunked variant (128 x 64 tiles, one TMEM accumulator per K/8 chunk summed on the CUDA cores) was measured in round 2: 5.3e-5 vs
// 2.8e-5 (fp32 CUDA cores) on R_body at B = 960, time within 6 % -- removed (see the note in blend_tc.cu).
static int vposer_tc_mode() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("LEMO_VPOSER"); on = (e && strcmp(e, "tc") == 0) ? 1 : 0; }
    return on;
}
static bool vposer_tc_enabled() { return vposer_tc_mode() != 0; }
static int vp_gemm(int mode, const void* map_a, const void* m224, const void* l224, float* C, int M, int N, int K, int lo_col, const TcEpi& ep,
                   cudaStream_t st) {
    (void)mode;
    return tc_gemm_launch(map_a, m224, C, M, N, K, lo_col, ep, st, l224);
}

int vposer_decode(VPoser* v, const float* z, int B, float* R_body, float* aa, cudaStream_t st) {
    LEMO_CHECK(v && z && R_body && B > 0 && B <= v->maxB, "bad arguments / batch exceeds handle size");
    if (v->has_tc && vposer_tc_enabled()) {
        // TF32 tensor-core MLP: A operands are exact (hi|lo) splits, weights are rounded once to TF32
        k_vp_split<<<cdiv(B * 32, 256), 256, 0, st>>>(z, B, 32, 32, v->zs);
        TcEpi e1; e1.bias = v->b1; e1.act = 1; e1.split_out = v->h1s; e1.split_ld = 1024; e1.split_lo = 512;
        const int md = vposer_tc_mode();
        LEMO_TRY(vp_gemm(md, v->m_zs, v->m_w1, v->l_w1, v->h1, B, 512, 32, 32, e1, st));
        TcEpi e2; e2.bias = v->b2; e2.act = 1; e2.split_out = v->h2s; e2.split_ld = 1024; e2.split_lo = 512;
        LEMO_TRY(vp_gemm(md, v->m_h1s, v->m_w2, v->l_w2, v->h2, B, 512, 512, 512, e2, st));
        TcEpi e3; e3.bias = v->b3;
        LEMO_TRY(vp_gemm(md, v->m_h2s, v->m_w3, v->l_w3, v->o, B, 126, 512, 512, e3, st));
        k_vp_gs<<<cdiv(B * NBODY, 128), 128, 0, st>>>(v->o, B * NBODY, R_body, aa);
        LEMO_CUDA(cudaGetLastError());
        return 0;
    }
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
    const bool tc = v->has_tc && vposer_tc_enabled();
    k_vp_gs_bwd<<<cdiv(B * NBODY, 128), 128, 0, st>>>(v->o, dR_body, B * NBODY, v->d_o, tc ? v->dos : nullptr);
    LEMO_CUDA(cudaGetLastError());
    if (tc) {
        TcEpi e1; e1.act = 2; e1.mask_src = v->h2; e1.split_out = v->dh2s; e1.split_ld = 1024; e1.split_lo = 512;
        const int md = vposer_tc_mode();
        LEMO_TRY(vp_gemm(md, v->m_dos, v->m_w3t, v->l_w3t, v->dh2, B, 512, 128, 128, e1, st));
        TcEpi e2; e2.act = 2; e2.mask_src = v->h1; e2.split_out = v->dh1s; e2.split_ld = 1024; e2.split_lo = 512;
        LEMO_TRY(vp_gemm(md, v->m_dh2s, v->m_w2t, v->l_w2t, v->dh1, B, 512, 512, 512, e2, st));
        LEMO_TRY(vp_gemm(md, v->m_dh1s, v->m_w1t, v->l_w1t, dz, B, 32, 512, 512, TcEpi{}, st));
        return 0;
    }
    // dh2 = (d_o . W3) * lrelu'(h2) ;  W3 is [126,512] row-major = B operand [K=126, N=512]
    GemmP g = gemm_rowmajor(v->d_o, v->W3, v->dh2, B, 512, 126, false); g.act = 2; g.mask_src = v->h2; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->dh2, v->W2, v->dh1, B, 512, 512, false);        g.act = 2; g.mask_src = v->h1; LEMO_TRY(gemm_launch(g, st));
    g = gemm_rowmajor(v->dh1, v->W1, dz, B, 32, 512, false);             g.act = 0; LEMO_TRY(gemm_launch(g, st));
    return 0;
}

}  // namespace lemo
