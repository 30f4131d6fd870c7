// Generic strided fp32 GEMM used OFF the critical path (VPoser MLP, LBS backward contractions, the
// sparse-row blend).  C[m,n] (+)= sum_k A[m,k] B[k,n] with arbitrary element strides so every
// transpose is a stride choice.  64x64 tile, K-step 16, 256 threads, 4x4 register tile.
// blockIdx.z is either a batch index (batch strides) or a split-K slice (atomicAdd epilogue).
#pragma once
#include "common.cuh"

namespace lemo {

struct GemmP {
    const float* A; const float* B; float* C; const float* bias;   // bias[n] or null
    int M, N, K;
    long long sAm, sAk, sBk, sBn, sCm, sCn;
    long long bA, bB, bC;      // batch strides (nz batches) -- used when splitk == 0
    int nz;                    // number of batches or of K slices
    int splitk;                // 1: blockIdx.z slices K and the epilogue is atomicAdd into C; 2: slice z stores its partial at C[z][M][N]
    int act;                   // 0 none, 1 LeakyReLU(0.2), 2 multiply by LeakyReLU'(mask_src) (ignored for splitk)
    const float* mask_src;     // act==2: same layout as C; factor = mask_src>0 ? 1 : 0.2
    int accumulate;            // non-split: C += result instead of C = result
};

int gemm_launch(const GemmP& p, cudaStream_t st);

// convenience: C[M,N] = act(A[M,K] * B + bias).  B given as W[N,K] row-major (nn.Linear weight) if b_is_nk.
inline GemmP gemm_rowmajor(const float* A, const float* B, float* C, int M, int N, int K, bool b_is_nk) {
    GemmP p{};
    p.A = A; p.B = B; p.C = C; p.bias = nullptr;
    p.M = M; p.N = N; p.K = K;
    p.sAm = K; p.sAk = 1;
    if (b_is_nk) { p.sBk = 1; p.sBn = K; } else { p.sBk = N; p.sBn = 1; }
    p.sCm = N; p.sCn = 1;
    p.nz = 1; p.splitk = 0; p.act = 0; p.accumulate = 0;
    return p;
}

}  // namespace lemo
