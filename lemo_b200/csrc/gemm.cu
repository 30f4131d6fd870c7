// Generic strided fp32 GEMM (see gemm.cuh).  Plain CUDA-core code: this kernel is deliberately the
// "utility" path; the roofline kernels (fused LBS forward, conv3x3) have their own files.
#include "gemm.cuh"
#include <cstdlib>

namespace lemo {

constexpr int GBN = 64, GBK = 16;

// GBM x 64 tile, 4 x 4 outputs per thread, GBM * 4 threads.  GBM = 32 halves the tile for the small problems of the fit (M = S*T <= 960 rows):
// twice the CTAs, several resident per SM, so one CTA's barrier bubbles are covered by another's math.  The K order per output element
// is the same for both tiles, so results are bit-identical.
template <int GBM>
__global__ void __launch_bounds__(GBM * 4) k_gemm(GemmP p) {
    constexpr int NT = GBM * 4, NB = GBN * GBK / NT;
    __shared__ float As[GBK][GBM + 4];
    __shared__ float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const float* A = p.A;
    const float* B = p.B;
    float* C = p.C;
    int k_begin = 0, k_end = p.K;
    if (p.splitk) {
        const int chunk = ((p.K + p.nz - 1) / p.nz + GBK - 1) / GBK * GBK;
        k_begin = blockIdx.z * chunk;
        k_end = min(p.K, k_begin + chunk);
    } else {
        A += (long long)blockIdx.z * p.bA;
        B += (long long)blockIdx.z * p.bB;
        C += (long long)blockIdx.z * p.bC;
    }
    const bool a_kfast = (p.sAk == 1);
    const bool b_nfast = (p.sBn == 1);
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // Register prefetch two K-blocks ahead: these GEMMs run with at most one CTA (8 warps) per SM, so nothing else hides the ~1 us
    // global-load latency, and one K-block of math is only ~500 cycles.  Block k+2 is requested while block k is multiplied; the loop is
    // unrolled by two so both register stages are statically indexed.  (ncu: 52 us -> 36 us with distance 1 for the 0.5 GFLOP layers.)
    float ra0[4], rb0[NB], ra1[4], rb1[NB];
    auto load_tile = [&](int k0, float (&ra)[4], float (&rb)[NB]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * NT;
            int m, k;
            if (a_kfast) { k = idx & 15; m = idx >> 4; } else { m = idx & (GBM - 1); k = idx / GBM; }
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < p.M && gk < k_end) ? A[gm * p.sAm + gk * p.sAk] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = tid + i * NT;
            int n, kb;
            if (b_nfast) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            const int gn = n0 + n, gkb = k0 + kb;
            rb[i] = (gn < p.N && gkb < k_end) ? B[gkb * p.sBk + gn * p.sBn] : 0.f;
        }
    };
    auto store_tile = [&](const float (&ra)[4], const float (&rb)[NB]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * NT;
            int m, k;
            if (a_kfast) { k = idx & 15; m = idx >> 4; } else { m = idx & (GBM - 1); k = idx / GBM; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int idx = tid + i * NT;
            int n, kb;
            if (b_nfast) { n = idx & 63; kb = idx >> 6; } else { kb = idx & 15; n = idx >> 4; }
            Bs[kb][n] = rb[i];
        }
    };
    auto mul_tile = [&]() {
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    };
    if (k_begin < k_end) load_tile(k_begin, ra0, rb0);
    if (k_begin + GBK < k_end) load_tile(k_begin + GBK, ra1, rb1);
    for (int k0 = k_begin; k0 < k_end; k0 += 2 * GBK) {
        store_tile(ra0, rb0);
        __syncthreads();
        if (k0 + 2 * GBK < k_end) load_tile(k0 + 2 * GBK, ra0, rb0);
        mul_tile();
        __syncthreads();
        if (k0 + GBK >= k_end) break;
        store_tile(ra1, rb1);
        __syncthreads();
        if (k0 + 3 * GBK < k_end) load_tile(k0 + 3 * GBK, ra1, rb1);
        mul_tile();
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= p.N) continue;
            float* c = C + gm * p.sCm + gn * p.sCn;
            float v = acc[i][j];
            if (p.splitk) {
                if (p.bias && blockIdx.z == 0) v += p.bias[gn];
                atomicAdd(c, v);
            } else {
                if (p.bias) v += p.bias[gn];
                if (p.act == 1) v = lrelu(v);
                else if (p.act == 2) v *= (p.mask_src[gm * p.sCm + gn * p.sCn] > 0.f ? 1.f : 0.2f);
                if (p.accumulate) v += *c;
                *c = v;
            }
        }
    }
}

int gemm_launch(const GemmP& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
    static int force_bm = -1;                    // LEMO_GEMM_BM=32|64 pins the tile (A/B measurements); default: fill heuristic
    if (force_bm < 0) { const char* e = getenv("LEMO_GEMM_BM"); force_bm = e ? atoi(e) : 0; }
    const int nz = p.nz > 0 ? p.nz : 1;
    const bool small = force_bm ? force_bm == 32 : (long long)cdiv(p.N, GBN) * cdiv(p.M, 64) * nz < 2 * 148;
    if (small) k_gemm<32><<<dim3(cdiv(p.N, GBN), cdiv(p.M, 32), nz), 128, 0, st>>>(p);
    else k_gemm<64><<<dim3(cdiv(p.N, GBN), cdiv(p.M, 64), nz), 256, 0, st>>>(p);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
