// Generic strided fp32 GEMM (see gemm.cuh).  Plain CUDA-core code: this kernel is deliberately the
// "utility" path; the roofline kernels (fused LBS forward, conv3x3) have their own files.
#include "gemm.cuh"
#include <cstdlib>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace lemo {

constexpr int GBN = 64, GBK = 16;

// GBM x 64 tile, 4 x 4 outputs per thread, GBM * 4 threads.  GBM = 32 halves the tile for the small problems of the fit (M = S*T <= 960 rows):
// twice the CTAs, several resident per SM, so one CTA's barrier bubbles are covered by another's math.  The K order per output element
// is the same for both tiles, so results are bit-identical.
// CLUSTER (thread-block cluster along z = K slices, <= 8 CTAs): each CTA multiplies its slice of K, parks its 32x64 partial tile in shared
// memory, and after cluster.sync() every CTA reduces a quarter-row share of the tile across the cluster through distributed shared
// memory in slice order (deterministic), applies the full epilogue (bias / LeakyReLU / mask / accumulate) and stores.  The fit's GEMMs
// have M = S*T <= 960 rows: 240 CTAs x 4 warps left 1.7 warps per scheduler and the FMA pipe 29 % busy (ncu: short-scoreboard stalls on
// the shared-memory operands); slicing K four ways gives every scheduler ~7 warps without a second pass or atomics.
// Measured: the narrow GEMMs (N = 126 / 32) halve (28 -> 11-15 us), the 512 x 512 ones stay at 36-51 us, i.e. those are bound by
// per-SM throughput, not latency.  A 64 x 64 tile with 8 x 8 outputs per thread (251 registers, 8 warps per SM) was tried for them and
// is slower end to end (1.505 vs 1.479 ms per fitting step); the tensor-core route is closed by accuracy (vposer.cu).
template <int GBM, bool CLUSTER = false>
__global__ void __launch_bounds__(GBM * 4) k_gemm(GemmP p) {
    constexpr int NT = GBM * 4, NB = GBN * GBK / NT;
    __shared__ float As[GBK][GBM + 4];
    __shared__ float Bs[GBK][GBN + 4];
    __shared__ float red[CLUSTER ? GBM * GBN : 1];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const float* A = p.A;
    const float* B = p.B;
    float* C = p.C;
    int k_begin = 0, k_end = p.K;
    if (p.splitk || CLUSTER) {
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
    if (CLUSTER) {
        cg::cluster_group cluster = cg::this_cluster();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[(ty * 4 + i) * GBN + tx * 4 + j] = acc[i][j];
        cluster.sync();
        const int nz = p.nz, rank = (int)cluster.block_rank();
        // row r of the tile is finished by CTA r % nz; a thread handles 4 consecutive columns (float4 over DSMEM)
        for (int e = tid; e < GBM * (GBN / 4); e += NT) {
            const int r = e / (GBN / 4), c4 = (e - r * (GBN / 4)) * 4;
            if (r % nz != rank) continue;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int sl = 0; sl < nz; ++sl) {
                const float4 q = *reinterpret_cast<const float4*>(cluster.map_shared_rank(red, sl) + r * GBN + c4);
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
            const int gm = m0 + r;
            if (gm >= p.M) continue;
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gn = n0 + c4 + j;
                if (gn >= p.N) continue;
                float* c = C + gm * p.sCm + gn * p.sCn;
                float o = vv[j];
                if (p.bias) o += p.bias[gn];
                if (p.act == 1) o = lrelu(o);
                else if (p.act == 2) o *= (p.mask_src[gm * p.sCm + gn * p.sCn] > 0.f ? 1.f : 0.2f);
                if (p.accumulate) o += *c;
                *c = o;
            }
        }
        cluster.sync();                       // peers may still be reading this CTA's partial tile
        return;
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
            if (p.splitk == 2) {                 // slice z parks its partial in C[z][M][N] (row-major); the caller adds the slices in order
                C[((size_t)blockIdx.z * p.M + gm) * p.N + gn] = v;
            } else if (p.splitk) {
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
    static int use_cluster = -1;                 // LEMO_GEMM_CLUSTER=0 disables the cluster split-K path (A/B measurements)
    if (use_cluster < 0) { const char* e = getenv("LEMO_GEMM_CLUSTER"); use_cluster = (e && e[0] == '0') ? 0 : 1; }
    // unbatched GEMMs (and single-slice "split-K" accumulations) whose 32x64 tiles cannot fill the GPU: slice K across a cluster
    const long long tiles32 = (long long)cdiv(p.N, GBN) * cdiv(p.M, 32);
    if (use_cluster && nz == 1 && !force_bm && p.K >= 96 && tiles32 < 4 * 148) {
        GemmP q = p;
        if (p.splitk) { q.splitk = 0; q.accumulate = 1; }      // one atomic slice onto C == C += result
        int cz = p.K >= 512 ? 4 : 2;
        while (cz < 8 && tiles32 * cz < 2 * 148 && p.K / (cz * 2) >= 64) cz *= 2;
        q.nz = cz;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cdiv(p.N, GBN), cdiv(p.M, 32), cz);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = cz;
        cfg.attrs = at; cfg.numAttrs = 1;
        LEMO_CUDA(cudaLaunchKernelEx(&cfg, k_gemm<32, true>, q));
        return 0;
    }
    const bool small = force_bm ? force_bm == 32 : (long long)cdiv(p.N, GBN) * cdiv(p.M, 64) * nz < 2 * 148;
    if (small) k_gemm<32><<<dim3(cdiv(p.N, GBN), cdiv(p.M, 32), nz), 128, 0, st>>>(p);
    else k_gemm<64><<<dim3(cdiv(p.N, GBN), cdiv(p.M, 64), nz), 256, 0, st>>>(p);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
