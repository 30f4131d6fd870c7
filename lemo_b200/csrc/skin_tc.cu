// Linear blend skinning on the tensor cores (reference lbs.py:106-117):
//     T[v, (b,k)] = sum_j W[v,j] . A[b,j,k]        (k = the 12 entries of the 3x4 transform)
//     verts[b,v]  = T[v,b,:3,:3] . v_posed[b,v] + T[v,b,:3,3] + transl[b]
// The reference materialises W.repeat(B) (274 MB at B=119) and T [B,V,4,4] (80 MB); the CUDA-core kernel k_skin_fwd spends 660 FMAs
// per (frame, vertex) and is FMA-bound (23 us floor for the dense synthetic weights, 43 us measured at B=120).  Here the per-vertex
// weighted-transform reduction is one GEMM  M = vertices (128 per tile), N = frames x 12 (96 = 8 frames per unit), K = 55 -> 56 joints,
// on tcgen05.mma kind::tf32 with the accumulator in TMEM, and the 3x4 apply is the epilogue: T never leaves the SM.
//
//   unit       (vertex tile of 128, chunk of 8 frames); persistent CTAs walk a contiguous range of units (vertex tile major), so the
//              64 KB weight tile stays in shared memory across the ~8 units a CTA owns
//   operands   both pre-split for a 3-term TF32 product (hi.hi + lo.hi + hi.lo, fp32-grade: residual 2^-21) and stored box by box,
//              so every TMA box (32 floats = one 128 B swizzle row, x 128 / 96 rows) is one contiguous run of memory:
//                W2 [vertex tile][hi k0-31 | hi k32-63 | lo k0-31 | lo k32-63][128][32]   rn_tf32(W) and W - hi, made once at model create
//                A2 [frame chunk][same four sub-tiles][96 = 8 frames x 12][32]             hi(A^T) and A^T - hi, written by k_pose_chain_fwd
//              joints 55..63 are zero padding; k-step 7 (joints 56..63) is never multiplied
//   warps      0: TMA producer   1: TMEM alloc + MMA issuer   2..9: epilogue (lane = vertex), two groups of four; a group copies its
//              accumulator to registers and releases it at once, so the MMAs of unit i+2 overlap the 3x4 apply of unit i
//   HBM        reads v_posed (4.B.3V) + W2 (5.4 MB, L2 resident), writes verts (4.B.3V)
// Measured on B200 (tools/diag_skin_tl.py, CTA-0 timeline): 14 us warm at B=120; a unit costs ~1.2 us of which 0.85 us is the 21 MMAs
// (78 cycles each: M=128 x N=96 x K=8 TF32 with both operands read from shared memory), the rest the hand-over to the next unit.
// Tried and dropped: A2 streamed as plain fp32 with A_lo = A - trunc(A) written next to each landed tile by two converter warps
// (the tensor core truncates an fp32 operand to TF32 itself; results identical, 9.5e-7 vs the CUDA-core kernel) -- it halves the
// re-streamed operand but the kernel is not L2-bandwidth bound, 20.0 vs 19.5 us under ncu.
#include "body.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

namespace lemo {

constexpr int SK_M = 128;                 // vertices per tile
constexpr int SK_FR = SKIN_TC_FR;         // frames per unit (8)
constexpr int SK_N = SK_FR * 12;          // 96 accumulator columns
constexpr int SK_KP = 64;                 // padded joints per split half (two 32-float swizzle atoms)
constexpr int SK_KSTEPS = 7;              // 7 x 8 = 56 >= 55 joints: the last k-step of the second atom is all padding
constexpr int SK_W_SUB = SK_M * 128;      // bytes of one 32-float-wide sub-tile of W2 (128 rows x 128 B)
constexpr int SK_A_SUB = SK_N * 128;      // 12 KB
constexpr int SK_W_BYTES = 4 * SK_W_SUB;  // hi atom0, hi atom1, lo atom0, lo atom1
constexpr int SK_A_BYTES = 4 * SK_A_SUB;  // 48 KB
constexpr int SK_STAGES = 3;              // A2 ring
constexpr int SK_ACC = 2;                 // TMEM accumulators (one per epilogue group)
constexpr int SK_ACC_COLS = 128;          // column pitch between accumulators
constexpr int SK_EPI_WARPS = 8;
constexpr int SK_STAGE_OUT = SK_EPI_WARPS * 4 * 96 * 4;   // per epilogue warp: 4 frames x 96 floats
constexpr size_t SK_SMEM = 1024 + SK_W_BYTES + SK_STAGES * SK_A_BYTES + SK_STAGE_OUT + 256;
static_assert(SK_A_SUB % 1024 == 0 && SK_N % 16 == 0, "swizzled sub-tiles are 1024 B aligned; UMMA N is a multiple of 16");

__device__ __forceinline__ uint32_t sk_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mb_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void sk_mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void sk_mb_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void sk_tma2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t sk_desc(uint32_t saddr) {           // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t sk_idesc(int M, int N) {        // D=F32, A=B=TF32, both K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void sk_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void sk_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sk_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// TL (debug instantiation, LEMO_SKIN_TL=1): CTA 0 records globaltimer stamps per unit and prints them at the end.
__device__ unsigned long long g_sk_tl[32 * 8];
__device__ __forceinline__ unsigned long long sk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define SK_STAMP(slot) do { if (TL && blockIdx.x == 0 && i < 32) g_sk_tl[i * 8 + (slot)] = sk_now() - tl0; } while (0)

// barriers: [0] W full, [1] W empty, [2..4] A full, [5..7] A empty, [8..9] accumulator full, [10..11] accumulator empty
template <bool TL>
__global__ void __launch_bounds__(64 + 32 * SK_EPI_WARPS, 1) k_skin_tc(const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_a2,
                                                    const float* __restrict__ VP, const float* __restrict__ transl, int V, int B,
                                                    int n_fc, int n_units, float* __restrict__ verts) {
    extern __shared__ uint8_t sk_smem_raw[];
    uint8_t* smem = sk_smem_raw + ((1024u - (sk_u32(sk_smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS)
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + SK_W_BYTES;
    uint8_t* s_stage = smem + SK_W_BYTES + SK_STAGES * SK_A_BYTES;
    uint64_t* bars = (uint64_t*)(s_stage + SK_STAGE_OUT);
    uint32_t* tmem_slot = (uint32_t*)(bars + 14);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long tl0 = TL ? sk_now() : 0ull;
    // contiguous, balanced range of units for this CTA; unit u = (vertex tile u / n_fc, frame chunk u % n_fc)
    const int u_begin = (int)((long long)blockIdx.x * n_units / gridDim.x);
    const int u_end = (int)((long long)(blockIdx.x + 1) * n_units / gridDim.x);

    if (warp == 0 && lane == 0) {
        sk_mb_init(sk_u32(&bars[0]), 1);
        sk_mb_init(sk_u32(&bars[1]), 1);
        for (int s = 0; s < SK_STAGES; ++s) {
            sk_mb_init(sk_u32(&bars[2 + s]), 1);
            sk_mb_init(sk_u32(&bars[5 + s]), 1);
        }
        for (int s = 0; s < SK_ACC; ++s) {
            sk_mb_init(sk_u32(&bars[8 + s]), 1);
            sk_mb_init(sk_u32(&bars[10 + s]), 4);          // one arrival per epilogue warp of the group
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a2) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sk_u32(tmem_slot)), "r"(SK_ACC * SK_ACC_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int cur_vt = -1, n_w = 0;
            for (int u = u_begin, i = 0; u < u_end; ++u, ++i) {
                const int vt = u / n_fc, fc = u - vt * n_fc;
                if (vt != cur_vt) {
                    // the weight tile changes: wait until every MMA that read the old one has retired
                    sk_mb_wait(sk_u32(&bars[1]), (n_w & 1) ^ 1);
                    const uint32_t full = sk_u32(&bars[0]);
                    sk_mb_expect_tx(full, SK_W_BYTES);
                    for (int h = 0; h < 4; ++h) sk_tma2d(sk_u32(s_w + h * SK_W_SUB), &map_w2, full, 0, (vt * 4 + h) * SK_M);
                    cur_vt = vt;
                    ++n_w;
                }
                const int s = i % SK_STAGES;
                sk_mb_wait(sk_u32(&bars[5 + s]), ((i / SK_STAGES) & 1) ^ 1);
                const uint32_t full = sk_u32(&bars[2 + s]);
                sk_mb_expect_tx(full, SK_A_BYTES);
                for (int h = 0; h < 4; ++h) sk_tma2d(sk_u32(s_a + s * SK_A_BYTES + h * SK_A_SUB), &map_a2, full, 0, (fc * 4 + h) * SK_N);
                SK_STAMP(0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = sk_idesc(SK_M, SK_N);
            int cur_vt = -1, n_w = 0;
            for (int u = u_begin, i = 0; u < u_end; ++u, ++i) {
                const int vt = u / n_fc;
                const bool last_of_tile = (u + 1 == u_end) || ((u + 1) / n_fc != vt);
                if (vt != cur_vt) {
                    sk_mb_wait(sk_u32(&bars[0]), n_w & 1);
                    cur_vt = vt;
                    ++n_w;
                }
                const int s = i % SK_STAGES, ac = i & 1;
                sk_mb_wait(sk_u32(&bars[2 + s]), (i / SK_STAGES) & 1);               // A2 chunk landed
                SK_STAMP(1);
                sk_mb_wait(sk_u32(&bars[10 + ac]), ((i >> 1) & 1) ^ 1);              // accumulator copied out by the epilogue
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                SK_STAMP(2);
                const uint32_t acc = tmem_base + (uint32_t)(ac * SK_ACC_COLS);
                const uint32_t wb = sk_u32(s_w), ab = sk_u32(s_a + s * SK_A_BYTES);
                uint32_t first = 0;
#pragma unroll
                for (int term = 0; term < 3; ++term) {                               // W_hi.A_hi, W_lo.A_hi, W_hi.A_lo
                    const uint32_t w0 = wb + (term == 1 ? 2 * SK_W_SUB : 0);
                    const uint32_t a0 = ab + (term == 2 ? 2 * SK_A_SUB : 0);
#pragma unroll
                    for (int ks = 0; ks < SK_KSTEPS; ++ks) {
                        const uint32_t atom = ks >> 2, off = (uint32_t)(ks & 3) * 32u;   // +32 B per UMMA_K = 8 floats inside the swizzle atom
                        sk_mma(acc, sk_desc(w0 + atom * SK_W_SUB + off), sk_desc(a0 + atom * SK_A_SUB + off), idesc, first);
                        first = 1;
                    }
                }
                sk_commit(sk_u32(&bars[5 + s]));                                     // A2 stage reusable when these MMAs retire
                sk_commit(sk_u32(&bars[8 + ac]));                                    // accumulator complete
                if (last_of_tile) sk_commit(sk_u32(&bars[1]));                       // weight tile reusable
                SK_STAMP(3);
            }
        }
    } else {
        // ===================== epilogue: lane = vertex, 12 columns per frame =====================
        // Two groups of four warps; group e owns accumulator e (units i with i & 1 == e).  [B,V,3] rows are 12 B per vertex: a warp's
        // 32 vertices of one frame are 96 contiguous floats, moved as three fully coalesced 128 B requests and transposed to/from the
        // per-lane (x,y,z) through a small staging buffer (the direct 12 B-stride form issued 3 partial writes per sector).
        const int ew = warp - 2, grp = ew >> 2, lq = warp & 3;
        float* stg = reinterpret_cast<float*>(s_stage) + ew * (4 * 96);
        for (int u = u_begin + grp, i = grp; u < u_end; u += 2, i += 2) {
            const int vt = u / n_fc, fc = u - vt * n_fc;
            const int v0 = vt * SK_M + lq * 32;                                      // first vertex of this warp
            const int b0 = fc * SK_FR;
            const int nval = min(96, 3 * (V - v0));                                  // valid floats of the warp's 96 (<= 0: none)
            // v_posed of the warp's vertices for the unit's frames: issued before the accumulator wait so the loads overlap the MMAs
            float q[SK_FR][3];
#pragma unroll
            for (int f = 0; f < SK_FR; ++f) {
                const float* src = VP + ((size_t)(b0 + f) * V + v0) * 3;
#pragma unroll
                for (int k = 0; k < 3; ++k) q[f][k] = ((b0 + f) < B && 32 * k + lane < nval) ? __ldg(src + 32 * k + lane) : 0.f;
            }
            if (lq == 0 && lane == 0) SK_STAMP(4);
            sk_mb_wait(sk_u32(&bars[8 + grp]), (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lq == 0 && lane == 0) SK_STAMP(5);
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(grp * SK_ACC_COLS);
            uint32_t r[SK_N];
#pragma unroll
            for (int c = 0; c < SK_N / 16; ++c) sk_tmem_ld16(taddr + 16 * c, r + 16 * c);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) sk_mb_arrive(sk_u32(&bars[10 + grp]));                    // the accumulator may be overwritten
            if (lq == 0 && lane == 0) SK_STAMP(6);
#pragma unroll
            for (int g = 0; g < SK_FR / 4; ++g) {                                    // 4 frames per pass through the staging buffer
#pragma unroll
                for (int ff = 0; ff < 4; ++ff)
#pragma unroll
                    for (int k = 0; k < 3; ++k) stg[ff * 96 + 32 * k + lane] = q[g * 4 + ff][k];
                __syncwarp();
                float o[4][3];
#pragma unroll
                for (int ff = 0; ff < 4; ++ff) {
                    const int b = b0 + g * 4 + ff;
                    const float p0 = stg[ff * 96 + 3 * lane], p1 = stg[ff * 96 + 3 * lane + 1], p2 = stg[ff * 96 + 3 * lane + 2];
                    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
                    if (transl && b < B) { t0 = __ldg(transl + b * 3); t1 = __ldg(transl + b * 3 + 1); t2 = __ldg(transl + b * 3 + 2); }
                    const uint32_t* T = r + (g * 4 + ff) * 12;
                    o[ff][0] = __uint_as_float(T[0]) * p0 + __uint_as_float(T[1]) * p1 + __uint_as_float(T[2]) * p2 + __uint_as_float(T[3]) + t0;
                    o[ff][1] = __uint_as_float(T[4]) * p0 + __uint_as_float(T[5]) * p1 + __uint_as_float(T[6]) * p2 + __uint_as_float(T[7]) + t1;
                    o[ff][2] = __uint_as_float(T[8]) * p0 + __uint_as_float(T[9]) * p1 + __uint_as_float(T[10]) * p2 + __uint_as_float(T[11]) + t2;
                }
                __syncwarp();
#pragma unroll
                for (int ff = 0; ff < 4; ++ff) {
                    stg[ff * 96 + 3 * lane] = o[ff][0]; stg[ff * 96 + 3 * lane + 1] = o[ff][1]; stg[ff * 96 + 3 * lane + 2] = o[ff][2];
                }
                __syncwarp();
#pragma unroll
                for (int ff = 0; ff < 4; ++ff) {
                    const int b = b0 + g * 4 + ff;
                    if (b < B) {
                        float* dst = verts + ((size_t)b * V + v0) * 3;
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            if (32 * k + lane < nval) dst[32 * k + lane] = stg[ff * 96 + 32 * k + lane];
                    }
                }
                __syncwarp();
            }
            if (lq == 0 && lane == 0) SK_STAMP(7);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (TL && blockIdx.x == 0 && threadIdx.x == 0) {
        printf("[sk_tl] CTA 0, units %d..%d, kernel end %llu ns\n[sk_tl] unit: A2 issued | A2 landed, acc free, MMAs issued | epi ready, acc full, acc copied, unit stored\n",
               u_begin, u_end, sk_now() - tl0);
        for (int i = 0; i < u_end - u_begin && i < 32; ++i)
            printf("[sk_tl] %2d: %6llu | %6llu %6llu %6llu | %6llu %6llu %6llu %6llu\n", i, g_sk_tl[i * 8], g_sk_tl[i * 8 + 1], g_sk_tl[i * 8 + 2],
                   g_sk_tl[i * 8 + 3], g_sk_tl[i * 8 + 4], g_sk_tl[i * 8 + 5], g_sk_tl[i * 8 + 6], g_sk_tl[i * 8 + 7]);
    }
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SK_ACC * SK_ACC_COLS) : "memory");
}

// W2, box by box: sub-tile h = 2*(lo half) + (j >> 5) of vertex tile vt holds [128 vertices][32 joints]; hi = rn_tf32(w), lo = w - hi;
// joints >= 55 and vertices >= V are zero
__global__ void k_skin_tc_prep_w(const float* __restrict__ w_jm, int V, int Vpad, float* __restrict__ W2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Vpad * SK_KP) return;
    const int v = i / SK_KP, j = i - v * SK_KP;
    float w = 0.f;
    if (v < V && j < NJ) w = w_jm[(size_t)j * V + v];
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(w));
    const float hi = __uint_as_float(t);
    const size_t tile = (size_t)(v / SK_M) * 4, row = v % SK_M;
    W2[((tile + (j >> 5)) * SK_M + row) * 32 + (j & 31)] = hi;
    W2[((tile + 2 + (j >> 5)) * SK_M + row) * 32 + (j & 31)] = w - hi;
}

typedef CUresult (*PFN_encodeTiledSk)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int sk_make_map(void* out_map, const float* base, long long rows, int box_rows) {
    static PFN_encodeTiledSk enc = nullptr;
    if (!enc) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            enc = (PFN_encodeTiledSk)p;
    }
    LEMO_CHECK(enc, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {32u, (cuuint64_t)rows};                  // box-by-box storage: rows of 32 floats
    const cuuint64_t gstr[1] = {32u * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LEMO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (skin_tc)");
    return 0;
}

int skin_tc_vpad(int V) { return cdiv(V, SK_M) * SK_M; }
int skin_tc_prep_w(const float* w_jm, int V, float* W2, void* map_w2) {
    const int Vpad = skin_tc_vpad(V);
    k_skin_tc_prep_w<<<cdiv(Vpad * SK_KP, 256), 256>>>(w_jm, V, Vpad, W2);
    LEMO_CUDA(cudaGetLastError());
    return sk_make_map(map_w2, W2, (long long)(Vpad / SK_M) * 4 * SK_M, SK_M);
}
size_t skin_tc_a2_floats(int maxB) { return (size_t)cdiv(maxB, SK_FR) * 4 * SK_N * 32; }
int skin_tc_map_a(const float* A2, int maxB, void* map_a2) { return sk_make_map(map_a2, A2, (long long)cdiv(maxB, SK_FR) * 4 * SK_N, SK_N); }

int skin_tc_launch(const void* map_w2, const void* map_a2, const float* VP, const float* transl, int V, int B, float* verts,
                   cudaStream_t st) {
    static int n_sm = 0;
    static bool tl = false;
    if (!n_sm) {
        const char* e = getenv("LEMO_SKIN_TL");
        tl = e && e[0] == '1';
        LEMO_CUDA(cudaFuncSetAttribute(k_skin_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM));
        LEMO_CUDA(cudaFuncSetAttribute(k_skin_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM));
        int dev = 0;
        LEMO_CUDA(cudaGetDevice(&dev));
        LEMO_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int n_fc = cdiv(B, SK_FR), n_units = cdiv(V, SK_M) * n_fc;
    const int grid = n_units < n_sm ? n_units : n_sm;
    if (tl) k_skin_tc<true><<<grid, 64 + 32 * SK_EPI_WARPS, SK_SMEM, st>>>(*(const CUtensorMap*)map_w2, *(const CUtensorMap*)map_a2, VP, transl, V, B, n_fc, n_units, verts);
    else k_skin_tc<false><<<grid, 64 + 32 * SK_EPI_WARPS, SK_SMEM, st>>>(*(const CUtensorMap*)map_w2, *(const CUtensorMap*)map_a2, VP, transl, V, B, n_fc, n_units, verts);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
