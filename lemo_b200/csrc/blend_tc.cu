// Blend-shape contraction on the 5th-generation tensor cores:  VP[B, 3V] = X[B,512] . Wt[512,3V]
// (shape + pose blend shapes of SMPL-X in one GEMM, reference lbs.py:81,94-99).  This is the ONE place the
// north star puts tensor cores: tcgen05.mma kind::tf32, fp32 accumulators in TMEM, both operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) through a 3-stage mbarrier pipeline.
//
//   CTA tile   128 (frames) x 224 (vertex coordinates), K = 512 in 16 blocks of 32 floats (= one 128 B swizzle row), 3 stages
//   grid       ceil(3V/224) x ceil(B/128)   -> 141 CTAs for the full mesh: one wave of the 148 SMs
//   warps      0: TMA producer   1: TMEM alloc + MMA issuer (one elected lane)   2-5: epilogue (TMEM -> smem -> coalesced stores)
//   operands   A = X2  [B,1024] K-major (hi | lo);  B = WtT [3V,512] K-major (transposed copy of Wt made at model-create time)
//   HBM        reads WtT once (64.3 MB) + X (L2 resident), writes VP (4*B*3V bytes): HBM-bound, 78.7 MB at B=120
// Precision: TF32 keeps 10 explicit mantissa bits.  X is split by k_pose_chain_fwd into X2 = [Xhi | Xlo] (Xhi = X with the low 13
// bits cleared, Xlo = X - Xhi, both exactly representable products of the split), and every Wt block is multiplied by both
// halves, so the only rounding left is the tensor core's conversion of the model constant Wt -- see DESIGN.md section 4.
#include "body.cuh"
#include <cuda.h>
#include <cstdlib>

namespace lemo {

constexpr int TC_BM = 128, TC_BN = 224, TC_BK = 32, TC_STAGES = 3, TC_UMMA_K = 8;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;      // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;      // 28 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + TC_B_BYTES;     // X_hi block, X_lo block, Wt block
constexpr int TC_TMEM_COLS = 256;
constexpr int TC_OUT_PITCH = TC_BN + 1;
constexpr size_t TC_SMEM = 1024 /*align slack*/ + (size_t)TC_STAGES * TC_STAGE_BYTES + 256 /*barriers*/;
static_assert(2 * (TC_STAGE_BYTES + TC_B_BYTES) <= TC_STAGES * TC_STAGE_BYTES, "the 3-term variant (2 stages of 88 KB) must fit in the same buffer");
static_assert((size_t)TC_BM * TC_OUT_PITCH * 4 <= (size_t)TC_STAGES * TC_STAGE_BYTES, "epilogue staging reuses the pipeline buffers");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// UMMA shared-memory descriptor: K-major tile, 128-byte swizzle, 8-row groups 1024 B apart (cute SmemDescriptor, sm_100 version 1)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address  [0,14)
    d |= (uint64_t)(1) << 16;                          // leading byte offset (unused for swizzled K-major), [16,30)
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;       // stride byte offset = 1024 B, [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor (cute InstrDescriptor): D=F32, A=B=TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// The same kernel is the generic TF32 GEMM of the library: C[M,N] = epi( [A_hi | A_lo][M, lo_col + K] . B[N,K]^T ), used for the VPoser
// MLP and its adjoint (vposer.cu) with bias / LeakyReLU / LeakyReLU' epilogues and an optional (hi|lo)-split copy of the result that
// feeds the next layer's A operand.
// three != 0: a second B tensor map_wlo = B - rn_tf32(B) is loaded next to every B block and a third MMA (A_hi . B_lo) is issued,
// i.e. the full 3-term split A_hi B_hi + A_lo B_hi + A_hi B_lo (fp32-grade; used where the result is not diluted by a larger term).
__global__ void __launch_bounds__(192, 1) k_blend_tf32(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                       const __grid_constant__ CUtensorMap map_wlo, float* __restrict__ VP, int B, int N, int K,
                                                       int lo_col, int three, int b_tiled, TcEpi ep) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024 B alignment.  Offsetting the __shared__ array (rather than rounding a uintptr_t) keeps the
    // address space known to the compiler: the epilogue staging then compiles to LDS/STS instead of generic LD/ST.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stage_bytes = three ? TC_STAGE_BYTES + TC_B_BYTES : TC_STAGE_BYTES;
    const int nstages = three ? 2 : TC_STAGES;                                         // 2 x 88 KB or 3 x 60 KB
    uint64_t* bars = (uint64_t*)(smem + (size_t)TC_STAGES * TC_STAGE_BYTES);
    // bars[0..S) full, bars[S..2S) empty, bars[2S] tmem_full, then the TMEM base address slot
    uint32_t* tmem_slot = (uint32_t*)(bars + 9);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int nkb = K / TC_BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[TC_STAGES + s]), 1); }
        mbar_init(smem_u32(&bars[2 * TC_STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % nstages;
                const uint32_t ph = (kb / nstages) & 1;
                mbar_wait(smem_u32(&bars[TC_STAGES + s]), ph ^ 1);                  // slot free (first pass returns immediately)
                const uint32_t full = smem_u32(&bars[s]);
                mbar_expect_tx(full, stage_bytes);
                const uint32_t a_dst = smem_u32(smem + (size_t)s * stage_bytes);
                if (three) tma_load_2d(a_dst + 2 * TC_A_BYTES + TC_B_BYTES, &map_wlo, full, kb * TC_BK, n0);
                tma_load_2d(a_dst, &map_x, full, kb * TC_BK, m0);                   // X_hi block
                tma_load_2d(a_dst + TC_A_BYTES, &map_x, full, lo_col + kb * TC_BK, m0);  // X_lo block
                // b_tiled: B was re-laid out as [n tile][k block][224 rows][32 floats], one contiguous 28 KB box per (tile, k block)
                if (b_tiled) tma_load_2d(a_dst + 2 * TC_A_BYTES, &map_w, full, 0, (blockIdx.x * nkb + kb) * TC_BN);
                else tma_load_2d(a_dst + 2 * TC_A_BYTES, &map_w, full, kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % nstages;
                const uint32_t ph = (kb / nstages) & 1;
                mbar_wait(smem_u32(&bars[s]), ph);                                  // TMA bytes have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t blo = umma_desc_sw128(a_addr + 2 * TC_A_BYTES + TC_B_BYTES);
                const uint64_t ahi = umma_desc_sw128(a_addr), alo = umma_desc_sw128(a_addr + TC_A_BYTES);
                const uint64_t bdesc = umma_desc_sw128(a_addr + 2 * TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    // advancing K inside the 128 B swizzle atom = +32 B on the start address (encoded >>4)
                    const uint64_t off = (uint64_t)(k * TC_UMMA_K * 4 >> 4);
                    umma_tf32(tmem_base, ahi + off, bdesc + off, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_tf32(tmem_base, alo + off, bdesc + off, idesc, 1u);
                    if (three) umma_tf32(tmem_base, ahi + off, blo + off, idesc, 1u);
                }
                umma_commit(smem_u32(&bars[TC_STAGES + s]));                        // frees the smem slot when these MMAs retire
            }
            umma_commit(smem_u32(&bars[2 * TC_STAGES]));                            // accumulator complete
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> smem (transpose) -> coalesced global stores =====================
        const int lq = warp & 3;                                                    // TMEM lane quarter this warp may access
        mbar_wait(smem_u32(&bars[2 * TC_STAGES]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* s_out = reinterpret_cast<float*>(smem);                              // pipeline buffers are idle now
        const int row = lq * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) s_out[row * TC_OUT_PITCH + c0 + j] = __uint_as_float(r[j]);
        }
        __syncwarp();
        const long long ldc = ep.ldc ? ep.ldc : N;
        if (ep.act == 0 && !ep.split_out && VP) {
            // plain / bias epilogue (the blend GEMM): bias hoisted into registers, 7 independent 128 B row stores per iteration
            // (the generic loop below re-loaded the bias and serialised LDS -> LDG -> STG per element: 2/3 of the kernel's 48 us)
            float bias_r[TC_BN / 32];
#pragma unroll
            for (int k = 0; k < TC_BN / 32; ++k) {
                const int gc = n0 + lane + 32 * k;
                bias_r[k] = (ep.bias && gc < N) ? __ldg(ep.bias + gc) : 0.f;
            }
            const int nrows = min(32, B - (m0 + lq * 32));
#pragma unroll 4
            for (int rr = 0; rr < nrows; ++rr) {
                const float* src = s_out + (lq * 32 + rr) * TC_OUT_PITCH + lane;
                float* dst = VP + (size_t)(m0 + lq * 32 + rr) * ldc + n0 + lane;
                float v[TC_BN / 32];
#pragma unroll
                for (int k = 0; k < TC_BN / 32; ++k) v[k] = src[32 * k] + bias_r[k];
#pragma unroll
                for (int k = 0; k < TC_BN / 32; ++k)
                    if (n0 + lane + 32 * k < N) dst[32 * k] = v[k];
            }
        } else {
            for (int rr = 0; rr < 32; ++rr) {                                       // this warp's 32 rows, lanes along columns
                const int gr = m0 + lq * 32 + rr;
                if (gr >= B) break;
                const float* src = s_out + (lq * 32 + rr) * TC_OUT_PITCH;
                for (int c = lane; c < TC_BN; c += 32) {
                    const int gc = n0 + c;
                    if (gc >= N) continue;
                    float v = src[c];
                    if (ep.bias) v += ep.bias[gc];
                    if (ep.act == 1) v = v > 0.f ? v : 0.2f * v;
                    else if (ep.act == 2) v *= ep.mask_src[(size_t)gr * ldc + gc] > 0.f ? 1.f : 0.2f;
                    if (VP) VP[(size_t)gr * ldc + gc] = v;
                    if (ep.split_out) {               // (hi|lo) TF32 split of the result = A operand of the next GEMM
                        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                        ep.split_out[(size_t)gr * ep.split_ld + gc] = hi;
                        ep.split_out[(size_t)gr * ep.split_ld + ep.split_lo + gc] = v - hi;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
}

// ------------------------------------------------------------------------------------------------ blend GEMM, version 2
// The blend-shape contraction proper (k_blend_tf32 above stays the generic TF32 GEMM).  Two changes, both aimed at bytes in flight,
// since the kernel is HBM-bound on the 64 MB of WtT and the first version spent 32 of every 60 KB stage on the (L2-resident) X tile:
//   * X_lo is multiplied only where it matters.  X_hi = rn_tf32(X) (k_pose_chain_fwd), and for the 486 pose-feature columns the
//     residual X_lo . W is 2^-12 of blend offsets that are themselves ~1 % of the body size: 5.0e-6 vs 3.9e-6 of max|v| with / without
//     it on the synthetic model (W's own TF32 rounding dominates either way).  Only k-blocks >= lo_from_kb (the block holding betas and
//     expression, whose offsets are 10x larger) get the second pass, as extra pipeline steps that re-read their W box from L2.
//   * stages shrink to 16 KB (X) + 28 KB (W) = 44 KB, so FIVE fit in shared memory: 140 KB of W in flight per SM instead of 84 KB, and
//     the tensor pipe does 17 instead of 32 MMA groups per tile.
constexpr int B2_STAGES = 5;
constexpr int B2_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr size_t B2_SMEM = 1024 + (size_t)B2_STAGES * B2_STAGE_BYTES + 256;
static_assert((size_t)TC_BM * TC_OUT_PITCH * 4 <= (size_t)B2_STAGES * B2_STAGE_BYTES, "epilogue staging reuses the pipeline buffers");

__global__ void __launch_bounds__(192, 1) k_blend_v2(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                     float* __restrict__ VP, int B, int N, int nkb, int lo_from_kb, int lo_col,
                                                     const float* __restrict__ bias) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = (uint64_t*)(smem + (size_t)B2_STAGES * B2_STAGE_BYTES);        // [0,5) full, [5,10) empty, [10] accumulator complete
    uint32_t* tmem_slot = (uint32_t*)(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int nsteps = nkb + (nkb - lo_from_kb);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < B2_STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[B2_STAGES + s]), 1); }
        mbar_init(smem_u32(&bars[2 * B2_STAGES]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int st = 0; st < nsteps; ++st) {
                const bool lo = st >= nkb;
                const int kb = lo ? lo_from_kb + (st - nkb) : st;
                const int s = st % B2_STAGES;
                mbar_wait(smem_u32(&bars[B2_STAGES + s]), ((st / B2_STAGES) & 1) ^ 1);
                const uint32_t full = smem_u32(&bars[s]);
                mbar_expect_tx(full, B2_STAGE_BYTES);
                const uint32_t dst = smem_u32(smem + (size_t)s * B2_STAGE_BYTES);
                tma_load_2d(dst + TC_A_BYTES, &map_w, full, 0, (blockIdx.x * nkb + kb) * TC_BN);      // contiguous 28 KB box of WtT
                tma_load_2d(dst, &map_x, full, (lo ? lo_col : 0) + kb * TC_BK, m0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
            for (int st = 0; st < nsteps; ++st) {
                const int s = st % B2_STAGES;
                mbar_wait(smem_u32(&bars[s]), (st / B2_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + (size_t)s * B2_STAGE_BYTES);
                const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + TC_A_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    const uint64_t off = (uint64_t)(k * TC_UMMA_K * 4 >> 4);
                    umma_tf32(tmem_base, adesc + off, bdesc + off, idesc, (st | k) != 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&bars[B2_STAGES + s]));
            }
            umma_commit(smem_u32(&bars[2 * B2_STAGES]));
        }
    } else {
        const int lq = warp & 3;
        // bias (v_template) for this lane's 7 columns: fetched while the main loop runs
        float bias_r[TC_BN / 32];
#pragma unroll
        for (int k = 0; k < TC_BN / 32; ++k) {
            const int gc = n0 + lane + 32 * k;
            bias_r[k] = (bias && gc < N) ? __ldg(bias + gc) : 0.f;
        }
        mbar_wait(smem_u32(&bars[2 * B2_STAGES]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* s_out = reinterpret_cast<float*>(smem);
        const int row = lq * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
                "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) s_out[row * TC_OUT_PITCH + c0 + j] = __uint_as_float(r[j]);
        }
        __syncwarp();
        const int nrows = min(32, B - (m0 + lq * 32));
#pragma unroll 4
        for (int rr = 0; rr < nrows; ++rr) {
            const float* src = s_out + (lq * 32 + rr) * TC_OUT_PITCH + lane;
            float* dst = VP + (size_t)(m0 + lq * 32 + rr) * N + n0 + lane;
            float v[TC_BN / 32];
#pragma unroll
            for (int k = 0; k < TC_BN / 32; ++k) v[k] = src[32 * k] + bias_r[k];
#pragma unroll
            for (int k = 0; k < TC_BN / 32; ++k)
                if (n0 + lane + 32 * k < N) dst[32 * k] = v[k];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
}

// (A K-chunked variant of this GEMM -- 128 x 64 tiles, one TMEM accumulator per K/8 chunk summed on the CUDA cores, for the VPoser MLP --
// was written in round 1 and MEASURED in round 2 (tools/diag_vposer_modes.py on B200): R_body error vs an fp64 oracle 5.3e-5 at B = 960
// against 2.8e-5 for the fp32 CUDA-core GEMMs and 3.6e-4 for the plain TF32 kernel, decode + adjoint time within 6 % of the fp32 path.
// Twice the error for no gain: removed, the VPoser GEMMs stay on gemm.cu's cluster split-K SGEMM.)

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// 2-D fp32 tensor [rows][512] (K contiguous), box = 32 floats x box_rows, 128-byte swizzle, OOB rows read as zero
int make_kmajor_map(void* out_map /*CUtensorMap, 128 B*/, const float* base, long long rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    LEMO_CHECK(enc, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LEMO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    return 0;
}

int blend_tc_map_x(const float* X2, int maxB, void* map_x) { return make_kmajor_map(map_x, X2, maxB, 2 * XK, TC_BM); }
// WtT is stored box by box (k_transpose_wt): a [n_tiles * 16 * 224][32] tensor whose boxes are contiguous 28 KB runs of HBM.
// (The plain [3V][512] layout made every box 224 separate 128 B pieces at a 2 KB stride: 47 us for the 64 MB read, 1.4 TB/s.)
int blend_tc_wtt_floats(int N) { return cdiv(N, TC_BN) * TC_BN * XK; }
int blend_tc_map_w(const float* WtT, int N, void* map_w) {
    return make_kmajor_map(map_w, WtT, (long long)cdiv(N, TC_BN) * (XK / TC_BK) * TC_BN, TC_BK, TC_BN);
}

int tc_gemm_launch(const void* map_a, const void* map_b, float* C, int M, int N, int K, int lo_col, const TcEpi& ep, cudaStream_t st,
                   const void* map_b_lo) {
    static bool configured = false;
    if (!configured) {
        LEMO_CUDA(cudaFuncSetAttribute(k_blend_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        configured = true;
    }
    LEMO_CHECK(K % TC_BK == 0 && K > 0, "tc_gemm: K must be a positive multiple of 32");
    dim3 grid(cdiv(N, TC_BN), cdiv(M, TC_BM));
    k_blend_tf32<<<grid, 192, TC_SMEM, st>>>(*(const CUtensorMap*)map_a, *(const CUtensorMap*)map_b,
                                             *(const CUtensorMap*)(map_b_lo ? map_b_lo : map_b), C, M, N, K, lo_col, map_b_lo ? 1 : 0,
                                             ep.b_tiled, ep);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
// LEMO_BLEND_V=1 selects the first kernel (X_lo on every k-block, 3 stages) for A/B measurements
static int blend_v2_launch(const void* map_x, const void* map_w, float* VP, int B, int N, const float* bias, cudaStream_t st) {
    static int ver = -1;
    if (ver < 0) {
        const char* e = getenv("LEMO_BLEND_V");
        ver = (e && e[0] == '1') ? 1 : 2;
        LEMO_CUDA(cudaFuncSetAttribute(k_blend_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B2_SMEM));
    }
    if (ver == 1) {
        TcEpi ep;
        ep.bias = bias;
        ep.b_tiled = 1;
        return tc_gemm_launch(map_x, map_w, VP, B, N, XK, XK, ep, st, nullptr);
    }
    // the last k-block (columns 480..511) holds the 6 last pose features, betas, expression and the zero padding
    k_blend_v2<<<dim3(cdiv(N, TC_BN), cdiv(B, TC_BM)), 192, B2_SMEM, st>>>(*(const CUtensorMap*)map_x, *(const CUtensorMap*)map_w, VP, B, N,
                                                                           XK / TC_BK, XK / TC_BK - 1, XK, bias);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int blend_tc_launch(const void* map_x, const void* map_w, float* VP, int B, int N, cudaStream_t st) {
    return blend_v2_launch(map_x, map_w, VP, B, N, nullptr, st);
}
// same GEMM with a per-column bias: bias = v_template gives v_posed directly (the tcgen05 skinning kernel reads it as is)
int blend_tc_launch_bias(const void* map_x, const void* map_w, float* VP, int B, int N, const float* bias, cudaStream_t st) {
    return blend_v2_launch(map_x, map_w, VP, B, N, bias, st);
}
int tc_map_a(void* map, const float* base, long long rows, int cols) { return make_kmajor_map(map, base, rows, cols, TC_BM); }
int tc_map_b(void* map, const float* base, long long rows, int cols) { return make_kmajor_map(map, base, rows, cols, TC_BN); }

// dst[n][kpad] = rn_tf32(src) with optional transpose: src is [rows_src][cols_src] row-major; transpose=0: dst[n=r][k=c]; 1: dst[n=c][k=r]
__global__ void k_tc_prep_b(const float* __restrict__ src, int rows_src, int cols_src, int transpose, int kpad, float* __restrict__ dst,
                            float* __restrict__ dst_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_src * cols_src) return;
    const int r = i / cols_src, c = i - r * cols_src;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(src[i]));
    const float v = __uint_as_float(t);
    const size_t o = !transpose ? (size_t)r * kpad + c : (size_t)c * kpad + r;
    dst[o] = v;
    if (dst_lo) dst_lo[o] = src[i] - v;
}
int tc_prep_b(const float* src, int rows_src, int cols_src, int transpose, int kpad, float* dst, cudaStream_t st, float* dst_lo) {
    k_tc_prep_b<<<cdiv(rows_src * cols_src, 256), 256, 0, st>>>(src, rows_src, cols_src, transpose, kpad, dst, dst_lo);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// WtT[c][p] = rn_tf32(Wt[p][c]): the model constant is rounded to TF32 ONCE, to nearest, at model-create time.  The tensor core
// would otherwise truncate it on every MMA (a biased error: measured 5.5e-5 of max|v| on the full mesh); pre-rounded values
// convert exactly, so the remaining error is unbiased and averages out over the 506 terms of the contraction.
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__global__ void k_transpose_wt(const float* __restrict__ Wt, float* __restrict__ WtT, int N) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + threadIdx.x;
        t[i][threadIdx.x] = c < N ? Wt[(size_t)(p0 + i) * N + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        if (c < N) {              // box (n tile, k block) = 224 rows x 32 floats, contiguous; p0 is a multiple of 32 = one k block
            const size_t box = (size_t)(c / TC_BN) * (XK / TC_BK) + p0 / TC_BK;
            WtT[(box * TC_BN + c % TC_BN) * TC_BK + threadIdx.x] = rn_tf32(t[threadIdx.x][i]);
        }
    }
}
int blend_tc_transpose(const float* Wt, float* WtT, int N) {
    k_transpose_wt<<<dim3(cdiv(N, 32), XK / 32), dim3(32, 8)>>>(Wt, WtT, N);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
