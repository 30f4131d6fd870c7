// Enc conv3x3 stack on the 5th-generation tensor cores (tcgen05, kind::f16 with bf16 operands, fp32 accumulators in TMEM).
//
// Why tensor cores here although the north star names only the blend-shape GEMM: the 3x3 stack is 95 % of a fitting iteration and
// the CUDA-core kernel (conv.cu) sits at ~50 % of the fp32 FMA roof; SURVEY.md section 7 flags exactly this ("the restriction has to
// be revisited").  Parity is kept by a bf16 x 3 split: every fp32 value v is carried as the pair (hi, lo) = (bf16(v), bf16(v - hi))
// (the same 4 bytes as one float) and every product is evaluated as hi*hi + hi*lo + lo*hi with fp32 accumulation, i.e. 16+ bits
// of mantissa per operand.  Measured on the full-size Enc (oracle/README of DESIGN.md section 4): 1.0e-5 of max|z| after 10 layers and
// 4e-6 on the smoothness loss, against the 1e-4 bar (plain fp32: 6e-7).  LEMO_CONV=simt selects the fp32 CUDA-core path.
//
// Implicit GEMM per output tile of 128 consecutive (padded pitch-linear) pixels:
//     D[128 px, 64 oc] = sum over 9 taps  A_tap[128 px, 64 ic] . W_tap[64 oc, 64 ic]^T
//   * activations are NHWC: one pixel = 64 channels = 128 B = exactly one 128-byte-swizzle row, so A_tap is ONE TMA box
//     (cp.async.bulk.tensor.2d) at row offset (ky-1)*Wp + (kx-1); the zero border of the plane layout supplies the padding and TMA
//     zero-fills the one row that can fall before the tensor.
//   * the 9 x {hi,lo} weight tiles (144 KB) are loaded once per CTA and stay in shared memory; CTAs are persistent over tiles.
//   * warp 0 = TMA producer (2-stage A ring), warp 1 = MMA issuer (108 tcgen05.mma per tile), warps 2-5 = epilogue
//     (tcgen05.ld -> bias/LeakyReLU or LeakyReLU' mask -> bf16 hi/lo split -> 128 B row stores), accumulators double-buffered in TMEM
//     so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "conv.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace lemo {

constexpr int CT_M = 128, CT_C = 64, CT_STAGES = 2;
constexpr int CT_A_TILE = CT_M * CT_C * 2;                 // 16 KB  (bf16)
constexpr int CT_W_TILE = CT_C * CT_C * 2;                 //  8 KB
constexpr int CT_W_BYTES = 18 * CT_W_TILE;                 // 9 taps x {hi,lo}
constexpr int CT_STAGE_BYTES = 2 * CT_A_TILE;              // hi + lo
constexpr size_t CT_SMEM = 1024 + CT_W_BYTES + CT_STAGES * CT_STAGE_BYTES + 256;
// row-reuse variant (MODE 1/2): one TMA box of 136 rows per ky serves the three kx taps through descriptor start addresses that are
// 0/128/256 B into the box (one pixel = one 128 B swizzle row); 3x less L2->SMEM traffic and 3x fewer pipeline round trips.
constexpr int CT_M2 = 136;
constexpr int CT_A_TILE2 = CT_M2 * CT_C * 2;               // 17 KB
constexpr int CT_STAGE_BYTES2 = 2 * CT_A_TILE2;
constexpr size_t CT_SMEM2 = 1024 + CT_W_BYTES + CT_STAGES * CT_STAGE_BYTES2 + 256;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {       // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
// (hi, lo) bf16 split of one fp32 value; returns hi in the low 16 bits of *h, lo in *l (as raw bf16 bits)
__device__ __forceinline__ void split_bf16(float v, uint32_t& h, uint32_t& l) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    h = (uint32_t)__bfloat16_as_ushort(hi);
    l = (uint32_t)__bfloat16_as_ushort(lo);
}

// per-tile epilogue shared by the kernel variants: TMEM -> registers -> bias/LeakyReLU or LeakyReLU' mask -> (hi,lo) rows
// per-tile epilogue shared by the kernel variants: TMEM -> registers -> bias/LeakyReLU or LeakyReLU' mask -> (hi,lo) rows.
// NCH channels per thread starting at c0 (64: one warp per TMEM lane quarter; 32: two warps per quarter, each half the channels).
template <bool STACK = false, int NCH = 64>
__device__ __forceinline__ void conv_tc_epilogue(uint32_t tmem_base, int acc, uint32_t empty_bar, int lq, int lane, int n, int q0, int qend,
                                                 int W, int Wp, int PS, int epi, const float* __restrict__ bias,
                                                 const __nv_bfloat16* __restrict__ aux_hi, __nv_bfloat16* __restrict__ out_hi,
                                                 __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32, int c0 = 0) {
    uint32_t r[NCH];
    const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * (STACK ? 2 * CT_C : CT_C)) + (uint32_t)c0;
#pragma unroll
    for (int h = 0; h < NCH / 32; ++h) tmem_ld32(taddr + 32 * h, r + 32 * h);
    if (STACK) {                      // columns [64,128) hold the hi*lo product of the stacked-weights MMA: fold them in
        uint32_t t2[32];
#pragma unroll
        for (int h = 0; h < NCH / 32; ++h) {
            tmem_ld32(taddr + 64 + 32 * h, t2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) r[32 * h + j] = __float_as_uint(__uint_as_float(r[32 * h + j]) + __uint_as_float(t2[j]));
        }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0 && empty_bar) mb_arrive(empty_bar);     // this warp's share of the accumulator is free (0: caller arrives later)
    const int q = q0 + lq * 32 + lane;
    if (q < qend) {
        const int col = q % Wp;
        const bool interior = col >= 1 && col <= W;
        const size_t row = (size_t)n * PS + q;
        uint4* oh = reinterpret_cast<uint4*>(out_hi + row * CT_C + c0);
        uint4* ol = reinterpret_cast<uint4*>(out_lo + row * CT_C + c0);
        const uint4* ax = aux_hi ? reinterpret_cast<const uint4*>(aux_hi + row * CT_C + c0) : nullptr;
#pragma unroll
        for (int c8 = 0; c8 < NCH / 8; ++c8) {                              // 8 channels = 16 B of bf16 per step
            uint32_t hw[4], lw[4];
            uint4 a = make_uint4(0, 0, 0, 0);
            if (epi == 1) a = ax[c8];
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float v[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = c8 * 8 + p * 2 + e;
                    float x = __uint_as_float(r[c]);
                    if (epi == 0) { x += __ldg(bias + c0 + c); x = x > 0.f ? x : 0.2f * x; }
                    else {
                        const uint32_t hb = (aw[p] >> (16 * e)) & 0xFFFFu;       // bf16 bits of the forward activation
                        const bool pos = hb != 0u && (hb & 0x8000u) == 0u;
                        x *= pos ? 1.f : 0.2f;
                    }
                    v[e] = interior ? x : 0.f;
                    if (out_f32) out_f32[row * CT_C + c0 + c] = v[e];
                }
                uint32_t h0, l0, h1, l1;
                split_bf16(v[0], h0, l0);
                split_bf16(v[1], h1, l1);
                hw[p] = h0 | (h1 << 16);
                lw[p] = l0 | (l1 << 16);
            }
            oh[c8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            ol[c8] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
    }
}

// epi: 0 = bias + LeakyReLU (forward), 1 = multiply by LeakyReLU'(aux) (input gradient)
// MODE 1 (default): one TMA box of 136 rows per ky; the three kx taps are descriptor start addresses 0/128/256 B into it.  The 128 B
//   swizzle is a function of the ABSOLUTE shared-memory address, so a start address that is not 1024 B aligned needs NO base_offset
//   (measured on B200: base_offset = 0 reproduces the per-tap result bit for bit, base_offset = (addr >> 7) & 7 gives garbage).
// MODE 0: one TMA box per tap (9 per tile) -- kept as the A/B reference for that experiment (tools/diag_conv_modes.py).
template <int MODE>
__global__ void __launch_bounds__(192, 1) k_conv_tc(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                    const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
                                                    const __nv_bfloat16* __restrict__ aux_hi, __nv_bfloat16* __restrict__ out_hi,
                                                    __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32, int N, int H, int W,
                                                    int Wp, int PS, int epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + CT_W_BYTES;
    constexpr int STAGE_BYTES = MODE == 0 ? CT_STAGE_BYTES : CT_STAGE_BYTES2;
    constexpr int A_TILE = MODE == 0 ? CT_A_TILE : CT_A_TILE2;
    constexpr int NLOADS = MODE == 0 ? 9 : 3;
    uint64_t* bars = (uint64_t*)(s_a + CT_STAGES * STAGE_BYTES);
    // bars: [0,1] A full, [2,3] A empty, [4,5] tmem full, [6,7] tmem empty, [8] weights
    uint32_t* tmem_slot = (uint32_t*)(bars + 9);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tps = (H * Wp + CT_M - 1) / CT_M;               // tiles per sample
    const int ntiles = N * tps;
    const int qend = (H + 1) * Wp;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) mb_init(s_u32(&bars[i]), 1);
        mb_init(s_u32(&bars[4]), 1); mb_init(s_u32(&bars[5]), 1);
        mb_init(s_u32(&bars[6]), 4); mb_init(s_u32(&bars[7]), 4);
        mb_init(s_u32(&bars[8]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (lane == 0) {
            const uint32_t wbar = s_u32(&bars[8]);
            mb_expect_tx(wbar, CT_W_BYTES);
            for (int t = 0; t < 18; ++t) tma2d(s_u32(s_w + t * CT_W_TILE), &map_w, wbar, 0, t * CT_C);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
                const int row0 = n * PS + q0;
                for (int ld = 0; ld < NLOADS; ++ld, ++it) {
                    const int s = it & 1;
                    mb_wait(s_u32(&bars[2 + s]), ((it >> 1) & 1) ^ 1);
                    const uint32_t full = s_u32(&bars[s]);
                    mb_expect_tx(full, STAGE_BYTES);
                    const int row = MODE == 0 ? row0 + (ld / 3 - 1) * Wp + (ld % 3 - 1) : row0 + (ld - 1) * Wp - 1;
                    const uint32_t dst = s_u32(s_a + s * STAGE_BYTES);
                    tma2d(dst, &map_hi, full, 0, row);
                    tma2d(dst + A_TILE, &map_lo, full, 0, row);
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (lane == 0) {
            // D=F32, A=B=BF16, K-major both, N=64, M=128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CT_C >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
            mb_wait(s_u32(&bars[8]), 0);
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                const int acc = lt & 1;
                mb_wait(s_u32(&bars[6 + acc]), ((lt >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * CT_C;
                for (int ld = 0; ld < NLOADS; ++ld, ++it) {
                    const int s = it & 1;
                    mb_wait(s_u32(&bars[s]), (it >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = s_u32(s_a + s * STAGE_BYTES);
#pragma unroll
                    for (int kx = 0; kx < (MODE == 0 ? 1 : 3); ++kx) {
                        const int tap = MODE == 0 ? ld : ld * 3 + kx;
                        const uint32_t ah = a0 + kx * 128, al = a0 + A_TILE + kx * 128;   // +1 pixel = +1 swizzle row
                        const uint64_t ahi = desc_sw128(ah), alo = desc_sw128(al);
                        const uint64_t whi = desc_sw128(s_u32(s_w + (tap * 2) * CT_W_TILE)), wlo = desc_sw128(s_u32(s_w + (tap * 2 + 1) * CT_W_TILE));
#pragma unroll
                        for (int k = 0; k < CT_C / 16; ++k) {                  // UMMA_K = 16 bf16 = 32 B inside the swizzle atom
                            const uint64_t o = (uint64_t)(k * 32 >> 4);
                            mma_bf16(d, ahi + o, whi + o, idesc, (tap | k) != 0 ? 1u : 0u);
                            mma_bf16(d, ahi + o, wlo + o, idesc, 1u);
                            mma_bf16(d, alo + o, whi + o, idesc, 1u);
                        }
                    }
                    mma_commit(s_u32(&bars[2 + s]));                          // A stage reusable once these MMAs retire
                }
                mma_commit(s_u32(&bars[4 + acc]));                            // accumulator of this tile complete
            }
        }
    } else {
        // ============================== epilogue ==============================
        const int lq = warp & 3;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
            mb_wait(s_u32(&bars[4 + acc]), (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            conv_tc_epilogue(tmem_base, acc, s_u32(&bars[6 + acc]), lq, lane, n, q0, qend, W, Wp, PS, epi, bias, aux_hi, out_hi, out_lo, out_f32);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
}

// ------------------------------------------------------------------------------------------------ deeper-pipeline variant
// Same math as k_conv_tc<1>, different shared-memory budget: ncu on k_conv_tc<1> shows the tensor pipe only 30-39 % active with L2 at
// 35 % and DRAM at 12 % -- the kernel waits on TMA->MMA round trips because 144 KB of resident weights leave room for just two A
// stages.  Here the weights of ONE ky row (3 taps x {hi,lo} = 48 KB) travel with each step through a 2-slot ring (they stay L2
// resident: 144 KB per layer), which frees shared memory for THREE 34 KB A stages.
constexpr int CW_NA = 3, CW_NW = 2;
constexpr int CW_WSLOT = 6 * CT_W_TILE;                                       // 48 KB
constexpr size_t CW_SMEM = 1024 + CW_NW * CW_WSLOT + CW_NA * CT_STAGE_BYTES2 + 256;
// STACK: the MMA atom M128 x N64 x K16 reads 6 KB of shared memory for 32 cycles of math (128 B/clk => 48 cycles): operand-bandwidth
// bound.  Stacking [W_hi ; W_lo] (adjacent 64-row tiles = one 128-row K-major operand) turns the two products that share A_hi into ONE
// N=128 MMA whose halves [hi*hi | hi*lo] are summed in the epilogue: A_hi is read once instead of twice (14 KB instead of 18 KB per K step).
// NEPI epilogue warps (4: one per TMEM lane quarter; 8: two per quarter, 32 channels each -- with 4, every SM sub-partition runs ONE
// epilogue warp with no latency hiding and the ~800-instruction epilogue, not the MMAs, paces the tile).
template <bool STACK, int NEPI>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1) k_conv_tc_ws(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                       const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
                                                       const __nv_bfloat16* __restrict__ aux_hi, __nv_bfloat16* __restrict__ out_hi,
                                                       __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32, int N, int H, int W,
                                                       int Wp, int PS, int epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + CW_NW * CW_WSLOT;
    uint64_t* bars = (uint64_t*)(s_a + CW_NA * CT_STAGE_BYTES2);
    // bars: [0..2] step full (A stage + weight slot), [3..5] A empty, [6,7] W empty, [8,9] tmem full, [10,11] tmem empty
    uint32_t* tmem_slot = (uint32_t*)(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tps = (H * Wp + CT_M - 1) / CT_M;
    const int ntiles = N * tps;
    const int qend = (H + 1) * Wp;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 10; ++i) mb_init(s_u32(&bars[i]), 1);
        mb_init(s_u32(&bars[10]), NEPI); mb_init(s_u32(&bars[11]), NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(STACK ? 256 : 128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
                const int row0 = n * PS + q0;
                for (int ky = 0; ky < 3; ++ky, ++it) {
                    const int sa = it % CW_NA, sw = it % CW_NW;
                    mb_wait(s_u32(&bars[3 + sa]), ((it / CW_NA) & 1) ^ 1);
                    mb_wait(s_u32(&bars[6 + sw]), ((it / CW_NW) & 1) ^ 1);
                    const uint32_t full = s_u32(&bars[sa]);
                    mb_expect_tx(full, CT_STAGE_BYTES2 + CW_WSLOT);
                    const int row = row0 + (ky - 1) * Wp - 1;
                    const uint32_t dst = s_u32(s_a + sa * CT_STAGE_BYTES2);
                    tma2d(dst, &map_hi, full, 0, row);
                    tma2d(dst + CT_A_TILE2, &map_lo, full, 0, row);
                    const uint32_t wdst = s_u32(s_w + sw * CW_WSLOT);
                    for (int t = 0; t < 6; ++t) tma2d(wdst + t * CT_W_TILE, &map_w, full, 0, (ky * 6 + t) * CT_C);   // taps ky*3+kx, {hi,lo}
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CT_C >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
            constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * CT_C >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
                const int acc = lt & 1;
                mb_wait(s_u32(&bars[10 + acc]), ((lt >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * (STACK ? 2 * CT_C : CT_C);
                for (int ky = 0; ky < 3; ++ky, ++it) {
                    const int sa = it % CW_NA, sw = it % CW_NW;
                    mb_wait(s_u32(&bars[sa]), (it / CW_NA) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = s_u32(s_a + sa * CT_STAGE_BYTES2), w0 = s_u32(s_w + sw * CW_WSLOT);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint64_t ahi = desc_sw128(a0 + kx * 128), alo = desc_sw128(a0 + CT_A_TILE2 + kx * 128);
                        const uint64_t whi = desc_sw128(w0 + (kx * 2) * CT_W_TILE), wlo = desc_sw128(w0 + (kx * 2 + 1) * CT_W_TILE);
#pragma unroll
                        for (int k = 0; k < CT_C / 16; ++k) {
                            const uint64_t o = (uint64_t)(k * 32 >> 4);
                            if (STACK) {
                                mma_bf16(d, ahi + o, whi + o, idesc128, (ky | kx | k) != 0 ? 1u : 0u);   // [hi*W_hi | hi*W_lo], N = 128
                                mma_bf16(d, alo + o, whi + o, idesc, 1u);                               // lo*W_hi into columns [0,64)
                            } else {
                                mma_bf16(d, ahi + o, whi + o, idesc, (ky | kx | k) != 0 ? 1u : 0u);
                                mma_bf16(d, ahi + o, wlo + o, idesc, 1u);
                                mma_bf16(d, alo + o, whi + o, idesc, 1u);
                            }
                        }
                    }
                    mma_commit(s_u32(&bars[3 + sa]));
                    mma_commit(s_u32(&bars[6 + sw]));
                }
                mma_commit(s_u32(&bars[8 + acc]));
            }
        }
    } else {
        const int lq = warp & 3;
        const int c0 = NEPI == 8 ? ((warp - 2) >> 2) * 32 : 0;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
            mb_wait(s_u32(&bars[8 + acc]), (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            conv_tc_epilogue<STACK, NEPI == 8 ? 32 : 64>(tmem_base, acc, s_u32(&bars[10 + acc]), lq, lane, n, q0, qend, W, Wp, PS, epi, bias, aux_hi,
                                                          out_hi, out_lo, out_f32, c0);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(STACK ? 256 : 128) : "memory");
}

// ------------------------------------------------------------------------------------------------ pair variant
// ncu on k_conv_tc_ws<1,8> (profiles/r01_ncu_full_summary.md): tensor pipe 47 % active, operand reads 51 %, L2 45 % -- nothing saturated;
// the warp-stall samples sit in the EPILOGUE warps (only 13 % of their samples wait for an accumulator): the ~1000-instruction epilogue
// (runtime `epi` switch, per-channel bias LDGs, scalar F2F conversions, aux loads issued after the wait) paces the tile, the MMA warp
// waits for a free accumulator.  This variant attacks both ends:
//   * one 48 KB weight slot serves TWO consecutive 128-pixel tiles (per ky step: W(ky) + A(ky, tile 0), then A(ky, tile 1)); the four
//     stacked accumulators (2 tiles x 2 buffers x 128 columns) fill TMEM's 512 columns exactly; L2->SMEM traffic per tile drops from
//     246 KB to 174 KB;
//   * EPI / F32OUT are template parameters, the bias lives in registers, the forward activation bits (aux) are fetched BEFORE the wait
//     on the accumulator, the (hi,lo) split uses the packed F2FP conversion (2 values per instruction) and every accumulator half has its
//     own full/empty barrier so the epilogue of tile 0 starts while tile 1 still runs.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// two fp32 values -> packed bf16 hi pair and packed bf16 lo pair (bit-identical to two split_bf16 calls)
__device__ __forceinline__ void split_bf16x2(float v0, float v1, uint32_t& hw, uint32_t& lw) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
    hw = *reinterpret_cast<const uint32_t*>(&h2);
    const float h0 = __uint_as_float(hw << 16), h1 = __uint_as_float(hw & 0xFFFF0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - h0, v1 - h1);
    lw = *reinterpret_cast<const uint32_t*>(&l2);
}

template <int NEPI, int EPI, bool F32OUT>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1) k_conv_tc_pair(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                         const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
                                                         const __nv_bfloat16* __restrict__ aux_hi, __nv_bfloat16* __restrict__ out_hi,
                                                         __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32, int N, int H, int W,
                                                         int Wp, int PS, int dbg, int kmax) {
    constexpr int NCH = 256 / NEPI;                       // channels per epilogue thread: 32 (8 warps) or 16 (16 warps)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;
    uint8_t* s_a = smem + CW_NW * CW_WSLOT;
    uint64_t* bars = (uint64_t*)(s_a + CW_NA * CT_STAGE_BYTES2);
    // bars: [0..2] A stage full (+ weight slot on the first half), [3..5] A empty, [6,7] W empty, [8..11] accumulator full (buffer*2 + half),
    //       [12..15] accumulator empty
    uint32_t* tmem_slot = (uint32_t*)(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tps = (H * Wp + CT_M - 1) / CT_M;
    const int ntiles = N * tps;
    const int npairs = (ntiles + 1) >> 1;
    const int qend = (H + 1) * Wp;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 12; ++i) mb_init(s_u32(&bars[i]), 1);
        for (int i = 12; i < 16; ++i) mb_init(s_u32(&bars[i]), NEPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Warps 0 and 1 run their loops CONVERGED (all 32 lanes wait on the barriers and compute the same addresses) and elect one lane
    // only around the TMA / tcgen05 instructions: descriptors then live in uniform registers and every UTCHMMA issues back to back.
    // With the whole loop under `if (lane == 0)` the compiler wrapped EACH UTCHMMA in an ELECT/R2UR/BRA.U.ANY "waterfall" loop --
    // ~90 cycles of issue per MMA, which (not the tensor pipe, not shared memory, not L2) paced the earlier variants
    // (tools/diag_conv_modes.py --experiments: time independent of the MMA shapes issued, halved by halving the MMA count).
    if (warp == 0) {
        uint32_t ia = 0, iw = 0;
        for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
            const int nh = 2 * pr + 1 < ntiles ? 2 : 1;
            for (int ky = 0; ky < 3; ++ky, ++iw) {
                const int sw = iw % CW_NW;
                for (int h = 0; h < nh; ++h, ++ia) {
                    const int tile = 2 * pr + h;
                    const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
                    const int sa = ia % CW_NA;
                    mb_wait(s_u32(&bars[3 + sa]), ((ia / CW_NA) & 1) ^ 1);
                    if (h == 0) mb_wait(s_u32(&bars[6 + sw]), ((iw / CW_NW) & 1) ^ 1);
                    const uint32_t full = s_u32(&bars[sa]);
                    const int row = n * PS + q0 + (ky - 1) * Wp - 1;
                    const uint32_t dst = s_u32(s_a + sa * CT_STAGE_BYTES2);
                    const uint32_t wdst = s_u32(s_w + sw * CW_WSLOT);
                    if (elect_one()) {
                        if (dbg & 1) mb_arrive(full);                    // timing experiment: no loads
                        else {
                            mb_expect_tx(full, h == 0 ? CT_STAGE_BYTES2 + CW_WSLOT : CT_STAGE_BYTES2);
                            tma2d(dst, &map_hi, full, 0, row);
                            tma2d(dst + CT_A_TILE2, &map_lo, full, 0, row);
                            if (h == 0)
                                for (int t = 0; t < 6; ++t) tma2d(wdst + t * CT_W_TILE, &map_w, full, 0, (ky * 6 + t) * CT_C);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CT_C >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
        constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * CT_C >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
        uint32_t ia = 0, iw = 0, lt = 0;
        for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x, ++lt) {
            const int nh = 2 * pr + 1 < ntiles ? 2 : 1;
            const int buf = lt & 1;
            for (int ky = 0; ky < 3; ++ky, ++iw) {
                const int sw = iw % CW_NW;
                const uint32_t w0 = s_u32(s_w + sw * CW_WSLOT);
                for (int h = 0; h < nh; ++h, ++ia) {
                    const int sa = ia % CW_NA;
                    const uint32_t d = tmem_base + (uint32_t)((buf * 2 + h) * 2 * CT_C);
                    if (ky == 0) mb_wait(s_u32(&bars[12 + buf * 2 + h]), ((lt >> 1) & 1) ^ 1);
                    mb_wait(s_u32(&bars[sa]), (ia / CW_NA) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = s_u32(s_a + sa * CT_STAGE_BYTES2);
                    if (elect_one()) {
                        const int reps = (dbg & 128) ? 2 : (dbg & 256) ? 4 : 1;      // timing experiment: repeat the step's MMAs
                        if (!(dbg & 2)) for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint64_t ahi = desc_sw128(a0 + kx * 128), alo = desc_sw128(a0 + CT_A_TILE2 + kx * 128);
                                const uint64_t whi = desc_sw128(w0 + (kx * 2) * CT_W_TILE);
#pragma unroll
                                for (int k = 0; k < CT_C / 16; ++k) {
                                    if (k >= kmax) continue;             // input channels >= 16 kmax are structurally zero (32-channel layers)
                                    const uint64_t o = (uint64_t)(k * 32 >> 4);
                                    if (dbg >= 8) {                       // timing experiments on the tensor pipe (tools/diag_conv_modes.py)
                                        constexpr uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(CT_M >> 4) << 24);
                                        const uint64_t a1 = (dbg & 16) ? desc_sw128(a0) + o : ahi + o;
                                        // 512: rotate the destination over 4 accumulator blocks (independent accumulate chains)
                                        const uint32_t dd = (dbg & 512) ? tmem_base + (uint32_t)(((kx + k) & 3) * 128) : d;
                                        if (dbg & 32) mma_bf16(tmem_base + (uint32_t)(((dbg & 512) ? (k & 1) : buf) * 256), a1, desc_sw128(w0) + o, idesc256, 1u);
                                        else {
                                            if (!(dbg & 64)) mma_bf16(dd, a1, whi + o, idesc128, 1u);
                                            if (!(dbg & 8)) mma_bf16((dbg & 512) ? dd + 64 : dd, (dbg & 16) ? desc_sw128(a0 + CT_A_TILE2) + o : alo + o, whi + o, idesc, 1u);
                                        }
                                        continue;
                                    }
                                    mma_bf16(d, ahi + o, whi + o, idesc128, (ky | kx | k) != 0 ? 1u : 0u);   // [hi*W_hi | hi*W_lo], N = 128
                                    mma_bf16(d, alo + o, whi + o, idesc, 1u);                               // lo*W_hi into columns [0,64)
                                }
                            }
                        }
                        mma_commit(s_u32(&bars[3 + sa]));
                        if (ky == 2) mma_commit(s_u32(&bars[8 + buf * 2 + h]));
                        if (h == nh - 1) mma_commit(s_u32(&bars[6 + sw]));
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        const int lq = warp & 3;
        const int c0 = ((warp - 2) >> 2) * NCH;
        float breg[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) breg[c] = EPI == 0 ? __ldg(bias + c0 + c) : 0.f;
        uint32_t lt = 0;
        for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x, ++lt) {
            const int nh = 2 * pr + 1 < ntiles ? 2 : 1;
            const int buf = lt & 1;
            for (int h = 0; h < nh; ++h) {
                const int tile = 2 * pr + h;
                const int n = tile / tps, q0 = Wp + (tile - n * tps) * CT_M;
                const int q = q0 + lq * 32 + lane;
                const bool live = q < qend;
                const int col = q % Wp;
                const bool interior = live && col >= 1 && col <= W;
                const size_t row = (size_t)n * PS + q;
                uint4 ax[NCH / 8];
                if (EPI == 1) {                               // forward activation bits: issued before the wait, consumed after it
                    const uint4* axp = reinterpret_cast<const uint4*>(aux_hi + row * CT_C + c0);
#pragma unroll
                    for (int i = 0; i < NCH / 8; ++i) ax[i] = live ? __ldg(axp + i) : make_uint4(0, 0, 0, 0);
                }
                mb_wait(s_u32(&bars[8 + buf * 2 + h]), (lt >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t r[NCH], t2[NCH];
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)((buf * 2 + h) * 2 * CT_C + c0);
                if (NCH == 32) { tmem_ld32(taddr, r); tmem_ld32(taddr + 64, t2); }
                else { tmem_ld16(taddr, r); tmem_ld16(taddr + 64, t2); }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mb_arrive(s_u32(&bars[12 + buf * 2 + h]));     // this warp's share of the accumulator is free
                if (!live || (dbg & 4)) continue;
                uint4* oh = reinterpret_cast<uint4*>(out_hi + row * CT_C + c0);
                uint4* ol = reinterpret_cast<uint4*>(out_lo + row * CT_C + c0);
#pragma unroll
                for (int c8 = 0; c8 < NCH / 8; ++c8) {
                    uint32_t hw[4], lw[4];
                    float v[8];
                    const uint32_t aw[4] = {ax[c8].x, ax[c8].y, ax[c8].z, ax[c8].w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int c = c8 * 8 + e;
                        float x = __uint_as_float(r[c]) + __uint_as_float(t2[c]);
                        if (EPI == 0) { x += breg[c]; x = fmaxf(x, 0.2f * x); }
                        else {
                            const uint32_t hb = (aw[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;    // bf16 bits of the forward activation
                            x *= (hb - 1u < 0x7FFFu) ? 1.f : 0.2f;                            // positive and non-zero
                        }
                        v[e] = interior ? x : 0.f;
                    }
                    if (F32OUT) {
                        float4* of = reinterpret_cast<float4*>(out_f32 + row * CT_C + c0 + c8 * 8);
                        of[0] = make_float4(v[0], v[1], v[2], v[3]);
                        of[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
#pragma unroll
                    for (int p2 = 0; p2 < 4; ++p2) split_bf16x2(v[2 * p2], v[2 * p2 + 1], hw[p2], lw[p2]);
                    oh[c8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    ol[c8] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ weights-in-TMEM variant
// ncu on k_conv_tc_pair (profiles/r01_ncu_full_summary.md): tensor pipe 42 %, tensor-core shared-memory reads 49 %, L2 38 %, DRAM 17 % --
// nothing saturated.  The kernel is pipeline-depth bound: 96 KB of streamed weight slots leave three 34 KB activation stages, i.e. ~1.2 us
// of look-ahead against an L2 round trip + 82 KB of serialisation per step.  This variant removes the weights from shared memory
// altogether by transposing the implicit GEMM:
//     D^T[128 = {W_hi ; W_lo} x 64 oc, 96 px] += Wstack[128, 16 ic] . act^T[16 ic, 96 px]      (tcgen05.mma with A in TENSOR MEMORY)
//   * the A operand is the stacked weight matrix [W_hi ; W_lo] of all 9 taps: 128 lanes x 288 columns of TMEM (2 bf16 per 32-bit
//     column), written ONCE per CTA with tcgen05.st and resident for every tile of the persistent loop;
//   * the B operand is the activation box (pixel rows of 64 channels = one 128 B swizzle row, exactly the NHWC tile the other variants
//     use as A): N = 96 pixels per tile, the three kx taps are descriptor start addresses 0/128/256 B into one 104-row TMA box per ky;
//   * each K step issues the SAME weight columns against act_hi and act_lo: rows [0,64) accumulate W_hi.(a_hi + a_lo), rows [64,128)
//     W_lo.(a_hi + a_lo); the epilogue adds the two halves = the full (hi+lo) x (hi+lo) product (4 bf16 products, one more than the
//     3-term split of the other variants, at the full M=128 rate and with no shared-memory operand traffic for the weights);
//   * shared memory holds nothing but EIGHT 26 KB activation stages (almost three tiles of look-ahead); L2->SMEM traffic per output
//     pixel drops from 1.36 KB (pair kernel) to 0.83 KB;
//   * TMEM: accumulators at columns [0,96) and [128,224) (double buffered), weights at [224,512).
// Epilogue (8 warps, NO shared memory, no CTA-level barrier).  The stacked rows are interleaved so that the two halves of one output
// channel sit in the SAME warp: TMEM lane 32 q + l holds part (l >> 4) of channel 16 q + (l & 15).  A warp loads its 32 lanes x 48 pixel
// columns, exchanges with lane ^ 16 (one SHFL per pixel pair: each lane keeps the pixels of its own parity and receives the other half
// of them), applies bias + LeakyReLU or the LeakyReLU' mask, splits into bf16 (hi,lo) and stores 2-byte values: the 16 lanes of one
// parity write 32 contiguous bytes (one full sector) of a pixel row per instruction.  Measured on B200 (tools/diag_conv_wt.py): the
// first version of this kernel transposed through a 48 KB shared-memory buffer with two named barriers per tile; that epilogue alone
// took 2.2 us per tile against 1.5 us of MMAs and did not overlap with them (66 us per layer, "MMA only" 61 us, "epilogue only" 42 us).
constexpr int WT_N = 96;                                   // pixels per tile = MMA N
constexpr int WT_BOX = 104;                                // TMA box rows: 96 + 2 halo rows, rounded up to the 8-row swizzle atom
constexpr int WT_PLANE = WT_BOX * 128;                     // 13 KB
constexpr int WT_STAGE = 2 * WT_PLANE;                     // hi + lo
constexpr int WT_NST = 8;
constexpr int WT_ACC1 = 128, WT_W0 = 224;                  // TMEM columns
constexpr int WT_TL_MAX = 24;                              // timeline (debug) slots: tiles per CTA recorded
constexpr size_t WT_SMEM = 1024 + WT_NST * WT_STAGE + 256 + WT_TL_MAX * 8 * 8;

__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same, with the shared-memory descriptor given as its LOW word only: the high word of every K-major SWIZZLE_128B descriptor of this file is the
// constant 0x40004040 (SBO = 1024 B, version 1, layout 2), and (address >> 4) < 2^14 for any shared-memory address, so advancing a
// descriptor is a 32-bit add on the low word -- one uniform-datapath instruction per MMA instead of a 64-bit mask/shift/add chain.
__device__ __forceinline__ void mma_bf16_ts32(uint32_t tmem_d, uint32_t tmem_a, uint32_t bdesc_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 bd, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "r"(bdesc_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
          "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// wq: the layer's TMEM weight image (k_tc_prep_wt).
// KMAX = real input channels / 16 (2 for the 32-channel layers: the other K steps are structurally zero and never issued).
// NGRP epilogue groups of 8 warps: with 2, group g owns accumulator g and every other tile, so a group has TWO tile times for its tile.
// var (timing experiments, tools/diag_conv_wt.py; results invalid): bit 1 = no epilogue stores, bit 2 = no MMAs, bit 3 = no TMA loads,
// bit 4 = epilogue only hand-shakes (no TMEM load, no arithmetic), bit 5 = TMEM load but no arithmetic.
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// TL (timeline, debug builds of the kernel only): CTA 0 prints globaltimer stamps of its producer / MMA / epilogue warps (var bit 7).
template <int EPI, bool F32OUT, int KMAX, int NGRP, bool TL = false>
__global__ void __launch_bounds__(64 + 256 * NGRP, 1) k_conv_tc_wt(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                                                       const uint32_t* __restrict__ wq, const float* __restrict__ bias,
                                                       const __nv_bfloat16* __restrict__ aux_hi, __nv_bfloat16* __restrict__ out_hi,
                                                       __nv_bfloat16* __restrict__ out_lo, float* __restrict__ out_f32, int N, int H, int W,
                                                       int Wp, int PS, int var) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned long long tl0 = TL ? gtime_ns() : 0ull;
    const bool tl = TL && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
    // timeline slots (shared memory, printed once at the end so the stamps do not perturb the run): per local tile
    // [0] loads issued [1] accumulator free (MMA warp) [2] first stage ready [3] MMAs issued [4] accumulator complete (epilogue) [5] TMEM read [6] stored
    unsigned long long* s_tl = (unsigned long long*)(smem_raw + WT_SMEM - WT_TL_MAX * 8 * 8);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;
    uint64_t* bars = (uint64_t*)(smem + WT_NST * WT_STAGE);
    // bars: [0..NST) stage full, [NST..2 NST) stage empty, then accumulator full [2], accumulator empty [2], weights of tap row ky in TMEM [3]
    constexpr int B_EMPTY = WT_NST, B_AFULL = 2 * WT_NST, B_AEMPTY = 2 * WT_NST + 2, B_W = 2 * WT_NST + 4;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * WT_NST + 7);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tps = (H * Wp + WT_N - 1) / WT_N;               // tiles per sample
    const int ntiles = N * tps;
    const int qend = (H + 1) * Wp;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < B_AEMPTY; ++i) mb_init(s_u32(&bars[i]), 1);
        mb_init(s_u32(&bars[B_AEMPTY]), 8); mb_init(s_u32(&bars[B_AEMPTY + 1]), 8);
        for (int i = 0; i < 3; ++i) mb_init(s_u32(&bars[B_W + i]), 12);       // 3 taps x 4 lane quarters
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_lo) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const unsigned long long tl1 = TL ? gtime_ns() : 0ull;

    if (warp >= 2) {
        // stacked weights -> TMEM: lane 32 lq + l holds W_part[oc][tap][ic 0..63], part = l >> 4, oc = 16 lq + (l & 15), as 32 bf16 pairs per
        // tap (K ascending along the columns, even K in the low half of a column).  The 2 NGRP warps of a lane quarter take the taps round
        // robin and signal each tap row (ky) separately, so the first tile's ky = 0 MMAs start after one L2 round trip instead of after the
        // whole 144 KB image (measured: 6.8 us of the 60 us kernel when the CTA waited for all of it), and the producer never waits.
        // wq is the TMEM image prepared by k_tc_prep_wt: [tap][16-byte chunk i][TMEM lane r][4 words], so one warp-level load reads 512
        // contiguous bytes (with the row-major tiles every lane read its own 128 B line: 32 tag look-ups per instruction, 4.7 us per CTA).
        const int lq = warp & 3;
        for (int tap = (warp - 2) >> 2; tap < 9; tap += 2 * NGRP) {
            uint32_t wv[32];
            const uint4* src = reinterpret_cast<const uint4*>(wq) + (size_t)tap * 8 * 128 + lq * 32 + lane;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 v = __ldg(src + i * 128);
                wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
            }
            tmem_st32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(WT_W0 + tap * 32), wv);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mb_arrive(s_u32(&bars[B_W + tap / 3]));
        }
    }
    const unsigned long long tl2 = TL ? gtime_ns() : 0ull;

    if (warp == 0) {
        // ============================== TMA producer (converged warp, one elected lane issues) ==============================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n = tile / tps, q0 = Wp + (tile - n * tps) * WT_N;
            for (int ky = 0; ky < 3; ++ky, ++it) {
                const int s = it % WT_NST;
                mb_wait(s_u32(&bars[B_EMPTY + s]), ((it / WT_NST) & 1) ^ 1);
                const uint32_t full = s_u32(&bars[s]);
                const int row = n * PS + q0 + (ky - 1) * Wp - 1;
                const uint32_t dst = s_u32(s_a + s * WT_STAGE);
                if (elect_one()) {
                    if (var & 8) mb_arrive(full);
                    else {
                        mb_expect_tx(full, WT_STAGE);
                        tma2d(dst, &map_hi, full, 0, row);
                        tma2d(dst + WT_PLANE, &map_lo, full, 0, row);
                    }
                }
                __syncwarp();
            }
            if (TL && tl && it / 3 - 1 < WT_TL_MAX) s_tl[(it / 3 - 1) * 8 + 0] = gtime_ns() - tl0;
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // D = F32, A = B = BF16, both K-major, N = 96, M = 128
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(WT_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t wbase = tmem_base + (uint32_t)WT_W0;
        uint32_t it = 0, lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            const uint32_t d = tmem_base + (uint32_t)(acc * WT_ACC1);
            mb_wait(s_u32(&bars[B_AEMPTY + acc]), ((lt >> 1) & 1) ^ 1);        // epilogue has drained this accumulator
            const unsigned long long tm0 = TL ? gtime_ns() : 0ull;
            unsigned long long tm1 = 0ull;
            for (int ky = 0; ky < 3; ++ky, ++it) {
                const int s = it % WT_NST;
                mb_wait(s_u32(&bars[s]), (it / WT_NST) & 1);
                mb_wait(s_u32(&bars[B_W + ky]), 0);                             // weights of this tap row are in TMEM (only ever waits on the first tile)
                if (TL && ky == 0) tm1 = gtime_ns();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // low descriptor word of the stage's hi plane (start address >> 4, LBO field = 1); every operand of the step is this + a constant
                const uint32_t b0 = ((s_u32(s_a + s * WT_STAGE) >> 4) & 0x3FFFu) | (1u << 16);
                const uint32_t wa = wbase + (uint32_t)(ky * 96);
                if (elect_one()) {
                    if (!(var & 4)) {
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) {                    // UMMA_K = 16 bf16 = 32 B in the swizzle atom = 8 TMEM columns of A
                                // +1 pixel (kx) = +1 swizzle row = +128 B; +1 K step = +32 B; lo plane = +WT_PLANE
                                const uint32_t bh = b0 + (uint32_t)((kx * 128 + k * 32) >> 4);
                                const uint32_t bl = bh + (uint32_t)(WT_PLANE >> 4);
                                const uint32_t a = wa + (uint32_t)(kx * 32 + k * 8);
                                mma_bf16_ts32(d, a, bh, idesc, (kx | k) != 0 ? 1u : (ky != 0 ? 1u : 0u));
                                mma_bf16_ts32(d, a, bl, idesc, 1u);
                            }
                        }
                    }
                    mma_commit(s_u32(&bars[B_EMPTY + s]));                     // stage reusable once these MMAs retire
                    if (ky == 2) mma_commit(s_u32(&bars[B_AFULL + acc]));      // accumulator of this tile complete
                }
                __syncwarp();
            }
            if (TL && tl && lt < WT_TL_MAX) { s_tl[lt * 8 + 1] = tm0 - tl0; s_tl[lt * 8 + 2] = tm1 - tl0; s_tl[lt * 8 + 3] = gtime_ns() - tl0; }
        }
    } else {
        // ============================== epilogue ==============================
        const int lq = warp & 3, half = ((warp - 2) >> 2) & 1; // TMEM lane quarter; pixel half [48 half, 48 half + 48)
        const int grp = (warp - 2) >> 3;                       // epilogue group: tiles lt = grp, grp + NGRP, ...
        const int par = lane >> 4;                             // 0: this lane holds the W_hi rows and keeps the even pixels, 1: W_lo rows, odd pixels
        const int oc = lq * 16 + (lane & 15);
        const float breg = EPI == 0 ? __ldg(bias + oc) : 0.f;
        for (uint32_t lt = grp; (long long)blockIdx.x + (long long)lt * gridDim.x < ntiles; lt += NGRP) {
            const int tile = blockIdx.x + lt * gridDim.x;
            const int acc = lt & 1;
            const int n = tile / tps, q0 = Wp + (tile - n * tps) * WT_N;
            const int qa = q0 + half * 48 + par;               // this lane's pixels: qa + 2 m, m = 0..23
            const size_t obase = ((size_t)n * PS + qa) * CT_C + oc;
            unsigned short ax[24];
            if (EPI == 1) {                                    // forward activation bits: issued before the wait, consumed after it
#pragma unroll
                for (int m = 0; m < 24; ++m)
                    ax[m] = qa + 2 * m < qend ? __ldg(reinterpret_cast<const unsigned short*>(aux_hi) + obase + (size_t)m * 2 * CT_C) : (unsigned short)0;
            }
            mb_wait(s_u32(&bars[B_AFULL + acc]), (lt >> 1) & 1);
            const unsigned long long te0 = TL ? gtime_ns() : 0ull;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[48];
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * WT_ACC1 + half * 48);
            if (!(var & 16)) {
                tmem_ld32(taddr, r);
                tmem_ld16(taddr + 32, r + 32);
            } else {
#pragma unroll
                for (int j = 0; j < 48; ++j) r[j] = 0u;
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mb_arrive(s_u32(&bars[B_AEMPTY + acc]));            // this warp's share of the accumulator is free
            const unsigned long long te1 = TL ? gtime_ns() : 0ull;
            if (var & 48) continue;
            int col = qa % Wp;
#pragma unroll
            for (int m = 0; m < 24; ++m) {
                // keep the pixel of this lane's parity, hand the other one to lane ^ 16 (which holds the other half of the same channel)
                const float own = __uint_as_float(par ? r[2 * m + 1] : r[2 * m]);
                const float give = __uint_as_float(par ? r[2 * m] : r[2 * m + 1]);
                float x = own + __shfl_xor_sync(0xFFFFFFFFu, give, 16);
                const int q = qa + 2 * m;
                const bool interior = col >= 1 && col <= W;
                col += 2;
                if (col >= Wp) col -= Wp;
                if (EPI == 0) { x += breg; x = fmaxf(x, 0.2f * x); }
                else x *= ((uint32_t)ax[m] - 1u < 0x7FFFu) ? 1.f : 0.2f;                  // forward activation positive and non-zero
                x = interior ? x : 0.f;
                if (q < qend && !(var & 2)) {
                    const size_t o = obase + (size_t)m * 2 * CT_C;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
                    out_hi[o] = hi;
                    out_lo[o] = __float2bfloat16_rn(x - __bfloat162float(hi));
                    if (F32OUT) out_f32[o] = x;
                }
            }
            if (TL && tl && (warp == 2 || warp == 10) && lt < WT_TL_MAX) {
                s_tl[lt * 8 + 4] = te0 - tl0; s_tl[lt * 8 + 5] = te1 - tl0; s_tl[lt * 8 + 6] = gtime_ns() - tl0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    if (TL && blockIdx.x == 0 && threadIdx.x == 0) {
        printf("[tl] CTA 0 (ns since entry): TMEM alloc + barriers %llu | weights in TMEM %llu | kernel end %llu\n", tl1 - tl0, tl2 - tl0, gtime_ns() - tl0);
        printf("[tl] tile: loads issued | accumulator free, first stage ready, MMAs issued | accumulator complete, TMEM read, stored\n");
        const int nloc = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int i = 0; i < nloc && i < WT_TL_MAX; ++i)
            printf("[tl] %2d: %6llu | %6llu %6llu %6llu | %6llu %6llu %6llu\n", i, s_tl[i * 8], s_tl[i * 8 + 1], s_tl[i * 8 + 2], s_tl[i * 8 + 3], s_tl[i * 8 + 4],
                   s_tl[i * 8 + 5], s_tl[i * 8 + 6]);
    }
}

// ------------------------------------------------------------------------------------------------ SIMT helpers (first / last layer, loss)
// layer 0 (1 -> 32 channels): x planar fp32 [N][1][PS] -> NHWC (hi,lo) rows, channels 32..63 stay zero.
// One thread per pixel; weights in shared memory; the 32 outputs leave as 4 + 4 16-byte stores (64 B of hi, 64 B of lo per pixel).
__global__ void __launch_bounds__(128) k_tc_first(const float* __restrict__ x, const float* __restrict__ w /*[32][1][9]*/, const float* __restrict__ b,
                                                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int H, int W, int Wp, int PS) {
    __shared__ float s_w[32 * 9 + 32];
    for (int i = threadIdx.x; i < 288; i += blockDim.x) s_w[i] = w[i];
    if (threadIdx.x < 32) s_w[288 + threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.y;
    const int q = Wp + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (H + 1) * Wp) return;
    const int col = q % Wp;
    const bool interior = col >= 1 && col <= W;
    float xin[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {            // the first / last border pixel of a plane has a neighbour outside it (its output is zeroed below)
        const int qq = q + (k / 3 - 1) * Wp + (k % 3 - 1);
        xin[k] = (qq >= 0 && qq < PS) ? x[(size_t)n * PS + qq] : 0.f;
    }
    uint4* oh = reinterpret_cast<uint4*>(out_hi + ((size_t)n * PS + q) * CT_C);
    uint4* ol = reinterpret_cast<uint4*>(out_lo + ((size_t)n * PS + q) * CT_C);
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c8 * 8 + p * 2 + e;
                float a = s_w[288 + c];
#pragma unroll
                for (int k = 0; k < 9; ++k) a = fmaf(s_w[c * 9 + k], xin[k], a);
                a = a > 0.f ? a : 0.2f * a;
                v[e] = interior ? a : 0.f;
            }
            uint32_t h0, l0, h1, l1;
            split_bf16(v[0], h0, l0);
            split_bf16(v[1], h1, l1);
            hw[p] = h0 | (h1 << 16);
            lw[p] = l0 | (l1 << 16);
        }
        oh[c8] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        ol[c8] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
}
// input gradient of layer 0 (32 -> 1): dpre NHWC (hi,lo) -> dx planar fp32.
// Two phases per CTA of LB_TILE output pixels: (1) every pixel of the tile + halo reads ITS OWN 128 B row once and reduces it against the
// 9 weight columns into shared memory T[pixel][9]; (2) dx[q] = sum_k T[q - off_k][k].  The halo is 2 Wp + 2 pixels whatever the tile, so
// the tile is large (1024 pixels = 8 image rows at Wp = 128: 25 % redundant rows; the first version used 256 pixels = 100 %).
constexpr int LB_TILE = 1024;
__global__ void __launch_bounds__(256) k_tc_last_bwd(const __nv_bfloat16* __restrict__ g_hi, const __nv_bfloat16* __restrict__ g_lo,
                                                     const float* __restrict__ w /*[32][1][9]*/, float* __restrict__ dx, int H, int W, int Wp, int PS) {
    extern __shared__ float s_t[];                 // [LB_TILE + 2*Wp + 2][9]
    __shared__ float s_w[32 * 9];
    const int n = blockIdx.y;
    const int q0 = Wp + blockIdx.x * LB_TILE;
    const int span = LB_TILE + 2 * Wp + 2;
    for (int i = threadIdx.x; i < 288; i += 256) s_w[i] = w[i];
    __syncthreads();
    for (int e = threadIdx.x; e < span; e += 256) {
        const int q = q0 - Wp - 1 + e;
        float t[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) t[k] = 0.f;
        if (q >= 0 && q < PS) {
            const uint4* ph = reinterpret_cast<const uint4*>(g_hi + ((size_t)n * PS + q) * CT_C);
            const uint4* pl = reinterpret_cast<const uint4*>(g_lo + ((size_t)n * PS + q) * CT_C);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {       // channels 0..31
                const uint4 h = ph[c8], l = pl[c8];
                const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float g0 = __uint_as_float(hw[p] << 16) + __uint_as_float(lw[p] << 16);
                    const float g1 = __uint_as_float(hw[p] & 0xFFFF0000u) + __uint_as_float(lw[p] & 0xFFFF0000u);
                    const int c = c8 * 8 + p * 2;
#pragma unroll
                    for (int k = 0; k < 9; ++k) t[k] = fmaf(g0, s_w[c * 9 + k], fmaf(g1, s_w[(c + 1) * 9 + k], t[k]));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) s_t[e * 9 + k] = t[k];
    }
    __syncthreads();
    for (int o = threadIdx.x; o < LB_TILE; o += 256) {
        const int q = q0 + o;
        if (q >= (H + 1) * Wp) break;
        const int col = q % Wp;
        float a = 0.f;
        if (col >= 1 && col <= W) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {          // dx[q] = sum_oc,k dpre[oc][q - off_k] * W[oc][k]
                const int e = o + Wp + 1 - ((k / 3 - 1) * Wp + (k % 3 - 1));
                a += s_t[e * 9 + k];
            }
        }
        dx[(size_t)n * PS + q] = a;
    }
}
// smoothness loss on the fp32 NHWC output + dpre of the last layer in (hi,lo) form (same math as fit.cu:k_smooth_loss).
// One thread per (pixel, 8 channels): float4 loads of z[q-1], z[q], z[q+1], one 16-byte store per plane.
__global__ void __launch_bounds__(256) k_tc_smooth_loss(const float* __restrict__ zf, int H, int W, int Wp, int PS, float w, int acc_stride,
                                                        int acc_slot, __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                                                        float* __restrict__ acc) {
    __shared__ float sred[32];
    const int s = blockIdx.y;
    const float inv_n = 1.f / (64.f * (float)H * (float)(W - 1));
    float part = 0.f;
    const long long tot = (long long)H * Wp * 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i & 7);
        const int q = Wp + (int)(i >> 3);
        const int col = q % Wp, x = col - 1;
        const size_t o = ((size_t)s * PS + q) * 64 + c8 * 8;
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        if (x >= 0 && x < W) {
            float zc[8], zl[8], zr[8];
            *reinterpret_cast<float4*>(zc) = *reinterpret_cast<const float4*>(zf + o);
            *reinterpret_cast<float4*>(zc + 4) = *reinterpret_cast<const float4*>(zf + o + 4);
            if (x >= 1) { *reinterpret_cast<float4*>(zl) = *reinterpret_cast<const float4*>(zf + o - 64); *reinterpret_cast<float4*>(zl + 4) = *reinterpret_cast<const float4*>(zf + o - 60); }
            if (x <= W - 2) { *reinterpret_cast<float4*>(zr) = *reinterpret_cast<const float4*>(zf + o + 64); *reinterpret_cast<float4*>(zr + 4) = *reinterpret_cast<const float4*>(zf + o + 68); }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
                if (x >= 1) a += zc[j] - zl[j];
                if (x <= W - 2) { const float d = zr[j] - zc[j]; a -= d; part += d * d; }
                g[j] = w * 2.f * inv_n * a * (zc[j] > 0.f ? 1.f : 0.2f);
            }
        }
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            uint32_t h0, l0, h1, l1;
            split_bf16(g[2 * p], h0, l0);
            split_bf16(g[2 * p + 1], h1, l1);
            hw[p] = h0 | (h1 << 16);
            lw[p] = l0 | (l1 << 16);
        }
        *reinterpret_cast<uint4*>(g_hi + o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(g_lo + o) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    part = block_sum(part, sred);
    if (threadIdx.x == 0) atomicAdd(&acc[s * acc_stride + acc_slot], part * inv_n);
}
// module API helpers: NHWC fp32 -> dense NCHW ; dense NCHW gradient (x LeakyReLU'(z)) -> NHWC (hi,lo)
__global__ void k_tc_unpack_z(const float* __restrict__ zf, float* __restrict__ dense, int N, int H, int W, int Wp, int PS) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * 64 * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % 64), n = (int)(i / ((long long)W * H * 64));
    dense[i] = zf[((size_t)n * PS + (y + 1) * Wp + x + 1) * 64 + c];
}
__global__ void k_tc_pack_dz(const float* __restrict__ dz, const float* __restrict__ zf, __nv_bfloat16* __restrict__ g_hi,
                             __nv_bfloat16* __restrict__ g_lo, int N, int H, int W, int Wp, int PS) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * 64 * H * W) return;
    const int x = (int)(i % W), y = (int)((i / W) % H), c = (int)((i / ((long long)W * H)) % 64), n = (int)(i / ((long long)W * H * 64));
    const size_t o = ((size_t)n * PS + (y + 1) * Wp + x + 1) * 64 + c;
    const float g = dz[i] * (zf[o] > 0.f ? 1.f : 0.2f);
    uint32_t h, l;
    split_bf16(g, h, l);
    g_hi[o] = __ushort_as_bfloat16((unsigned short)h);
    g_lo[o] = __ushort_as_bfloat16((unsigned short)l);
}
// weights: Conv2d W[oc][ic][9] fp32 -> forward tiles [tap][{hi,lo}][oc 64][ic 64] and input-gradient tiles [tap][{hi,lo}][ic 64][oc 64] (flipped taps)
__global__ void k_tc_prep_w(const float* __restrict__ w, const float* __restrict__ b, int Cin, int Cout, __nv_bfloat16* __restrict__ wf,
                            __nv_bfloat16* __restrict__ wb, float* __restrict__ bias64) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64) bias64[i] = i < Cout ? b[i] : 0.f;
    if (i >= 9 * 64 * 64) return;
    const int tap = i / 4096, r = (i / 64) % 64, c = i % 64;
    {   // forward: row = oc, col = ic
        const float v = (r < Cout && c < Cin) ? w[((size_t)r * Cin + c) * 9 + tap] : 0.f;
        uint32_t h, l;
        split_bf16(v, h, l);
        wf[((size_t)(tap * 2) * 64 + r) * 64 + c] = __ushort_as_bfloat16((unsigned short)h);
        wf[((size_t)(tap * 2 + 1) * 64 + r) * 64 + c] = __ushort_as_bfloat16((unsigned short)l);
    }
    {   // input gradient: row = ic (output channel of the adjoint conv), col = oc, tap flipped
        const float v = (c < Cout && r < Cin) ? w[((size_t)c * Cin + r) * 9 + (8 - tap)] : 0.f;
        uint32_t h, l;
        split_bf16(v, h, l);
        wb[((size_t)(tap * 2) * 64 + r) * 64 + c] = __ushort_as_bfloat16((unsigned short)h);
        wb[((size_t)(tap * 2 + 1) * 64 + r) * 64 + c] = __ushort_as_bfloat16((unsigned short)l);
    }
}

// TMEM image of a layer for k_conv_tc_wt: dst[tap][i 8][r 128][w 4] (uint32 = two consecutive-K bf16) = tile word (4 i + w) of row
// (part, oc) of tap, where TMEM lane r = 32 q + l holds part = l >> 4 of channel oc = 16 q + (l & 15).
__global__ void k_tc_prep_wt(const __nv_bfloat16* __restrict__ tiles /*[tap][{hi,lo}][64][64]*/, uint32_t* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 9 * 8 * 128 * 4) return;
    const int w = idx & 3, r = (idx >> 2) & 127, i = (idx >> 9) & 7, tap = idx >> 12;
    const int q = r >> 5, l = r & 31, part = l >> 4, oc = q * 16 + (l & 15);
    dst[idx] = reinterpret_cast<const uint32_t*>(tiles)[((size_t)(tap * 2 + part) * 64 + oc) * 32 + i * 4 + w];
}

// ------------------------------------------------------------------------------------------------ host side
struct EncTC {
    int maxN = 0;
    PlaneGeom g{};
    __nv_bfloat16 *a_hi[11] = {}, *a_lo[11] = {};       // a_*[l] = input of layers[l] for l = 1..9, a_*[10] = final activation
    float* zf = nullptr;                                // final activation, fp32 NHWC
    __nv_bfloat16 *g_hi[2] = {}, *g_lo[2] = {};
    __nv_bfloat16 *wf[10] = {}, *wb[10] = {};
    uint32_t *wtf[10] = {}, *wtb[10] = {};              // TMEM images of wf / wb for the weights-in-TMEM kernel
    float* bias[10] = {};
    CUtensorMap m_a_hi[11], m_a_lo[11], m_g_hi[2], m_g_lo[2], m_wf[10], m_wb[10];
    CUtensorMap r_a_hi[11], r_a_lo[11], r_g_hi[2], r_g_lo[2];   // 136-row boxes for the row-reuse variant
    CUtensorMap w_a_hi[11], w_a_lo[11], w_g_hi[2], w_g_lo[2];   // 104-row boxes for the weights-in-TMEM variant
    int sm_count = 148;
};

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_bf16_map(CUtensorMap* m, const void* base, long long rows, int box_rows) {
    static PFN_enc fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (PFN_enc)p;
    }
    LEMO_CHECK(fn, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LEMO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
    return 0;
}
template <typename T>
static int zalloc(T** p, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    LEMO_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}

// 0 = CUDA-core fp32 path (conv.cu).  Tensor-core variants, in the order they were developed and measured on B200 (64->64 layer,
// S=8, T=120; tools/diag_conv_modes.py, tools/diag_conv_wt.py): 2 = one TMA box per tap, resident weights (92.6 us); 3 = row reuse (76.7 us);
// 4 = + streamed weights, 3 A stages (71.2 us); 5 = + stacked [W_hi;W_lo] N=128 MMA (68.9 us); 6 = + 8 epilogue warps (61.8 us);
// 1 (LEMO_CONV=pair) = pair kernel: weight slot shared by two tiles, lean epilogue, converged MMA/TMA issue (67.9 us forward / 64.7 us input
// gradient in the final measurement series); 7 = same with 16 epilogue warps.
// 8..8191: timing experiments on the pair kernel (bit mask, see tools/diag_conv_modes.py) -- results are garbage by construction.
// 8192 (LEMO_CONV=wt) = DEFAULT = weights-in-TMEM kernel k_conv_tc_wt (57.6 us forward / 59.8 us input gradient); 8192 + bits = its
// experiments (tools/diag_conv_wt.py).
constexpr int CONV_TC_DEFAULT = 8192;
static int g_conv_tc = -1;
static void conv_tc_init() {
    if (g_conv_tc < 0) {
        const char* e = getenv("LEMO_CONV");
        g_conv_tc = (e && strcmp(e, "simt") == 0) ? 0 : (e && strcmp(e, "wt") == 0) ? 8192 : (e && strcmp(e, "pair") == 0) ? 1
                    : (e && e[0] >= '0' && e[0] <= '9') ? atoi(e) : CONV_TC_DEFAULT;
    }
}
bool conv_tc_enabled() { conv_tc_init(); return g_conv_tc >= 1; }
void conv_tc_set(int on) { g_conv_tc = on < 0 ? -1 : (on > 16383 ? 16383 : on); }     // -1: re-read LEMO_CONV / default at the next use
static inline int conv_tc_maps() { return g_conv_tc >= 8192 ? 2 : g_conv_tc == 2 ? 0 : 1; }   // which activation box: 128 / 136 / 104 rows

int enc_tc_refresh_weights(ConvNet* n, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    for (int l = 1; l < 10; ++l) {
        const ConvLayer& L = n->layers[l];
        k_tc_prep_w<<<cdiv(9 * 64 * 64, 256), 256, 0, st>>>(n->w_flat + L.w_off, n->w_flat + L.b_off, L.Cin, L.Cout, t->wf[l], t->wb[l], t->bias[l]);
        k_tc_prep_wt<<<cdiv(9 * 8 * 128 * 4, 256), 256, 0, st>>>(t->wf[l], t->wtf[l]);
        k_tc_prep_wt<<<cdiv(9 * 8 * 128 * 4, 256), 256, 0, st>>>(t->wb[l], t->wtb[l]);
    }
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int enc_tc_create(ConvNet* n) {
    EncTC* t = new EncTC();
    n->tc = t;
    t->maxN = n->maxN; t->g = n->geom[0];
    const size_t rows = (size_t)n->maxN * t->g.PS;
    for (int l = 1; l <= 10; ++l) {
        LEMO_TRY(zalloc(&t->a_hi[l], rows * 64)); LEMO_TRY(zalloc(&t->a_lo[l], rows * 64));
        LEMO_TRY(make_bf16_map(&t->m_a_hi[l], t->a_hi[l], rows, CT_M)); LEMO_TRY(make_bf16_map(&t->m_a_lo[l], t->a_lo[l], rows, CT_M));
        LEMO_TRY(make_bf16_map(&t->r_a_hi[l], t->a_hi[l], rows, CT_M2)); LEMO_TRY(make_bf16_map(&t->r_a_lo[l], t->a_lo[l], rows, CT_M2));
        LEMO_TRY(make_bf16_map(&t->w_a_hi[l], t->a_hi[l], rows, WT_BOX)); LEMO_TRY(make_bf16_map(&t->w_a_lo[l], t->a_lo[l], rows, WT_BOX));
    }
    LEMO_TRY(zalloc(&t->zf, rows * 64));
    if (n->with_backward)
        for (int i = 0; i < 2; ++i) {
            LEMO_TRY(zalloc(&t->g_hi[i], rows * 64)); LEMO_TRY(zalloc(&t->g_lo[i], rows * 64));
            LEMO_TRY(make_bf16_map(&t->m_g_hi[i], t->g_hi[i], rows, CT_M)); LEMO_TRY(make_bf16_map(&t->m_g_lo[i], t->g_lo[i], rows, CT_M));
            LEMO_TRY(make_bf16_map(&t->r_g_hi[i], t->g_hi[i], rows, CT_M2)); LEMO_TRY(make_bf16_map(&t->r_g_lo[i], t->g_lo[i], rows, CT_M2));
            LEMO_TRY(make_bf16_map(&t->w_g_hi[i], t->g_hi[i], rows, WT_BOX)); LEMO_TRY(make_bf16_map(&t->w_g_lo[i], t->g_lo[i], rows, WT_BOX));
        }
    for (int l = 1; l < 10; ++l) {
        LEMO_TRY(zalloc(&t->wf[l], (size_t)18 * 64 * 64)); LEMO_TRY(zalloc(&t->wb[l], (size_t)18 * 64 * 64)); LEMO_TRY(zalloc(&t->bias[l], 64));
        LEMO_TRY(zalloc(&t->wtf[l], (size_t)9 * 8 * 128 * 4)); LEMO_TRY(zalloc(&t->wtb[l], (size_t)9 * 8 * 128 * 4));
        LEMO_TRY(make_bf16_map(&t->m_wf[l], t->wf[l], 18 * 64, CT_C)); LEMO_TRY(make_bf16_map(&t->m_wb[l], t->wb[l], 18 * 64, CT_C));
    }
    cudaDeviceProp prop;
    LEMO_CUDA(cudaGetDeviceProperties(&prop, n->device));
    t->sm_count = prop.multiProcessorCount;
    LEMO_CUDA(cudaFuncSetAttribute(k_tc_last_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM2));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_ws<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_ws<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_ws<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<8, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<8, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<8, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<16, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<16, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_pair<16, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW_SMEM));
#define LEMO_WT_ATTR(E, F, K) \
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_wt<E, F, K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM)); \
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_wt<E, F, K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM))
    LEMO_CUDA(cudaFuncSetAttribute(k_conv_tc_wt<0, false, 4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM));
    LEMO_WT_ATTR(0, false, 4); LEMO_WT_ATTR(0, true, 4); LEMO_WT_ATTR(1, false, 4); LEMO_WT_ATTR(0, false, 2); LEMO_WT_ATTR(1, false, 2);
#undef LEMO_WT_ATTR
    return enc_tc_refresh_weights(n, 0);
}
void enc_tc_free(ConvNet* n) {
    EncTC* t = (EncTC*)n->tc;
    if (!t) return;
    for (int l = 0; l <= 10; ++l) { cudaFree(t->a_hi[l]); cudaFree(t->a_lo[l]); }
    cudaFree(t->zf);
    for (int i = 0; i < 2; ++i) { cudaFree(t->g_hi[i]); cudaFree(t->g_lo[i]); }
    for (int l = 0; l < 10; ++l) { cudaFree(t->wf[l]); cudaFree(t->wb[l]); cudaFree(t->bias[l]); cudaFree(t->wtf[l]); cudaFree(t->wtb[l]); }
    delete t;
    n->tc = nullptr;
}

static int launch_tc(const EncTC* t, const CUtensorMap& mh, const CUtensorMap& ml, const CUtensorMap& mw, const uint32_t* wimg, const float* bias,
                     const __nv_bfloat16* aux, __nv_bfloat16* oh, __nv_bfloat16* ol, float* of32, int N, int epi, cudaStream_t st, int kin = CT_C) {
    const PlaneGeom& g = t->g;
    const int kmax = std::min(CT_C / 16, (kin + 15) / 16);     // real input channels of this layer, in MMA K steps
    const int ntiles = N * cdiv((long long)g.H * g.Wp, CT_M);
    const int grid = std::min(ntiles, t->sm_count);
    conv_tc_init();
    if (g_conv_tc >= 8192) {
        const int var = g_conv_tc - 8192;
        const int wtiles = N * cdiv((long long)g.H * g.Wp, WT_N);
        const int wg = std::min(wtiles, t->sm_count);
        const uint32_t* wq = wimg;
#define LEMO_WT(E, F, K) do { if (var & 64) k_conv_tc_wt<E, F, K, 1><<<wg, 320, WT_SMEM, st>>>(mh, ml, wq, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, var); \
                              else k_conv_tc_wt<E, F, K, 2><<<wg, 576, WT_SMEM, st>>>(mh, ml, wq, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, var); } while (0)
        // var bit 6: one epilogue group of 8 warps instead of two (A/B measurement); bit 7: timeline print of CTA 0 (64->64 forward layers only)
        if ((var & 128) && kmax > 2 && epi == 0 && !of32)
            k_conv_tc_wt<0, false, 4, 2, true><<<wg, 576, WT_SMEM, st>>>(mh, ml, wq, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, var & ~128);
        else if (kmax <= 2) { if (epi == 1) LEMO_WT(1, false, 2); else LEMO_WT(0, false, 2); }      // 32-channel layers (never the fp32-output layer)
        else if (epi == 1) LEMO_WT(1, false, 4);
        else if (of32) LEMO_WT(0, true, 4);
        else LEMO_WT(0, false, 4);
#undef LEMO_WT
    } else if (g_conv_tc == 1 || g_conv_tc >= 7) {
        const int dbg = g_conv_tc >= 8 ? g_conv_tc - 8 : 0;      // 8 + bit mask: timing experiments (tools/diag_conv_modes.py), results invalid
        const int pg = std::min((ntiles + 1) / 2, t->sm_count);
#define LEMO_PAIR(NE, E, F) k_conv_tc_pair<NE, E, F><<<pg, 64 + 32 * NE, CW_SMEM, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, dbg, kmax)
        if (g_conv_tc != 7) { if (epi == 1) LEMO_PAIR(8, 1, false); else if (of32) LEMO_PAIR(8, 0, true); else LEMO_PAIR(8, 0, false); }
        else { if (epi == 1) LEMO_PAIR(16, 1, false); else if (of32) LEMO_PAIR(16, 0, true); else LEMO_PAIR(16, 0, false); }
#undef LEMO_PAIR
    } else if (g_conv_tc == 5) k_conv_tc_ws<true, 4><<<grid, 192, CW_SMEM, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, epi);
    else if (g_conv_tc == 4) k_conv_tc_ws<false, 4><<<grid, 192, CW_SMEM, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, epi);
    else if (g_conv_tc == 3) k_conv_tc<1><<<grid, 192, CT_SMEM2, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, epi);
    else if (g_conv_tc == 2) k_conv_tc<0><<<grid, 192, CT_SMEM, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, epi);
    else k_conv_tc_ws<true, 8><<<grid, 320, CW_SMEM, st>>>(mh, ml, mw, bias, aux, oh, ol, of32, N, g.H, g.W, g.Wp, g.PS, epi);    // 6
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// x planar fp32 [N][1][PS] -> a_*[10] / zf
int enc_tc_forward(ConvNet* n, const float* x_planes, int N, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    const PlaneGeom& g = t->g;
    const ConvLayer& L0 = n->layers[0];
    k_tc_first<<<dim3(cdiv((long long)g.H * g.Wp, 128), N), 128, 0, st>>>(x_planes, n->w_flat + L0.w_off, n->w_flat + L0.b_off, t->a_hi[1], t->a_lo[1],
                                                                          g.H, g.W, g.Wp, g.PS);
    conv_tc_init();
    const int mk = conv_tc_maps();
    for (int l = 1; l < 10; ++l)
        LEMO_TRY(launch_tc(t, mk == 2 ? t->w_a_hi[l] : mk ? t->r_a_hi[l] : t->m_a_hi[l], mk == 2 ? t->w_a_lo[l] : mk ? t->r_a_lo[l] : t->m_a_lo[l],
                           t->m_wf[l], t->wtf[l], t->bias[l], nullptr, t->a_hi[l + 1], t->a_lo[l + 1], l == 9 ? t->zf : nullptr, N, 0, st,
                           n->layers[l].Cin));
    n->launches += 10;
    return 0;
}
int enc_tc_smooth_loss(ConvNet* n, int N, float w, int acc_stride, int acc_slot, float* acc, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    const PlaneGeom& g = t->g;
    k_tc_smooth_loss<<<dim3(256, N), 256, 0, st>>>(t->zf, g.H, g.W, g.Wp, g.PS, w, acc_stride, acc_slot, t->g_hi[0], t->g_lo[0], acc);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
// dpre of the last layer in g_*[0]  ->  dx planar fp32
int enc_tc_backward(ConvNet* n, int N, float* dx_planes, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    const PlaneGeom& g = t->g;
    int cur = 0;
    conv_tc_init();
    const int mk = conv_tc_maps();
    for (int l = 9; l >= 1; --l) {
        LEMO_TRY(launch_tc(t, mk == 2 ? t->w_g_hi[cur] : mk ? t->r_g_hi[cur] : t->m_g_hi[cur], mk == 2 ? t->w_g_lo[cur] : mk ? t->r_g_lo[cur] : t->m_g_lo[cur],
                           t->m_wb[l], t->wtb[l], nullptr, t->a_hi[l], t->g_hi[cur ^ 1], t->g_lo[cur ^ 1], nullptr, N, 1, st, n->layers[l].Cout));
        cur ^= 1;
    }
    const ConvLayer& L0 = n->layers[0];
    const size_t lb_smem = (size_t)(LB_TILE + 2 * g.Wp + 2) * 9 * sizeof(float);
    LEMO_CHECK(lb_smem <= 200 * 1024, "plane pitch too wide for k_tc_last_bwd");
    k_tc_last_bwd<<<dim3(cdiv((long long)g.H * g.Wp, LB_TILE), N), 256, lb_smem, st>>>(
        t->g_hi[cur], t->g_lo[cur], n->w_flat + L0.w_off, dx_planes, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    n->launches += 10;
    return 0;
}
// measurement hook: relaunch one tensor-core layer `reps` times on resident buffers (forward layer l in 1..9, or its input gradient)
int enc_tc_profile_layer(ConvNet* n, int layer, int N, int backward, int reps, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    LEMO_CHECK(layer >= 1 && layer <= 9, "tensor-core layers are 1..9");
    conv_tc_init();
    for (int r = 0; r < reps; ++r) {
        const int mk = conv_tc_maps();
        if (!backward)
            LEMO_TRY(launch_tc(t, mk == 2 ? t->w_a_hi[layer] : mk ? t->r_a_hi[layer] : t->m_a_hi[layer],
                               mk == 2 ? t->w_a_lo[layer] : mk ? t->r_a_lo[layer] : t->m_a_lo[layer], t->m_wf[layer], t->wtf[layer], t->bias[layer],
                               nullptr, t->a_hi[layer + 1], t->a_lo[layer + 1], nullptr, N, 0, st));
        else
            LEMO_TRY(launch_tc(t, mk == 2 ? t->w_g_hi[0] : mk ? t->r_g_hi[0] : t->m_g_hi[0], mk == 2 ? t->w_g_lo[0] : mk ? t->r_g_lo[0] : t->m_g_lo[0],
                               t->m_wb[layer], t->wtb[layer], nullptr, t->a_hi[layer], t->g_hi[1], t->g_lo[1], nullptr, N, 1, st));
    }
    return 0;
}
int enc_tc_unpack_z(ConvNet* n, int N, float* z_dense, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    const PlaneGeom& g = t->g;
    const long long tot = (long long)N * 64 * g.H * g.W;
    k_tc_unpack_z<<<cdiv(tot, 256), 256, 0, st>>>(t->zf, z_dense, N, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int enc_tc_pack_dz(ConvNet* n, int N, const float* dz_dense, cudaStream_t st) {
    EncTC* t = (EncTC*)n->tc;
    const PlaneGeom& g = t->g;
    const long long tot = (long long)N * 64 * g.H * g.W;
    k_tc_pack_dz<<<cdiv(tot, 256), 256, 0, st>>>(dz_dense, t->zf, t->g_hi[0], t->g_lo[0], N, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
