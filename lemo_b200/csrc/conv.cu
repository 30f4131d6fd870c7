// conv3x3 stack on padded pitch-linear planes (see conv.cuh for the layout argument).
// Reference semantics: models/AE_sep.py:11-30,77-99 (Enc: 10x conv3x3 p1 + LeakyReLU 0.2, no pooling) and
// models/AE.py:11-108 (AE: + MaxPool(3,2,1), ConvTranspose(s2, output_size)).  fp32 CUDA-core FMA:
// the north star keeps tensor cores for the blend-shape GEMM only, and 1e-4 parity rules out single-pass TF32.
#include "conv.cuh"
#include "../../include/lemo_b200.h"
#include <algorithm>
#include <mutex>

namespace lemo {

constexpr int TP = 256;   // output pixels (linear) per CTA
constexpr int CT = 8;     // input channels staged per smem pass

// ------------------------------------------------------------------------------------------------
// main kernel: CTA = NW warps; warp w owns output channels ocb0 + 8w .. +7 for all 256 pixels of the tile;
// lane g owns pixels 4g..4g+3 and 128+4g..128+4g+3  -> 8 oc x 8 px register tile, weights are warp-uniform
// (broadcast LDS.128), inputs are conflict-free LDS.128 + LDS.64 per (ic,ky).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 4 : 0;                         // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NW>
__global__ void __launch_bounds__(NW * 32) k_conv3x3(const float* __restrict__ in, const float* __restrict__ wk,
                                                     const float* __restrict__ bias, const float* __restrict__ aux,
                                                     float* __restrict__ out, int Cin, int Cout, int H, int W, int Wp, int PS,
                                                     int SW, int epi, int KS, int cps, int Nn) {
    // KS > 1: blockIdx.z = n * KS + ks; this CTA contracts input channels [ks*cps, (ks+1)*cps) only and stores its RAW partial sums
    // into out = scratch[ks][n][oc][q]; k_conv_splitk_finish adds the slices in order and applies the epilogue (deterministic).
    // (Fusing that finish into this kernel -- the last-arriving CTA of a tile sums the slices, arrival counter per tile -- was measured:
    // it removes 34 launches per AE step but confines each tile's 16 KS float4 loads per thread to ONE SM; on the deep levels (4 tiles,
    // 32 slices) the tail is longer than the launch it saves: 71 -> 117 ms per clip.  The separate kernel spreads the sums over all SMs.)
    // Small feature maps (the deep AE levels: 16x16 planes, 256 channels) would otherwise run on 4-30 CTAs.
    constexpr int OCB = NW * 8, NT = NW * 32;
    extern __shared__ __align__(16) float smem[];
    const int stage_floats = CT * SW + CT * 9 * OCB;    // one pipeline stage: input halo tile [CT][SW] + weights [CT][9][OCB]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = blockIdx.z / KS, ks = blockIdx.z - n * KS, ocb0 = blockIdx.y * OCB;
    const int q0 = Wp + blockIdx.x * TP;
    const float* in_n = in + (size_t)n * Cin * PS;
    const int ic_lo = ks * cps, ic_hi = min(Cin, ic_lo + cps);

    float acc[8][8];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[o][i] = 0.f;

    // cp.async double buffering: the global->shared copy of channel block k+1 overlaps the FMAs of block k
    // (ncu on the single-buffered version: long_scoreboard was the top stall, fma pipe 42 % active)
    auto issue_stage = [&](int ic0, float* st) {
        float* s_in = st;
        float* s_w = st + CT * SW;
        const int nic = min(CT, ic_hi - ic0);
        for (int ic = 0; ic < nic; ++ic) {
            const float* src = in_n + (size_t)(ic0 + ic) * PS;
            for (int e = tid; e < SW; e += NT) {
                const int q = q0 - Wp - 1 + e;
                const bool ok = (q >= 0 && q < PS);
                cp_async4(s_in + ic * SW + e, ok ? src + q : src, ok);
            }
        }
        const int nw4 = nic * 9 * OCB / 4;
        for (int idx = tid; idx < nw4; idx += NT) {
            const int o4 = idx % (OCB / 4), r = idx / (OCB / 4);
            cp_async16(s_w + r * OCB + o4 * 4, wk + (size_t)(ic0 * 9 + r) * Cout + ocb0 + o4 * 4);
        }
        cp_async_commit();
    };

    const int ntiles = (ic_hi - ic_lo + CT - 1) / CT;
    issue_stage(ic_lo, smem);
    for (int it = 0; it < ntiles; ++it) {
        float* st = smem + (it & 1) * stage_floats;
        if (it + 1 < ntiles) { issue_stage(ic_lo + (it + 1) * CT, smem + ((it + 1) & 1) * stage_floats); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* s_in = st;
        const float* s_w = st + CT * SW;
        const int nic = min(CT, ic_hi - ic_lo - it * CT);
        for (int ic = 0; ic < nic; ++ic) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* p = s_in + ic * SW + ky * Wp + 4 * lane;
                const float4 a0 = *reinterpret_cast<const float4*>(p);
                const float2 a1 = *reinterpret_cast<const float2*>(p + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(p + 128);
                const float2 b1 = *reinterpret_cast<const float2*>(p + 132);
                const float xa[6] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y};
                const float xb[6] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float* wp = s_w + (ic * 9 + ky * 3 + kx) * OCB + warp * 8;
                    const float4 w0 = *reinterpret_cast<const float4*>(wp);
                    const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int o = 0; o < 8; ++o)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            acc[o][i] = fmaf(wv[o], xa[i + kx], acc[o][i]);
                            acc[o][4 + i] = fmaf(wv[o], xb[i + kx], acc[o][4 + i]);
                        }
                }
            }
        }
        __syncthreads();
    }

    const int qend = (H + 1) * Wp;
    const size_t obase = ((size_t)n * Cout + ocb0 + warp * 8) * PS;
    if (KS > 1) {                                        // raw partials; epilogue in k_conv_splitk_finish
        const size_t pbase = (((size_t)ks * Nn + n) * Cout + ocb0 + warp * 8) * PS;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int q = q0 + half * 128 + 4 * lane;
            if (q >= qend) continue;
#pragma unroll
            for (int o = 0; o < 8; ++o)
                *reinterpret_cast<float4*>(out + pbase + (size_t)o * PS + q) =
                    make_float4(acc[o][half * 4], acc[o][half * 4 + 1], acc[o][half * 4 + 2], acc[o][half * 4 + 3]);
        }
        return;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int q = q0 + half * 128 + 4 * lane;
        if (q >= qend) continue;
        const int col = q % Wp;
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ok[i] = (col + i >= 1) && (col + i <= W);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            float v[4];
            const float bo = (epi == EPI_BIAS_LRELU || epi == EPI_BIAS) ? __ldg(bias + ocb0 + warp * 8 + o) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = acc[o][half * 4 + i] + bo;
            if (epi == EPI_BIAS_LRELU) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = lrelu(v[i]);
            } else if (epi == EPI_MASK) {
                const float4 a = *reinterpret_cast<const float4*>(aux + obase + (size_t)o * PS + q);
                v[0] *= a.x > 0.f ? 1.f : 0.2f; v[1] *= a.y > 0.f ? 1.f : 0.2f;
                v[2] *= a.z > 0.f ? 1.f : 0.2f; v[3] *= a.w > 0.f ? 1.f : 0.2f;
            }
            float4 r;
            r.x = ok[0] ? v[0] : 0.f; r.y = ok[1] ? v[1] : 0.f; r.z = ok[2] ? v[2] : 0.f; r.w = ok[3] ? v[3] : 0.f;
            *reinterpret_cast<float4*>(out + obase + (size_t)o * PS + q) = r;
        }
    }
}

// out[n][oc][q] = epi( sum_ks partial[ks][n][oc][q] ) on the interior rows; the slices are added in slice order (bitwise reproducible)
__global__ void __launch_bounds__(256) k_conv_splitk_finish(const float* __restrict__ partial, const float* __restrict__ bias,
                                                            const float* __restrict__ aux, float* __restrict__ out, int KS, int N, int Cout,
                                                            int H, int W, int Wp, int PS, int epi) {
    const int span = H * Wp;                               // linear range [Wp, (H+1) Wp)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * Cout * span) return;
    const int q = Wp + (int)(i % span);
    const long long c = i / span;                           // n * Cout + oc
    const int oc = (int)(c % Cout);
    const size_t o = (size_t)c * PS + q;
    const size_t slice = (size_t)N * Cout * PS;
    float v = 0.f;
    for (int ks = 0; ks < KS; ++ks) v += partial[(size_t)ks * slice + o];
    if (epi == EPI_BIAS_LRELU || epi == EPI_BIAS) v += __ldg(bias + oc);
    if (epi == EPI_BIAS_LRELU) v = lrelu(v);
    else if (epi == EPI_MASK) v *= aux[o] > 0.f ? 1.f : 0.2f;
    const int col = q % Wp;
    out[o] = (col >= 1 && col <= W) ? v : 0.f;
}

// few output channels (the 32->1 input-gradient layer of Enc, AE's 32->1 / 1->1 output layers): thread per pixel
template <int CO>
__global__ void __launch_bounds__(256) k_conv3x3_small(const float* __restrict__ in, const float* __restrict__ wk,
                                                       const float* __restrict__ bias, const float* __restrict__ aux,
                                                       float* __restrict__ out, int Cin, int H, int W, int Wp, int PS, int epi) {
    const int n = blockIdx.z;
    const int q = Wp + blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (H + 1) * Wp) return;
    const int col = q % Wp;
    const bool ok = col >= 1 && col <= W;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
    if (ok) {
        const float* in_n = in + (size_t)n * Cin * PS;
        for (int ic = 0; ic < Cin; ++ic) {
            const float* p = in_n + (size_t)ic * PS + q;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float x = __ldg(p + (ky - 1) * Wp + (kx - 1));
#pragma unroll
                    for (int o = 0; o < CO; ++o) acc[o] = fmaf(x, __ldg(wk + (ic * 9 + ky * 3 + kx) * CO + o), acc[o]);
                }
        }
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        float v = acc[o];
        if (epi == EPI_BIAS_LRELU || epi == EPI_BIAS) v += bias[o];
        if (epi == EPI_BIAS_LRELU) v = lrelu(v);
        else if (epi == EPI_MASK) v *= aux[((size_t)n * CO + o) * PS + q] > 0.f ? 1.f : 0.2f;
        out[((size_t)n * CO + o) * PS + q] = ok ? v : 0.f;
    }
}

// split-K factor for a layer: enough CTAs for ~2 per SM, slices of whole 8-channel stages
int conv3x3_splitk(int N, int Cin, int Cout, const PlaneGeom& g) {
    if (Cout % 32 != 0) return 1;
    const int ocb = (Cout % 64 == 0) ? 64 : 32;
    const long long ctas = (long long)cdiv((long long)g.H * g.Wp, TP) * (Cout / ocb) * N;
    if (ctas >= 148 || Cin < 2 * CT) return 1;
    return (int)std::max(1LL, std::min<long long>(Cin / CT, (296 + ctas - 1) / ctas));
}
size_t conv3x3_splitk_floats(int N, int Cin, int Cout, const PlaneGeom& g) {
    const int ks = conv3x3_splitk(N, Cin, Cout, g);
    return ks > 1 ? (size_t)ks * N * Cout * g.PS : 0;
}

template <int NW>
static int conv_main_launch(const float* in, const float* wk, const float* bias, const float* aux, float* out, int N, int Cin,
                            int Cout, const PlaneGeom& g, ConvEpi epi, cudaStream_t st, float* scratch, size_t scratch_floats) {
    const int SW = (TP + 2 * g.Wp + 2 + 3) / 4 * 4;
    const size_t smem = 2 * (size_t)(CT * SW + CT * 9 * NW * 8) * sizeof(float);      // two pipeline stages
    // one process per GPU: a per-process cache is enough -- but several host threads may launch (InfillPool drives
    // every stage from its own thread), and the limit must only ever GROW: an unlocked check-then-set let a thread with a smaller layer
    // overwrite a larger limit another thread had just set, whose next launch then failed inside its stream capture
    {
        static std::mutex mu;
        static size_t configured = 0;
        std::lock_guard<std::mutex> lk(mu);
        if (smem > configured) {
            LEMO_CUDA(cudaFuncSetAttribute(k_conv3x3<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
    }
    int KS = scratch ? conv3x3_splitk(N, Cin, Cout, g) : 1;
    if (KS > 1 && (size_t)KS * N * Cout * g.PS > scratch_floats) KS = 1;
    if (KS > 1) {
        const int cps = cdiv(cdiv(Cin, KS), CT) * CT;             // channels per slice, whole stages
        KS = cdiv(Cin, cps);
        dim3 grid(cdiv((long long)g.H * g.Wp, TP), Cout / (NW * 8), N * KS);
        k_conv3x3<NW><<<grid, NW * 32, smem, st>>>(in, wk, bias, aux, scratch, Cin, Cout, g.H, g.W, g.Wp, g.PS, SW, (int)epi, KS, cps, N);
        const long long tot = (long long)N * Cout * g.H * g.Wp;
        k_conv_splitk_finish<<<cdiv(tot, 256), 256, 0, st>>>(scratch, bias, aux, out, KS, N, Cout, g.H, g.W, g.Wp, g.PS, (int)epi);
    } else {
        dim3 grid(cdiv((long long)g.H * g.Wp, TP), Cout / (NW * 8), N);
        k_conv3x3<NW><<<grid, NW * 32, smem, st>>>(in, wk, bias, aux, out, Cin, Cout, g.H, g.W, g.Wp, g.PS, SW, (int)epi, 1, Cin, N);
    }
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int conv3x3_launch(const float* in, const float* wk, const float* bias, const float* aux, float* out, int N, int Cin, int Cout,
                   const PlaneGeom& g, ConvEpi epi, cudaStream_t st, float* scratch, size_t scratch_floats) {
    if (Cout % 64 == 0) return conv_main_launch<8>(in, wk, bias, aux, out, N, Cin, Cout, g, epi, st, scratch, scratch_floats);
    if (Cout % 32 == 0) return conv_main_launch<4>(in, wk, bias, aux, out, N, Cin, Cout, g, epi, st, scratch, scratch_floats);
    dim3 grid(cdiv((long long)g.H * g.Wp, 256), 1, N);
    if (Cout == 1) k_conv3x3_small<1><<<grid, 256, 0, st>>>(in, wk, bias, aux, out, Cin, g.H, g.W, g.Wp, g.PS, (int)epi);
    else if (Cout == 4) k_conv3x3_small<4><<<grid, 256, 0, st>>>(in, wk, bias, aux, out, Cin, g.H, g.W, g.Wp, g.PS, (int)epi);
    else { set_error("conv3x3: unsupported output channel count"); return 2; }
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------------
__global__ void k_pack(const float* __restrict__ dense, float* __restrict__ planes, long long total, int H, int W, int Wp, int PS) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const long long c = r / H;
    planes[c * PS + (y + 1) * Wp + x + 1] = dense[i];
}
__global__ void k_unpack(const float* __restrict__ planes, float* __restrict__ dense, long long total, int H, int W, int Wp, int PS) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const long long c = r / H;
    dense[i] = planes[c * PS + (y + 1) * Wp + x + 1];
}
// planes = dz * LeakyReLU'(z_planes)   (dz dense)
__global__ void k_pack_mask(const float* __restrict__ dz, const float* __restrict__ zpl, float* __restrict__ planes, long long total,
                            int H, int W, int Wp, int PS) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const long long c = r / H;
    const long long q = c * PS + (y + 1) * Wp + x + 1;
    planes[q] = dz[i] * (zpl[q] > 0.f ? 1.f : 0.2f);
}
int pack_planes(const float* dense, float* planes, int NC, const PlaneGeom& g, cudaStream_t st) {
    const long long total = (long long)NC * g.H * g.W;
    k_pack<<<cdiv(total, 256), 256, 0, st>>>(dense, planes, total, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int unpack_planes(const float* planes, float* dense, int NC, const PlaneGeom& g, cudaStream_t st) {
    const long long total = (long long)NC * g.H * g.W;
    k_unpack<<<cdiv(total, 256), 256, 0, st>>>(planes, dense, total, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weights: state_dict order -> kernel layouts
// ------------------------------------------------------------------------------------------------
__global__ void k_prep_weights(const float* __restrict__ w, int Cin, int Cout, int transposed, float* __restrict__ wk_f,
                               float* __restrict__ wk_b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Cin * Cout * 9) return;
    const int k = i % 9, r = i / 9;
    if (!transposed) {                       // nn.Conv2d  W[oc][ic][k]
        const int ic = r % Cin, oc = r / Cin;
        wk_f[((size_t)ic * 9 + k) * Cout + oc] = w[i];
        wk_b[((size_t)oc * 9 + (8 - k)) * Cin + ic] = w[i];
    } else {                                 // nn.ConvTranspose2d  W[ic][oc][k]  (stride-1/pad-1 == conv with flipped taps)
        const int oc = r % Cout, ic = r / Cout;
        wk_f[((size_t)ic * 9 + (8 - k)) * Cout + oc] = w[i];
        wk_b[((size_t)oc * 9 + k) * Cin + ic] = w[i];
    }
}

// all layers in ONE launch (the AE fine-tune refreshes 20 layers after every Adam step: 20 launches of ~4 us each were 5 % of a step)
constexpr int PREP_MAX_LAYERS = 24;
struct PrepTab {
    int n;
    long long start[PREP_MAX_LAYERS + 1];        // prefix sums of Cin*Cout*9
    long long w_off[PREP_MAX_LAYERS];
    int Cin[PREP_MAX_LAYERS], Cout[PREP_MAX_LAYERS], transposed[PREP_MAX_LAYERS];
    float *wk_f[PREP_MAX_LAYERS], *wk_b[PREP_MAX_LAYERS];
};
__global__ void k_prep_weights_all(const __grid_constant__ PrepTab t, const float* __restrict__ w_flat) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= t.start[t.n]) return;
    int l = 0;
    while (g >= t.start[l + 1]) ++l;
    const int i = (int)(g - t.start[l]), Cin = t.Cin[l], Cout = t.Cout[l];
    const float v = w_flat[t.w_off[l] + i];
    const int k = i % 9, r = i / 9;
    if (!t.transposed[l]) {
        const int ic = r % Cin, oc = r / Cin;
        t.wk_f[l][((size_t)ic * 9 + k) * Cout + oc] = v;
        t.wk_b[l][((size_t)oc * 9 + (8 - k)) * Cin + ic] = v;
    } else {
        const int oc = r % Cout, ic = r / Cout;
        t.wk_f[l][((size_t)ic * 9 + (8 - k)) * Cout + oc] = v;
        t.wk_b[l][((size_t)oc * 9 + k) * Cin + ic] = v;
    }
}

int convnet_refresh_weights(ConvNet* n, cudaStream_t st) {
    if ((int)n->layers.size() <= PREP_MAX_LAYERS) {
        PrepTab t{};
        t.n = (int)n->layers.size();
        long long pos = 0;
        for (int l = 0; l < t.n; ++l) {
            const ConvLayer& L = n->layers[l];
            t.start[l] = pos; pos += (long long)L.Cin * L.Cout * 9;
            t.w_off[l] = L.w_off; t.Cin[l] = L.Cin; t.Cout[l] = L.Cout; t.transposed[l] = L.transposed ? 1 : 0;
            t.wk_f[l] = L.wk_f; t.wk_b[l] = L.wk_b;
        }
        t.start[t.n] = pos;
        k_prep_weights_all<<<(unsigned)cdiv(pos, 256), 256, 0, st>>>(t, n->w_flat);
    } else {
        for (auto& L : n->layers) {
            const int tot = L.Cin * L.Cout * 9;
            k_prep_weights<<<cdiv(tot, 256), 256, 0, st>>>(n->w_flat + L.w_off, L.Cin, L.Cout, L.transposed ? 1 : 0, L.wk_f, L.wk_b);
        }
    }
    LEMO_CUDA(cudaGetLastError());
    if (n->tc) LEMO_TRY(enc_tc_refresh_weights(n, st));
    return 0;
}

template <typename T>
static int dalloc(T** p, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    LEMO_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}

int convnet_create(int kind, int in_ch, const float* h_weights, long long n_weights, int maxN, int H, int W, bool with_backward,
                   int device, ConvNet** out) {
    LEMO_CHECK(out && h_weights && maxN > 0 && H > 0 && W > 0, "bad arguments");
    LEMO_CHECK(kind == 0, "AE (kind 1) is built by ae.cu");
    LEMO_CUDA(cudaSetDevice(device));
    ConvNet* n = new ConvNet();
    n->device = device; n->kind = kind; n->in_ch = in_ch; n->maxN = maxN; n->with_backward = with_backward;
    const int chans[11] = {in_ch, 32, 32, 64, 64, 64, 64, 64, 64, 64, 64};      // AE_sep.py:77-89, z_channel=64
    long long off = 0;
    for (int l = 0; l < 10; ++l) {
        ConvLayer L;
        L.Cin = chans[l]; L.Cout = chans[l + 1]; L.transposed = false;
        L.w_off = off; off += (long long)L.Cin * L.Cout * 9;
        L.b_off = off; off += L.Cout;
        n->layers.push_back(L);
    }
    LEMO_CHECK(off == n_weights, "weight vector length does not match Enc(downsample=False, z_channel=64)");
    n->n_weights = off;
    LEMO_CUDA(cudaMalloc((void**)&n->w_flat, off * sizeof(float)));
    LEMO_CUDA(cudaMemcpy(n->w_flat, h_weights, off * sizeof(float), cudaMemcpyHostToDevice));
    for (auto& L : n->layers) {
        LEMO_TRY(dalloc(&L.wk_f, (size_t)L.Cin * L.Cout * 9));
        LEMO_TRY(dalloc(&L.wk_b, (size_t)L.Cin * L.Cout * 9));
    }
    LEMO_TRY(convnet_refresh_weights(n, 0));
    const PlaneGeom g = make_geom(H, W);
    n->geom.push_back(g);
    n->act.resize(11, nullptr);
    LEMO_TRY(dalloc(&n->act[0], (size_t)maxN * in_ch * g.PS));
    for (int l = 0; l < 10; ++l) LEMO_TRY(dalloc(&n->act[l + 1], (size_t)maxN * chans[l + 1] * g.PS));
    if (with_backward) {
        n->grad.resize(2, nullptr);
        LEMO_TRY(dalloc(&n->grad[0], (size_t)maxN * 64 * g.PS));
        LEMO_TRY(dalloc(&n->grad[1], (size_t)maxN * 64 * g.PS));
    }
    if (in_ch == 1) LEMO_TRY(enc_tc_create(n));          // tensor-core path (conv_tc.cu); LEMO_CONV=simt keeps it idle
    LEMO_CUDA(cudaDeviceSynchronize());
    *out = n;
    return 0;
}

void convnet_free(ConvNet* n) {
    if (!n) return;
    cudaSetDevice(n->device);
    if (n->tc) enc_tc_free(n);
    cudaFree(n->w_flat); cudaFree(n->d_wflat); cudaFree(n->sk_scratch); cudaFree(n->wg_scratch); cudaFree(n->ft_sched);
    if (n->ft_gexec) cudaGraphExecDestroy((cudaGraphExec_t)n->ft_gexec);
    if (n->ft_graph) cudaGraphDestroy((cudaGraph_t)n->ft_graph);
    cudaFree(n->wg_scratch2);
    for (int i = 0; i < 2; ++i) if (n->bw_side[i]) cudaStreamDestroy((cudaStream_t)n->bw_side[i]);
    if (n->bw_ready) cudaEventDestroy((cudaEvent_t)n->bw_ready);
    for (int i = 0; i < 2; ++i) if (n->bw_done[i]) cudaEventDestroy((cudaEvent_t)n->bw_done[i]);
    for (auto p : n->glev) cudaFree(p);
    if (n->ft_stream) { cudaStreamDestroy((cudaStream_t)n->ft_stream); cudaEventDestroy((cudaEvent_t)n->ft_ev_in); cudaEventDestroy((cudaEvent_t)n->ft_ev_out); }
    for (auto& L : n->layers) { cudaFree(L.wk_f); cudaFree(L.wk_b); }
    for (auto p : n->act) cudaFree(p);
    for (auto p : n->grad) cudaFree(p);
    for (auto p : n->pool_in) cudaFree(p);
    for (auto p : n->pool_idx) cudaFree(p);
    for (auto p : n->up) cudaFree(p);
    delete n;
}

int enc_forward_planes(ConvNet* n, const float* x_planes, int N, cudaStream_t st) {
    LEMO_CHECK(n && n->kind == 0 && N > 0 && N <= n->maxN, "bad Enc handle / batch exceeds handle size");
    if (enc_uses_tc(n)) return enc_tc_forward(n, x_planes, N, st);
    const PlaneGeom& g = n->geom[0];
    const float* cur = x_planes;
    for (int l = 0; l < 10; ++l) {
        const ConvLayer& L = n->layers[l];
        LEMO_TRY(conv3x3_launch(cur, L.wk_f, n->w_flat + L.b_off, nullptr, n->act[l + 1], N, L.Cin, L.Cout, g, EPI_BIAS_LRELU, st));
        cur = n->act[l + 1];
    }
    n->launches += 10;
    return 0;
}

int enc_backward_planes(ConvNet* n, int N, float* dx_planes, cudaStream_t st) {
    LEMO_CHECK(n && n->kind == 0 && n->with_backward && N > 0 && N <= n->maxN, "Enc handle has no backward buffers");
    if (enc_uses_tc(n)) return enc_tc_backward(n, N, dx_planes, st);
    const PlaneGeom& g = n->geom[0];
    float* cur = n->grad[0];
    float* nxt = n->grad[1];
    for (int l = 9; l >= 1; --l) {          // dpre_l -> dpre_{l-1} = convT(dpre_l) * LeakyReLU'(a_{l-1})
        const ConvLayer& L = n->layers[l];
        LEMO_TRY(conv3x3_launch(cur, L.wk_b, nullptr, n->act[l], nxt, N, L.Cout, L.Cin, g, EPI_MASK, st));
        std::swap(cur, nxt);
    }
    const ConvLayer& L0 = n->layers[0];
    LEMO_TRY(conv3x3_launch(cur, L0.wk_b, nullptr, nullptr, dx_planes, N, L0.Cout, L0.Cin, g, EPI_NONE, st));
    n->launches += 10;
    return 0;
}

}  // namespace lemo

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace lemo;
#include "handles.cuh"

namespace lemo { int ae_create(int in_ch, const float* h_weights, long long n_weights, int maxN, int H, int W, bool with_backward,
                               int device, ConvNet** out); }

extern "C" {
int lemo_convnet_create(int32_t kind, int32_t in_channels, const float* h_weights, int64_t n_weights, int32_t max_n, int32_t H,
                        int32_t W, int32_t with_backward, int device, LemoConvNet** out) {
    LEMO_CHECK(out, "null out");
    ConvNet* n = nullptr;
    if (kind == 0) LEMO_TRY(convnet_create(0, in_channels, h_weights, n_weights, max_n, H, W, with_backward != 0, device, &n));
    else if (kind == 1) LEMO_TRY(ae_create(in_channels, h_weights, n_weights, max_n, H, W, with_backward != 0, device, &n));
    else { set_error("unknown convnet kind"); return 2; }
    LemoConvNet* h = new LemoConvNet{n, nullptr};
    if (kind == 0 && with_backward) {
        LEMO_CUDA(cudaMalloc((void**)&h->dx_planes, (size_t)max_n * in_channels * n->geom[0].PS * sizeof(float)));
        LEMO_CUDA(cudaMemset(h->dx_planes, 0, (size_t)max_n * in_channels * n->geom[0].PS * sizeof(float)));
    }
    *out = h;
    return 0;
}
int lemo_convnet_destroy(LemoConvNet* h) {
    if (!h) return 0;
    cudaFree(h->dx_planes);
    convnet_free(h->n);
    delete h;
    return 0;
}
int64_t lemo_convnet_num_weights(const LemoConvNet* h) { return h ? h->n->n_weights : 0; }
int lemo_convnet_set_weights(LemoConvNet* h, const float* w, void* stream) {
    LEMO_CHECK(h && w, "bad arguments");
    LEMO_CUDA(cudaMemcpyAsync(h->n->w_flat, w, h->n->n_weights * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return convnet_refresh_weights(h->n, (cudaStream_t)stream);
}
int lemo_convnet_get_weights(LemoConvNet* h, float* w, void* stream) {
    LEMO_CHECK(h && w, "bad arguments");
    LEMO_CUDA(cudaMemcpyAsync(w, h->n->w_flat, h->n->n_weights * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
int lemo_enc_forward(LemoConvNet* h, const float* x, int32_t N, float* z, void* stream) {
    LEMO_NVTX("lemo_enc_forward");
    LEMO_CHECK(h && h->n->kind == 0 && x && z, "bad arguments");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    LEMO_TRY(pack_planes(x, n->act[0], N * n->in_ch, n->geom[0], st));
    LEMO_TRY(enc_forward_planes(n, n->act[0], N, st));
    if (enc_uses_tc(n)) return enc_tc_unpack_z(n, N, z, st);
    LEMO_TRY(unpack_planes(n->act[10], z, N * 64, n->geom[0], st));
    return 0;
}
// debug: dpre of layers[stop_layer] (gradient w.r.t. its pre-activation), dense [N,Cout,H,W]
int lemo_enc_debug_backward(LemoConvNet* h, const float* dz, int32_t N, int32_t stop_layer, float* out, void* stream) {
    LEMO_CHECK(h && h->n->kind == 0 && dz && out && stop_layer >= 0 && stop_layer <= 9, "bad arguments");
    LEMO_CHECK(!enc_uses_tc(h->n), "the layer-wise debug hook inspects the fp32 CUDA-core path: call lemo_debug_set_conv_tc(0) first");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    const PlaneGeom& g = n->geom[0];
    const long long total = (long long)N * 64 * g.H * g.W;
    k_pack_mask<<<cdiv(total, 256), 256, 0, st>>>(dz, n->act[10], n->grad[0], total, g.H, g.W, g.Wp, g.PS);
    float* cur = n->grad[0];
    float* nxt = n->grad[1];
    for (int l = 9; l > stop_layer; --l) {
        const ConvLayer& L = n->layers[l];
        LEMO_TRY(conv3x3_launch(cur, L.wk_b, nullptr, n->act[l], nxt, N, L.Cout, L.Cin, g, EPI_MASK, st));
        std::swap(cur, nxt);
    }
    return unpack_planes(cur, out, N * n->layers[stop_layer].Cout, g, st);
}
int lemo_debug_set_conv_tc(int32_t on) { conv_tc_set(on); return 0; }
int lemo_convnet_profile_layer(LemoConvNet* h, int32_t layer, int32_t N, int32_t backward, int32_t reps, void* stream) {
    LEMO_CHECK(h && h->n->kind == 0 && layer >= 0 && layer < 10 && N > 0 && N <= h->n->maxN, "bad arguments");
    ConvNet* n = h->n;
    if (enc_uses_tc(n)) return enc_tc_profile_layer(n, layer, N, backward, reps, (cudaStream_t)stream);
    const ConvLayer& L = n->layers[layer];
    for (int r = 0; r < reps; ++r) {
        if (!backward) {
            LEMO_TRY(conv3x3_launch(n->act[layer], L.wk_f, n->w_flat + L.b_off, nullptr, n->act[layer + 1], N, L.Cin, L.Cout, n->geom[0],
                                    EPI_BIAS_LRELU, (cudaStream_t)stream));
        } else {
            LEMO_CHECK(n->with_backward && layer >= 1, "backward profiling needs backward buffers and layer >= 1");
            LEMO_TRY(conv3x3_launch(n->grad[0], L.wk_b, nullptr, n->act[layer], n->grad[1], N, L.Cout, L.Cin, n->geom[0], EPI_MASK,
                                    (cudaStream_t)stream));
        }
    }
    return 0;
}
int lemo_enc_backward_input(LemoConvNet* h, const float* dz, int32_t N, float* dx, void* stream) {
    LEMO_NVTX("lemo_enc_backward_input");
    LEMO_CHECK(h && h->n->kind == 0 && dz && dx && h->dx_planes, "bad arguments / handle created without backward");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    const PlaneGeom& g = n->geom[0];
    const long long total = (long long)N * 64 * g.H * g.W;
    if (enc_uses_tc(n)) LEMO_TRY(enc_tc_pack_dz(n, N, dz, st));
    else k_pack_mask<<<cdiv(total, 256), 256, 0, st>>>(dz, n->act[10], n->grad[0], total, g.H, g.W, g.Wp, g.PS);
    LEMO_CUDA(cudaGetLastError());
    LEMO_TRY(enc_backward_planes(n, N, h->dx_planes, st));
    LEMO_TRY(unpack_planes(h->dx_planes, dx, N * n->in_ch, g, st));
    return 0;
}
}
