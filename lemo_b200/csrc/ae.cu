// AE -- the motion-infilling prior (reference models/AE.py:11-108, AE(downsample=True, in_channel=4, kernel=3)):
//   5x [conv3x3+LeakyReLU, conv3x3+LeakyReLU, MaxPool(3,s2,p1)]  ->  z [N,256,7,5]
//   5x [ConvTranspose(s2, output_size)+LeakyReLU, ConvTranspose(s1)(+LeakyReLU except the last)]  ->  rec [N,1,H,W]
// plus the weight-gradient backward used by the 60-step self-supervised fine-tune
// (opt_amass_perframe.py:152-173, opt_amass_temp.py:152-196; PROX S3: fitting_temp_slide.py:868-885).
// Everything stays on the padded pitch-linear planes of conv.cuh, one geometry per resolution level; a stride-2
// ConvTranspose is a zero-upsample (value at (2y,2x)) followed by the stride-1 kernel with flipped taps.
#include "handles.cuh"
#include "fit_common.cuh"
#include "../../include/lemo_b200.h"
#include <algorithm>
#include <mutex>

namespace lemo {

// ---------------------------------------------------------------------------------------------- pooling / upsampling
// MaxPool2d(3, stride 2, pad 1) on planes; idx = linear index (within the source plane) of the arg-max (first max wins)
__global__ void k_maxpool_fwd(const float* __restrict__ src, int NC, PlaneGeom gs, PlaneGeom gd, float* __restrict__ dst, int* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)NC * gd.H * gd.W) return;
    const int ox = (int)(i % gd.W), oy = (int)((i / gd.W) % gd.H);
    const long long c = i / ((long long)gd.W * gd.H);
    const float* s = src + c * gs.PS;
    float best = -3.4e38f;
    int bi = -1;
    for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * oy + dy;
        if (y < 0 || y >= gs.H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int x = 2 * ox + dx;
            if (x < 0 || x >= gs.W) continue;
            const int q = (y + 1) * gs.Wp + x + 1;
            const float v = s[q];
            if (v > best || bi < 0) { best = v; bi = q; }
        }
    }
    const long long o = c * gd.PS + (oy + 1) * gd.Wp + ox + 1;
    dst[o] = best; idx[o] = bi;
}
// gather-form adjoint (no atomics): pre-pool pixel (y,x) collects from the <=4 windows that contain it; result is
// multiplied by LeakyReLU'(pre-pool activation) so it is directly the dpre of the conv that produced it.
__global__ void k_maxpool_bwd_mask(const float* __restrict__ gdst, const int* __restrict__ idx, const float* __restrict__ act, int NC,
                                   PlaneGeom gs, PlaneGeom gd, float* __restrict__ gsrc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)NC * gs.H * gs.W) return;
    const int x = (int)(i % gs.W), y = (int)((i / gs.W) % gs.H);
    const long long c = i / ((long long)gs.W * gs.H);
    const int q = (y + 1) * gs.Wp + x + 1;
    float a = 0.f;
    for (int oy = (y) / 2; oy <= (y + 1) / 2; ++oy) {          // windows with 2*oy-1 <= y <= 2*oy+1
        if (oy >= gd.H) continue;
        for (int ox = (x) / 2; ox <= (x + 1) / 2; ++ox) {
            if (ox >= gd.W) continue;
            const long long o = c * gd.PS + (oy + 1) * gd.Wp + ox + 1;
            if (idx[o] == q) a += gdst[o];
        }
    }
    gsrc[c * gs.PS + q] = a * (act[c * gs.PS + q] > 0.f ? 1.f : 0.2f);
}
// zero-upsample: dst(2y,2x) = src(y,x), everything else 0  (dst planes are zero-filled first)
__global__ void k_upsample_fwd(const float* __restrict__ src, int NC, PlaneGeom gs, PlaneGeom gd, float* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)NC * gs.H * gs.W) return;
    const int x = (int)(i % gs.W), y = (int)((i / gs.W) % gs.H);
    const long long c = i / ((long long)gs.W * gs.H);
    dst[c * gd.PS + (2 * y + 1) * gd.Wp + 2 * x + 1] = src[c * gs.PS + (y + 1) * gs.Wp + x + 1];
}
// adjoint of the zero-upsample, fused with LeakyReLU' of the tensor that was upsampled (mask == nullptr: no activation)
__global__ void k_upsample_bwd_mask(const float* __restrict__ gup, const float* __restrict__ act, int NC, PlaneGeom gs, PlaneGeom gd,
                                    float* __restrict__ gsrc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)NC * gs.H * gs.W) return;
    const int x = (int)(i % gs.W), y = (int)((i / gs.W) % gs.H);
    const long long c = i / ((long long)gs.W * gs.H);
    const long long q = c * gs.PS + (y + 1) * gs.Wp + x + 1;
    const float g = gup[c * gd.PS + (2 * y + 1) * gd.Wp + 2 * x + 1];
    gsrc[q] = act ? g * (act[q] > 0.f ? 1.f : 0.2f) : g;
}
__global__ void k_mask_inplace(float* __restrict__ g, const float* __restrict__ act, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] *= act[i] > 0.f ? 1.f : 0.2f;
}

// ---------------------------------------------------------------------------------------------- weight gradient
// G[oc][ic][k] = sum_{n,q} dpre[n][oc][q] * x[n][ic][q + off_k];  db[oc] = sum dpre.   16x16 (oc,ic) tile per CTA, the linear pixel
// range cut into 256-pixel chunks staged in shared memory.  Deterministic: the chunks are dealt to NG groups (blockIdx.z = n*NG + g),
// a CTA accumulates its chunks in registers and stores ONE partial per (group, weight); k_wgrad_finish adds the N*NG partials in
// order and writes the gradient in the state_dict layout (the first version atomicAdd-ed every chunk: order-dependent rounding).
constexpr int AE_GB = 6;            // gradient buffers per resolution level (most tensors one backward pass leaves on a level)
constexpr int WG_T = 16, WG_CP = 256, WG_DP = WG_CP + 4;      // WG_DP: pitch of the staged dpre rows (16-byte aligned rows)
__global__ void __launch_bounds__(256) k_wgrad(const float* __restrict__ x, const float* __restrict__ dpre, float* __restrict__ part,
                                               int Cin, int Cout, int H, int Wp, int PS, int nchunks, int NG, int SWX) {
    extern __shared__ __align__(16) float sm[];
    float* s_d = sm;                               // [16][WG_DP]
    float* s_x = sm + WG_T * WG_DP;                // [16][SWX]
    const int o = threadIdx.x & 15, i = threadIdx.x >> 4;
    const int oc0 = blockIdx.x * WG_T, ic0 = blockIdx.y * WG_T;
    const int n = blockIdx.z / NG, g = blockIdx.z - n * NG;
    const int qend = (H + 1) * Wp;
    const int span = WG_CP + 2 * Wp + 2;
    float acc[9], bsum = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
    for (int ch = g; ch < nchunks; ch += NG) {
        const int q0 = Wp + ch * WG_CP;
        __syncthreads();
        for (int e = threadIdx.x; e < WG_T * WG_CP; e += 256) {
            const int r = e / WG_CP, p = e % WG_CP;
            const int q = q0 + p;
            s_d[r * WG_DP + p] = (oc0 + r < Cout && q < qend) ? dpre[((size_t)n * Cout + oc0 + r) * PS + q] : 0.f;
        }
        for (int e = threadIdx.x; e < WG_T * span; e += 256) {
            const int r = e / span, p = e % span;
            const int q = q0 - Wp - 1 + p;
            s_x[r * SWX + p] = (ic0 + r < Cin && q >= 0 && q < PS) ? x[((size_t)n * Cin + ic0 + r) * PS + q] : 0.f;
        }
        __syncthreads();
        // four pixels per step: one 128-bit read of dpre and, per tap row, a 128-bit + 64-bit read of the input window feed 36 FMAs
        // (the scalar version issued 10 shared-memory reads per 9 FMAs and was load-issue bound: 30 us per launch).  Wp is a multiple
        // of 8 and the rows are 16-byte aligned, so p + ky*Wp stays a multiple of 4.
        const float* dr = s_d + o * WG_DP;
        const float* xr = s_x + i * SWX;
        for (int p = 0; p < WG_CP; p += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(dr + p);
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
            bsum += (d4.x + d4.y) + (d4.z + d4.w);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float4 x4 = *reinterpret_cast<const float4*>(xr + p + ky * Wp);
                const float2 x2 = *reinterpret_cast<const float2*>(xr + p + ky * Wp + 4);
                const float xv[6] = {x4.x, x4.y, x4.z, x4.w, x2.x, x2.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[ky * 3 + kx] = fmaf(dv[u], xv[u + kx], acc[ky * 3 + kx]);
            }
        }
    }
    const int oc = oc0 + o, ic = ic0 + i;
    float* pz = part + (size_t)blockIdx.z * ((size_t)Cout * Cin * 9 + Cout);
    if (oc < Cout && ic < Cin) {
#pragma unroll
        for (int k = 0; k < 9; ++k) pz[((size_t)oc * Cin + ic) * 9 + k] = acc[k];
    }
    if (blockIdx.y == 0 && i == 0 && oc < Cout) pz[(size_t)Cout * Cin * 9 + oc] = bsum;
}
// dW (state_dict layout) = sum of the partials in (n, group) order; Conv2d weight [oc][ic][k]; ConvTranspose2d weight
// [ic_t = our ic][oc_t = our oc][8-k]
__global__ void k_wgrad_finish(const float* __restrict__ part, int nparts, int Cin, int Cout, int transposed, float* __restrict__ dW,
                               float* __restrict__ db) {
    const int nw = Cout * Cin * 9;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nw + Cout) return;
    float a = 0.f;
    for (int p = 0; p < nparts; ++p) a += part[(size_t)p * (nw + Cout) + e];
    if (e >= nw) { db[e - nw] = a; return; }
    const int k = e % 9, ic = (e / 9) % Cin, oc = e / (9 * Cin);
    dW[transposed ? ((size_t)ic * Cout + oc) * 9 + (8 - k) : (size_t)e] = a;
}

static int wgrad_groups(int N, int Cin, int Cout, const PlaneGeom& g) {
    const int nchunks = cdiv((long long)g.H * g.Wp, WG_CP);
    const long long tiles = (long long)cdiv(Cout, WG_T) * cdiv(Cin, WG_T) * N;
    return (int)std::max(1LL, std::min<long long>(nchunks, (296 + tiles - 1) / tiles));
}
size_t conv3x3_wgrad_floats(int N, int Cin, int Cout, const PlaneGeom& g) {
    return (size_t)N * wgrad_groups(N, Cin, Cout, g) * ((size_t)Cout * Cin * 9 + Cout);
}

int conv3x3_wgrad_launch(const float* x, const float* dpre, float* dW, float* db, int N, int Cin, int Cout, const PlaneGeom& g,
                         bool transposed, cudaStream_t st, float* scratch, size_t scratch_floats) {
    LEMO_CHECK(scratch && scratch_floats >= conv3x3_wgrad_floats(N, Cin, Cout, g), "weight-gradient scratch too small");
    const int nchunks = cdiv((long long)g.H * g.Wp, WG_CP);
    const int NG = wgrad_groups(N, Cin, Cout, g);
    const int SWX = (WG_CP + 2 * g.Wp + 2 + 4 + 3) / 4 * 4;          // window + the 2 floats the last vector read runs over, 16-byte pitch
    const size_t smem = (size_t)(WG_T * WG_DP + WG_T * SWX) * sizeof(float);
    {                                             // monotonic under a lock: several host threads launch (see conv_main_launch)
        static std::mutex mu;
        static size_t configured = 0;
        std::lock_guard<std::mutex> lk(mu);
        if (smem > configured) {
            LEMO_CUDA(cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured = smem;
        }
    }
    dim3 grid(cdiv(Cout, WG_T), cdiv(Cin, WG_T), N * NG);
    k_wgrad<<<grid, 256, smem, st>>>(x, dpre, scratch, Cin, Cout, g.H, g.Wp, g.PS, nchunks, NG, SWX);
    k_wgrad_finish<<<cdiv(Cout * Cin * 9 + Cout, 256), 256, 0, st>>>(scratch, N * NG, Cin, Cout, transposed ? 1 : 0, dW, db);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------- network
// buffer map for kind==1 (index into ConvNet::act): 0 = input; encoder level i (0..4): 1+3i = after conv1, 2+3i = after conv2
// (pre-pool), 3+3i = pooled (level i+1);  decoder block b (0..4): 16+3b = upsampled input, 17+3b = after deconv1, 18+3b = after deconv2
static inline int E1(int i) { return 1 + 3 * i; }
static inline int E2(int i) { return 2 + 3 * i; }
static inline int EP(int i) { return 3 + 3 * i; }
static inline int DU(int b) { return 16 + 3 * b; }
static inline int D1(int b) { return 17 + 3 * b; }
static inline int D2(int b) { return 18 + 3 * b; }

int ae_create(int in_ch, const float* h_weights, long long n_weights, int maxN, int H, int W, bool with_backward, int device, ConvNet** out) {
    LEMO_CHECK(out && h_weights && maxN > 0 && H >= 32 && W >= 32, "bad arguments (AE needs H,W >= 32 for five 2x poolings)");
    LEMO_CUDA(cudaSetDevice(device));
    ConvNet* n = new ConvNet();
    n->device = device; n->kind = 1; n->in_ch = in_ch; n->maxN = maxN; n->with_backward = with_backward;
    const int ec[6] = {in_ch, 32, 64, 128, 256, 256};       // models/AE.py:82-86
    const int dc[6] = {256, 256, 128, 64, 32, 1};           // models/AE.py:88-92
    long long off = 0;
    auto add = [&](int ci, int co, bool tr) {
        ConvLayer L;
        L.Cin = ci; L.Cout = co; L.transposed = tr;
        L.w_off = off; off += (long long)ci * co * 9;
        L.b_off = off; off += co;
        n->layers.push_back(L);
    };
    for (int i = 0; i < 5; ++i) { add(ec[i], ec[i + 1], false); add(ec[i + 1], ec[i + 1], false); }
    for (int b = 0; b < 5; ++b) { add(dc[b], dc[b + 1], true); add(dc[b + 1], dc[b + 1], true); }
    LEMO_CHECK(off == n_weights, "weight vector length does not match AE(downsample=True, in_channel=C, kernel=3)");
    n->n_weights = off;
    LEMO_CUDA(cudaMalloc((void**)&n->w_flat, off * sizeof(float)));
    LEMO_CUDA(cudaMemcpy(n->w_flat, h_weights, off * sizeof(float), cudaMemcpyHostToDevice));
    for (auto& L : n->layers) {
        LEMO_TRY(dalloc(&L.wk_f, (size_t)L.Cin * L.Cout * 9));
        LEMO_TRY(dalloc(&L.wk_b, (size_t)L.Cin * L.Cout * 9));
    }
    LEMO_TRY(convnet_refresh_weights(n, 0));
    int h = H, w = W;
    for (int l = 0; l < 6; ++l) { n->geom.push_back(make_geom(h, w)); h = (h - 1) / 2 + 1; w = (w - 1) / 2 + 1; }
    n->act.resize(31, nullptr);
    n->pool_idx.resize(5, nullptr);
    const size_t N = maxN;
    size_t gmax = 0;
    LEMO_TRY(dalloc(&n->act[0], N * in_ch * n->geom[0].PS));
    for (int i = 0; i < 5; ++i) {
        const size_t c = ec[i + 1];
        LEMO_TRY(dalloc(&n->act[E1(i)], N * c * n->geom[i].PS));
        LEMO_TRY(dalloc(&n->act[E2(i)], N * c * n->geom[i].PS));
        LEMO_TRY(dalloc(&n->act[EP(i)], N * c * n->geom[i + 1].PS));
        LEMO_TRY(dalloc(&n->pool_idx[i], N * c * n->geom[i + 1].PS));
        gmax = std::max(gmax, N * c * n->geom[i].PS);
    }
    for (int b = 0; b < 5; ++b) {
        const PlaneGeom& g = n->geom[4 - b];
        LEMO_TRY(dalloc(&n->act[DU(b)], N * dc[b] * g.PS));
        LEMO_TRY(dalloc(&n->act[D1(b)], N * dc[b + 1] * g.PS));
        LEMO_TRY(dalloc(&n->act[D2(b)], N * dc[b + 1] * g.PS));
        gmax = std::max(gmax, N * (size_t)std::max(dc[b], dc[b + 1]) * g.PS);
    }
    // workspaces: split-K partials of the small-plane convolutions (forward and input gradient) and weight-gradient partials
    size_t sk = 0, wg = 0;
    for (int i = 0; i < 5; ++i) {
        const PlaneGeom& g = n->geom[i];
        for (int l = 0; l < 2; ++l) {
            const ConvLayer& L = n->layers[2 * i + l];
            sk = std::max(sk, std::max(conv3x3_splitk_floats(maxN, L.Cin, L.Cout, g), conv3x3_splitk_floats(maxN, L.Cout, L.Cin, g)));
            wg = std::max(wg, conv3x3_wgrad_floats(maxN, L.Cin, L.Cout, g));
        }
    }
    for (int b = 0; b < 5; ++b) {
        const PlaneGeom& g = n->geom[4 - b];
        for (int l = 0; l < 2; ++l) {
            const ConvLayer& L = n->layers[10 + 2 * b + l];
            sk = std::max(sk, std::max(conv3x3_splitk_floats(maxN, L.Cin, L.Cout, g), conv3x3_splitk_floats(maxN, L.Cout, L.Cin, g)));
            wg = std::max(wg, conv3x3_wgrad_floats(maxN, L.Cin, L.Cout, g));
        }
    }
    if (sk) { LEMO_TRY(dalloc(&n->sk_scratch, sk)); n->sk_floats = sk; }
    if (with_backward) {
        n->grad.resize(1, nullptr);
        LEMO_TRY(dalloc(&n->grad[0], N * n->geom[0].PS));            // dpre of the last layer (1 channel, level 0)
        const int cmax[6] = {32, 64, 128, 256, 256, 256};            // widest gradient tensor per level
        n->glev.resize(6 * AE_GB, nullptr);
        for (int l = 0; l < 6; ++l)
            for (int k = 0; k < AE_GB; ++k) LEMO_TRY(dalloc(&n->glev[l * AE_GB + k], N * cmax[l] * n->geom[l].PS));
        LEMO_TRY(dalloc(&n->wg_scratch, wg)); n->wg_floats = wg;
        LEMO_TRY(dalloc(&n->wg_scratch2, wg));
        for (int i = 0; i < 2; ++i) { cudaStream_t s; LEMO_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); n->bw_side[i] = s; }
        for (int i = 0; i < 3; ++i) {
            cudaEvent_t e; LEMO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            if (i < 2) n->bw_done[i] = e; else n->bw_ready = e;
        }
    }
    LEMO_CUDA(cudaDeviceSynchronize());
    *out = n;
    return 0;
}

static int ae_forward_planes(ConvNet* n, int N, cudaStream_t st) {
    const int ec[6] = {n->in_ch, 32, 64, 128, 256, 256}, dc[6] = {256, 256, 128, 64, 32, 1};
    const float* cur = n->act[0];
    for (int i = 0; i < 5; ++i) {
        const PlaneGeom& g = n->geom[i];
        const ConvLayer &A = n->layers[2 * i], &B = n->layers[2 * i + 1];
        LEMO_TRY(conv3x3_launch(cur, A.wk_f, n->w_flat + A.b_off, nullptr, n->act[E1(i)], N, A.Cin, A.Cout, g, EPI_BIAS_LRELU, st, n->sk_scratch, n->sk_floats));
        LEMO_TRY(conv3x3_launch(n->act[E1(i)], B.wk_f, n->w_flat + B.b_off, nullptr, n->act[E2(i)], N, B.Cin, B.Cout, g, EPI_BIAS_LRELU, st, n->sk_scratch, n->sk_floats));
        const PlaneGeom& gd = n->geom[i + 1];
        const long long tot = (long long)N * ec[i + 1] * gd.H * gd.W;
        k_maxpool_fwd<<<cdiv(tot, 256), 256, 0, st>>>(n->act[E2(i)], N * ec[i + 1], g, gd, n->act[EP(i)], n->pool_idx[i]);
        cur = n->act[EP(i)];
    }
    for (int b = 0; b < 5; ++b) {
        const PlaneGeom &gs = n->geom[5 - b], &gd = n->geom[4 - b];
        const ConvLayer &A = n->layers[10 + 2 * b], &B = n->layers[11 + 2 * b];
        // (act[DU(b)] was zeroed at create; only the (2y, 2x) positions are ever written, so the zeros in between persist)
        const long long tot = (long long)N * dc[b] * gs.H * gs.W;
        k_upsample_fwd<<<cdiv(tot, 256), 256, 0, st>>>(cur, N * dc[b], gs, gd, n->act[DU(b)]);
        LEMO_TRY(conv3x3_launch(n->act[DU(b)], A.wk_f, n->w_flat + A.b_off, nullptr, n->act[D1(b)], N, A.Cin, A.Cout, gd, EPI_BIAS_LRELU, st, n->sk_scratch, n->sk_floats));
        LEMO_TRY(conv3x3_launch(n->act[D1(b)], B.wk_f, n->w_flat + B.b_off, nullptr, n->act[D2(b)], N, B.Cin, B.Cout, gd,
                                b < 4 ? EPI_BIAS_LRELU : EPI_BIAS, st, n->sk_scratch, n->sk_floats));
        cur = n->act[D2(b)];
    }
    LEMO_CUDA(cudaGetLastError());
    n->launches += 45;
    return 0;
}

// d_rec already packed into grad[0] (1 channel at level 0); writes d_weights (flat, state_dict order; every element is written).
// The input-gradient chain runs on `st`; the weight gradient of each layer is forked onto one of two side streams (alternating, each with
// its own partials scratch) as soon as that layer's dpre exists, and both are joined at the end.  Every tensor of the chain gets its own
// buffer from a per-level pool (AE_GB per resolution level): nothing is overwritten within a pass, so the side streams need no
// intermediate joins, and -- the buffers being zeroed once and producers writing plane interiors only -- the zero-border invariant of
// the plane layout holds without the ~35 per-step clears the two-buffer version needed.  Inside the captured step these are graph branches.
static int ae_backward_planes(ConvNet* n, int N, float* dW, cudaStream_t st) {
    const int ec[6] = {n->in_ch, 32, 64, 128, 256, 256}, dc[6] = {256, 256, 128, 64, 32, 1};
    int ring[6] = {0, 0, 0, 0, 0, 0};
    int forks = 0;
    bool used[2] = {false, false};
    cudaEvent_t ready = (cudaEvent_t)n->bw_ready;
    auto out_buf = [&](int level) -> float* { return ring[level] < AE_GB ? n->glev[level * AE_GB + ring[level]++] : nullptr; };
    auto wgrad = [&](const float* x, const float* dpre, const ConvLayer& L, const PlaneGeom& g, bool transposed) -> int {
        const int k = forks++ & 1;
        cudaStream_t s = (cudaStream_t)n->bw_side[k];
        LEMO_CUDA(cudaEventRecord(ready, st));
        LEMO_CUDA(cudaStreamWaitEvent(s, ready, 0));
        LEMO_TRY(conv3x3_wgrad_launch(x, dpre, dW + L.w_off, dW + L.b_off, N, L.Cin, L.Cout, g, transposed, s,
                                      k ? n->wg_scratch2 : n->wg_scratch, n->wg_floats));
        used[k] = true;
        return 0;
    };
    const float* cur = n->grad[0];    // dpre of the layer being processed
    float* nxt;
    for (int b = 4; b >= 0; --b) {
        const PlaneGeom &gs = n->geom[5 - b], &gd = n->geom[4 - b];
        const ConvLayer &A = n->layers[10 + 2 * b], &B = n->layers[11 + 2 * b];
        // deconv2: dpre in cur (for b<4 it already carries LeakyReLU'(D2))
        LEMO_TRY(wgrad(n->act[D1(b)], cur, B, gd, true));
        LEMO_CHECK((nxt = out_buf(4 - b)) != nullptr, "gradient buffer pool exhausted");
        LEMO_TRY(conv3x3_launch(cur, B.wk_b, nullptr, n->act[D1(b)], nxt, N, B.Cout, B.Cin, gd, EPI_MASK, st, n->sk_scratch, n->sk_floats));
        cur = nxt;
        // deconv1
        LEMO_TRY(wgrad(n->act[DU(b)], cur, A, gd, true));
        LEMO_CHECK((nxt = out_buf(4 - b)) != nullptr, "gradient buffer pool exhausted");
        LEMO_TRY(conv3x3_launch(cur, A.wk_b, nullptr, nullptr, nxt, N, A.Cout, A.Cin, gd, EPI_NONE, st, n->sk_scratch, n->sk_floats));
        cur = nxt;
        // through the zero-upsample to the tensor that fed this block: D2(b-1) (LeakyReLU output) or the pooled code z
        const long long tot = (long long)N * dc[b] * gs.H * gs.W;
        const float* mask = b > 0 ? n->act[D2(b - 1)] : nullptr;
        LEMO_CHECK((nxt = out_buf(5 - b)) != nullptr, "gradient buffer pool exhausted");
        k_upsample_bwd_mask<<<cdiv(tot, 256), 256, 0, st>>>(cur, mask, N * dc[b], gs, gd, nxt);
        cur = nxt;
    }
    // cur = dL/dz on level-5 planes (pooled output of encoder level 4)
    for (int i = 4; i >= 0; --i) {
        const PlaneGeom &g = n->geom[i], &gd = n->geom[i + 1];
        const ConvLayer &A = n->layers[2 * i], &B = n->layers[2 * i + 1];
        const long long tot = (long long)N * ec[i + 1] * g.H * g.W;
        LEMO_CHECK((nxt = out_buf(i)) != nullptr, "gradient buffer pool exhausted");
        k_maxpool_bwd_mask<<<cdiv(tot, 256), 256, 0, st>>>(cur, n->pool_idx[i], n->act[E2(i)], N * ec[i + 1], g, gd, nxt);
        cur = nxt;                    // dpre of conv2 at level i
        LEMO_TRY(wgrad(n->act[E1(i)], cur, B, g, false));
        LEMO_CHECK((nxt = out_buf(i)) != nullptr, "gradient buffer pool exhausted");
        LEMO_TRY(conv3x3_launch(cur, B.wk_b, nullptr, n->act[E1(i)], nxt, N, B.Cout, B.Cin, g, EPI_MASK, st, n->sk_scratch, n->sk_floats));
        cur = nxt;                    // dpre of conv1 at level i
        const float* xin = i > 0 ? n->act[EP(i - 1)] : n->act[0];
        LEMO_TRY(wgrad(xin, cur, A, g, false));
        if (i > 0) {
            LEMO_CHECK((nxt = out_buf(i)) != nullptr, "gradient buffer pool exhausted");
            LEMO_TRY(conv3x3_launch(cur, A.wk_b, nullptr, nullptr, nxt, N, A.Cout, A.Cin, g, EPI_NONE, st, n->sk_scratch, n->sk_floats));
            cur = nxt;                // dL/d(pooled output of level i-1), on level-i planes
        }
    }
    for (int k = 0; k < 2; ++k)       // join: every weight gradient is complete before the caller's next operation on `st`
        if (used[k]) {
            LEMO_CUDA(cudaEventRecord((cudaEvent_t)n->bw_done[k], (cudaStream_t)n->bw_side[k]));
            LEMO_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)n->bw_done[k], 0));
        }
    LEMO_CUDA(cudaGetLastError());
    n->launches += 64;
    return 0;
}

// masked L1 of the fine-tune (opt_amass_perframe.py:162-171): res = rec[:,0] - x[:,0] on the selected rows; writes
// d_rec = sign(res)/count into 1-channel level-0 planes and accumulates the loss value
__global__ void __launch_bounds__(256) k_ae_l1(const float* __restrict__ rec_planes, const float* __restrict__ x_planes, const float* __restrict__ row_mask,
                                               int N, int C, PlaneGeom g, float inv_count, float* __restrict__ d_planes, float* __restrict__ loss) {
    __shared__ float sred[32];
    float part = 0.f;
    const long long tot = (long long)N * g.H * g.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), n = (int)(i / ((long long)g.W * g.H));
        const int q = (y + 1) * g.Wp + x + 1;
        float d = 0.f;
        if (row_mask[y] != 0.f) {
            const float r = rec_planes[(size_t)n * g.PS + q] - x_planes[(size_t)n * C * g.PS + q];
            part += fabsf(r);
            d = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * inv_count;
        }
        d_planes[(size_t)n * g.PS + q] = d;
    }
    part = block_sum(part, sred);
    if (threadIdx.x == 0 && loss) atomicAdd(loss, part * inv_count);
}
// k_ae_l1 for the graph-replayed driver: the loss of step t lands in losses[t-1] (k_sched has already advanced the step counter)
__global__ void __launch_bounds__(256) k_ae_l1_sched(const float* __restrict__ rec_planes, const float* __restrict__ x_planes,
                                                     const float* __restrict__ row_mask, int N, int C, PlaneGeom g,
                                                     float* __restrict__ d_planes, float* __restrict__ losses, const Sched* __restrict__ sc) {
    __shared__ float sred[32];
    const float inv_count = sc->aux;                       // 1 / (N * selected rows * W): a per-run value, read from the schedule
    float part = 0.f;
    const long long tot = (long long)N * g.H * g.W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.W), y = (int)((i / g.W) % g.H), n = (int)(i / ((long long)g.W * g.H));
        const int q = (y + 1) * g.Wp + x + 1;
        float d = 0.f;
        if (row_mask[y] != 0.f) {
            const float r = rec_planes[(size_t)n * g.PS + q] - x_planes[(size_t)n * C * g.PS + q];
            part += fabsf(r);
            d = (r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f)) * inv_count;
        }
        d_planes[(size_t)n * g.PS + q] = d;
    }
    part = block_sum(part, sred);
    if (threadIdx.x == 0 && losses) atomicAdd(&losses[sc->it - 1], part * inv_count);
}
__global__ void k_adam_flat(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                            float b1, float b2, float eps, float bc1, float bc2s) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (lr / bc1) * (mi / (sqrtf(vi) / bc2s + eps));
}

}  // namespace lemo

using namespace lemo;
extern "C" {

int lemo_ae_finetune_step(LemoConvNet* h, const float* x, const float* row_mask, int32_t n_rows_selected, int32_t N, double lr, int32_t t,
                          float* loss_out, void* stream) {
    LEMO_NVTX("lemo_ae_finetune_step");
    LEMO_CHECK(h && h->n->kind == 1 && x && row_mask && h->n->with_backward && t >= 1 && n_rows_selected > 0, "bad arguments");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    const PlaneGeom& g = n->geom[0];
    if (!n->d_wflat) {            // first fine-tune step on this handle: gradient + Adam moments
        LEMO_CUDA(cudaMalloc((void**)&n->d_wflat, 3 * n->n_weights * sizeof(float)));
    }
    float* dW = n->d_wflat;
    float* m1 = dW + n->n_weights;
    float* m2 = m1 + n->n_weights;
    if (t == 1) LEMO_CUDA(cudaMemsetAsync(m1, 0, 2 * n->n_weights * sizeof(float), st));   // fresh optim.Adam (opt_amass_perframe.py:127-129)
    LEMO_TRY(pack_planes(x, n->act[0], N * n->in_ch, g, st));
    LEMO_TRY(ae_forward_planes(n, N, st));
    if (loss_out) LEMO_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(n->grad[0], 0, (size_t)N * g.PS * sizeof(float), st));
    const float inv_count = 1.f / ((float)N * (float)n_rows_selected * (float)g.W);
    k_ae_l1<<<64, 256, 0, st>>>(n->act[18 + 3 * 4], n->act[0], row_mask, N, n->in_ch, g, inv_count, n->grad[0], loss_out);
    LEMO_TRY(ae_backward_planes(n, N, dW, st));
    const float bc1 = (float)(1.0 - pow(0.9, (double)t)), bc2s = (float)sqrt(1.0 - pow(0.999, (double)t));
    k_adam_flat<<<cdiv(n->n_weights, 256), 256, 0, st>>>(n->w_flat, dW, m1, m2, n->n_weights, (float)lr, 0.9f, 0.999f, 1e-8f, bc1, bc2s);
    LEMO_CUDA(cudaGetLastError());
    return convnet_refresh_weights(n, st);
}

// one fine-tune step with the step-dependent scalars on the device (replayable)
static int ae_finetune_graph_step(ConvNet* n, cudaStream_t st) {
    const PlaneGeom& g = n->geom[0];
    const int N = n->ft_N;
    float* dW = n->d_wflat;
    float* m1 = dW + n->n_weights;
    float* m2 = m1 + n->n_weights;
    k_sched<<<1, 1, 0, st>>>((Sched*)n->ft_sched);
    LEMO_TRY(ae_forward_planes(n, N, st));
    LEMO_CUDA(cudaMemsetAsync(n->grad[0], 0, (size_t)N * g.PS * sizeof(float), st));
    k_ae_l1_sched<<<64, 256, 0, st>>>(n->act[18 + 3 * 4], n->act[0], (const float*)n->ft_mask, N, n->in_ch, g, n->grad[0],
                                     n->ft_losses, (const Sched*)n->ft_sched);
    LEMO_TRY(ae_backward_planes(n, N, dW, st));
    k_adam_dev<<<cdiv(n->n_weights, 256), 256, 0, st>>>(n->w_flat, dW, m1, m2, (int)n->n_weights, (const Sched*)n->ft_sched);
    LEMO_CUDA(cudaGetLastError());
    return convnet_refresh_weights(n, st);
}

int lemo_ae_finetune_run(LemoConvNet* h, const float* x, const float* row_mask, int32_t n_rows_selected, int32_t N, double lr, int32_t steps,
                         float* losses_out, void* stream) {
    LEMO_NVTX("lemo_ae_finetune_run");
    LEMO_CHECK(h && h->n->kind == 1 && x && row_mask && h->n->with_backward && steps >= 0 && n_rows_selected > 0, "bad arguments");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    if (!n->d_wflat) LEMO_CUDA(cudaMalloc((void**)&n->d_wflat, 3 * n->n_weights * sizeof(float)));
    if (!n->ft_sched) LEMO_CUDA(cudaMalloc(&n->ft_sched, sizeof(Sched)));
    // fresh optim.Adam (opt_amass_perframe.py:127-129): zero moments, step counter 0, constant lr
    LEMO_CUDA(cudaMemsetAsync(n->d_wflat + n->n_weights, 0, 2 * n->n_weights * sizeof(float), st));
    Sched s{};
    s.it = 0; s.lr0 = s.lr1 = s.lr2 = (float)lr; s.sw1 = s.sw2 = 1 << 30;
    s.aux = 1.f / ((float)N * (float)n_rows_selected * (float)n->geom[0].W);     // scale of the masked L1 (read by k_ae_l1_sched)
    LEMO_CUDA(cudaMemcpyAsync(n->ft_sched, &s, sizeof(Sched), cudaMemcpyHostToDevice, st));
    if (losses_out) LEMO_CUDA(cudaMemsetAsync(losses_out, 0, (size_t)steps * sizeof(float), st));
    LEMO_TRY(pack_planes(x, n->act[0], N * n->in_ch, n->geom[0], st));          // the input is the same for every step (:164-184)
    // the captured step bakes these pointers / sizes: re-capture when they change (the number of selected rows is NOT baked: Sched::aux)
    const bool same = n->ft_gexec && n->ft_mask == row_mask && n->ft_losses == losses_out && n->ft_N == N;
    if (!same) {
        if (n->ft_gexec) { cudaGraphExecDestroy((cudaGraphExec_t)n->ft_gexec); n->ft_gexec = nullptr; }
        if (n->ft_graph) { cudaGraphDestroy((cudaGraph_t)n->ft_graph); n->ft_graph = nullptr; }
        n->ft_mask = row_mask; n->ft_losses = losses_out; n->ft_N = N; n->ft_rows = n_rows_selected;
    }
    if (steps == 0) return 0;
    if (!n->ft_stream) {
        cudaStream_t gs; cudaEvent_t e0, e1;
        LEMO_CUDA(cudaStreamCreateWithFlags(&gs, cudaStreamNonBlocking));
        LEMO_CUDA(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
        LEMO_CUDA(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        n->ft_stream = gs; n->ft_ev_in = e0; n->ft_ev_out = e1;
    }
    cudaStream_t gs = (cudaStream_t)n->ft_stream;
    LEMO_CUDA(cudaEventRecord((cudaEvent_t)n->ft_ev_in, st));
    LEMO_CUDA(cudaStreamWaitEvent(gs, (cudaEvent_t)n->ft_ev_in, 0));
    if (!n->ft_gexec) {
        LEMO_CUDA(cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal));
        const int r = ae_finetune_graph_step(n, gs);
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(gs, &g);
        if (r) { if (g) cudaGraphDestroy(g); return r; }
        LEMO_CUDA(e);
        cudaGraphExec_t ge = nullptr;
        LEMO_CUDA(cudaGraphInstantiate(&ge, g, 0));
        n->ft_graph = g; n->ft_gexec = ge;
    }
    for (int i = 0; i < steps; ++i) LEMO_CUDA(cudaGraphLaunch((cudaGraphExec_t)n->ft_gexec, gs));
    LEMO_CUDA(cudaEventRecord((cudaEvent_t)n->ft_ev_out, gs));
    LEMO_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)n->ft_ev_out, 0));
    return 0;
}

int lemo_ae_forward(LemoConvNet* h, const float* x, int32_t N, float* rec, float* z, void* stream) {
    LEMO_NVTX("lemo_ae_forward");
    LEMO_CHECK(h && h->n->kind == 1 && x && rec, "bad arguments (need an AE handle)");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    LEMO_TRY(pack_planes(x, n->act[0], N * n->in_ch, n->geom[0], st));
    LEMO_TRY(ae_forward_planes(n, N, st));
    LEMO_TRY(unpack_planes(n->act[18 + 3 * 4], rec, N * 1, n->geom[0], st));
    if (z) LEMO_TRY(unpack_planes(n->act[3 + 3 * 4], z, N * 256, n->geom[5], st));
    return 0;
}

int lemo_ae_backward_weights(LemoConvNet* h, const float* d_rec, int32_t N, float* d_weights, void* stream) {
    LEMO_CHECK(h && h->n->kind == 1 && d_rec && d_weights && h->n->with_backward, "bad arguments / AE handle created without backward");
    ConvNet* n = h->n;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_CHECK(N > 0 && N <= n->maxN, "batch exceeds handle size");
    LEMO_CUDA(cudaMemsetAsync(n->grad[0], 0, (size_t)N * n->geom[0].PS * sizeof(float), st));
    LEMO_TRY(pack_planes(d_rec, n->grad[0], N, n->geom[0], st));
    return ae_backward_planes(n, N, d_weights, st);
}
}
