// Pieces shared by the fused fitting drivers (fit.cu: AMASS temporal / per-frame stages; fit_prox.cu: PROX stage-2 window):
// device-resident schedule + Adam, and the smoothness-prior input / adjoint kernels (opt_amass_temp.py:377-391 ==
// fitting_temp_slide.py:1012-1031).  Kernels are `static` so that both translation units can include this header.
#pragma once
#include "common.cuh"

namespace lemo {

struct Sched {            // device-resident schedule: constants set per run, scalars updated per iteration
    int it;               // iterations done so far in the current run
    float lr, bc1, bc2s;  // this iteration's learning rate and Adam bias corrections
    float lr0, lr1, lr2;  // lr = it > sw2 ? lr2 : it > sw1 ? lr1 : lr0   (`if step > 60` semantics of the scripts)
    int sw1, sw2;
    int frame;            // per-frame mode: which frame of every sequence is being fitted
    float aux;            // per-run scalar of the driver (AE fine-tune: 1 / number of selected elements), so that it is not baked into the graph
};

static __global__ void k_sched(Sched* s) {
    const int it = s->it;                         // 0-based step index of this iteration
    s->lr = it > s->sw2 ? s->lr2 : (it > s->sw1 ? s->lr1 : s->lr0);
    const double t = (double)(it + 1);
    s->bc1 = (float)(1.0 - pow(0.9, t));
    s->bc2s = (float)sqrt(1.0 - pow(0.999, t));
    s->it = it + 1;
}

static __global__ void k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int n,
                           const Sched* __restrict__ s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = 0.9f * m[i] + (1.f - 0.9f) * gi;
    const float vi = 0.999f * v[i] + (1.f - 0.999f) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= (s->lr / s->bc1) * (mi / (sqrtf(vi) / s->bc2s + 1e-8f));
}

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

// Enc input (opt_amass_temp.py:377-387): canonicalise, normalise, temporal difference, reflect pad (8,8,1,1)
static __global__ void k_smooth_input(const float* __restrict__ Vr, const float* __restrict__ canon, const float* __restrict__ stats, int T, int NR,
                               int H, int W, int Wp, int PS, float* __restrict__ xin) {
    const int s = blockIdx.z;
    const int tt = blockIdx.x * blockDim.x + threadIdx.x, dd = blockIdx.y;
    if (tt >= W) return;
    const int d = reflect_idx(dd - 1, H - 2), t = reflect_idx(tt - 8, W - 16);
    const int mk = d / 3, c = d - mk * 3;
    const float* cn = canon + s * 12;
    const float mu = stats[d], sd = stats[243 + d];
    float val[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float* m = Vr + (((size_t)s * T + t + e) * NR + mk) * 3;
        const float a = (m[0] - cn[9]) * cn[c] + (m[1] - cn[10]) * cn[3 + c] + (m[2] - cn[11]) * cn[6 + c];
        val[e] = (a - mu) / sd;
    }
    xin[(size_t)s * PS + (dd + 1) * Wp + tt + 1] = val[1] - val[0];
}

// loss_smooth = mean((z[...,1:]-z[...,:-1])^2) (opt_amass_temp.py:390-391) and dL/dpre of the last Enc layer
static __global__ void __launch_bounds__(256) k_smooth_loss(const float* __restrict__ z, int C, int H, int W, int Wp, int PS, float w,
                                                     float* __restrict__ gpre, float* __restrict__ acc, int acc_stride, int acc_slot) {
    __shared__ float sred[32];
    const int s = blockIdx.z, c = blockIdx.y;
    const float inv_n = 1.f / ((float)C * (float)H * (float)(W - 1));
    const float* zp = z + ((size_t)s * C + c) * PS;
    float* gp = gpre + ((size_t)s * C + c) * PS;
    float part = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const int q = (y + 1) * Wp + x + 1;
        const float zc = zp[q];
        float g = 0.f;
        if (x >= 1) g += zc - zp[q - 1];
        if (x <= W - 2) { const float d = zp[q + 1] - zc; g -= d; part += d * d; }
        gp[q] = w * 2.f * inv_n * g * (zc > 0.f ? 1.f : 0.2f);
    }
    part = block_sum(part, sred);
    if (threadIdx.x == 0) atomicAdd(&acc[s * acc_stride + acc_slot], part * inv_n);
}

// adjoint of reflect pad: gv[s][d][t] = sum of gx over the padded positions that read (d,t)
static __global__ void k_smooth_bwd_a(const float* __restrict__ gx, int T, int H, int W, int Wp, int PS, float* __restrict__ gv) {
    const int s = blockIdx.z, d = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = T - 1;                        // == W - 16
    if (t >= n) return;
    int dds[2] = {d + 1, -1};
    if (d == 1) dds[1] = 0;
    if (d == H - 4) dds[1] = H - 1;             // reflect(H-2) = 2(H-3)-(H-2) = H-4   (H-2 = 243 rows before padding)
    int tts[3] = {t + 8, -1, -1};               // own slot, left reflection, right reflection (they overlap when T-1 < 18)
    if (t >= 1 && t <= 8) tts[1] = 8 - t;
    if (t >= n - 9 && t <= n - 2) tts[2] = 8 + 2 * (n - 1) - t;
    float a = 0.f;
    for (int i = 0; i < 2; ++i) {
        if (dds[i] < 0) continue;
        for (int j = 0; j < 3; ++j) {
            if (tts[j] < 0) continue;
            a += gx[(size_t)s * PS + (dds[i] + 1) * Wp + tts[j] + 1];
        }
    }
    gv[((size_t)s * (H - 2) + d) * n + t] = a;
}
// gv -> gradient on the 81 marker rows
static __global__ void k_smooth_bwd_b(const float* __restrict__ gv, const float* __restrict__ canon, const float* __restrict__ stats, int T, int NR,
                               int S, float* __restrict__ Grows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * T * 81) return;
    const int mk = i % 81, t = (i / 81) % T, s = i / (81 * T);
    const int n = T - 1;
    const float* cn = canon + s * 12;
    float gval[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = mk * 3 + c;
        const float* g = gv + ((size_t)s * 243 + d) * n;
        float a = 0.f;
        if (t >= 1) a += g[t - 1];
        if (t <= n - 1) a -= g[t];
        gval[c] = a / stats[243 + d];
    }
    float* o = Grows + (((size_t)s * T + t) * NR + mk) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] += gval[0] * cn[k * 3] + gval[1] * cn[k * 3 + 1] + gval[2] * cn[k * 3 + 2];
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static int dalloc(T** p, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
    LEMO_CUDA(cudaMemset(*p, 0, n * sizeof(T)));
    return 0;
}

}  // namespace lemo
