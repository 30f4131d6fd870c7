// Persistent per-frame fitting kernel (reference opt_amass_perframe.py:293-361; SURVEY.md section 7 "hard part 1").
//
// The per-frame stage is T sequential B=1 Adam problems x 100 steps per sequence, warm-started frame to frame: as a graph of
// ~23 small kernels per step it is pure launch/dependency latency (127 us per step for 8 chains, every kernel on 1-8 CTAs).
// Here ONE launch runs the whole stage -- all frames, all steps -- with one thread-block CLUSTER of 8 CTAs per sequence:
//   * the VPoser MLP (32 -> 512 -> 512 -> 126, 1.4 MB of fp32 weights) is split by output rows across the 8 CTAs, which stream their
//     slices from L2 every step (22 MB per step for 8 sequences: L2 bandwidth, not HBM) and exchange activations through small
//     per-sequence global (L2-resident) buffers between cluster barriers; the adjoint W^T products are computed as per-CTA partial
//     vectors that every CTA adds in rank order (deterministic);
//   * pose -> rotations -> kinematic chain and its adjoint are the SAME device bodies the stand-alone kernels run (body_dev.cuh),
//     executed redundantly by each CTA (55 joints: latency, not work) on SHARED-MEMORY copies of the model constants (joint
//     regressor directions, hand PCA bases, pose mean, tree) and shared-memory scratch, so no barrier and no L2 round trip is
//     spent on them (with global scratch these three bodies were 18 of the 41 us of a step);
//   * the 81 marker rows of the body model are dealt 11 per CTA: fp32 blend (243 of the 512 x 31425 blend-shape columns), skinning,
//     L1 marker loss and their adjoints; dA / dX / dtransl partials are combined like the MLP partials;
//   * parameters and Adam moments (65 per frame) are replicated in every CTA's shared memory: all CTAs apply the identical update,
//     nothing is broadcast; the learning-rate schedule (.1/.01 -> .01 @>60 -> .003 @>80) and bias corrections are computed in-kernel.
// Four cluster barriers per step, no host involvement until the last frame is done.
#pragma once
#include "body_dev.cuh"
#include <cooperative_groups.h>

namespace lemo {
namespace cgx = cooperative_groups;

constexpr int PM_CL = 8;          // CTAs per cluster = per sequence
constexpr int PM_NT = 256;        // threads per CTA
constexpr int PM_VPC = 11;        // marker rows per CTA (8 x 11 >= 81)
constexpr int PM_PART = 664;      // dA[660] + dtransl[3] + loss partial
constexpr size_t PM_DYN_BYTES = 4 * (size_t)(NJ * 3 * NBETA + 165 + 2 * 540 + 168 + TREE_N + 168 + 496 + 512 + 660 + 660 + 168 + 168 + 660 + 512 + 496 + 12 + 192 + 12 + 192 + 128 + 16 + XK * 33 + NJ * PM_VPC + 16 + 32 * 512 + 512);

struct MegaArgs {
    // loss-row sub-model (V = 81 marker rows)
    const float *vt, *Wt, *wjm, *Jt, *Jd, *hand_l, *hand_r, *pose_mean;
    const int* tree;
    int max_depth, V, npc;
    // VPoser decoder, nn.Linear layout [out][in]
    const float *W1, *b1, *W2, *b2, *W3, *b3;
    float *h2, *o;                               // [S,512] [S,126]
    float *dh1p, *dXp, *dAp;                     // per-CTA partials: [S][8][512], [S][8][512], [S][8][PM_PART]
    // fit state
    float *P, *Gp;
    const float *betas, *mrec;
    float *p72, *acc;
    int acc_n, acc_rec, acc_vp, acc_shape, acc_hand;
    int S, T, n_iters;
    float w_rec, w_vp, w_shape, w_hand;
    unsigned long long* tl;                      // debug timeline (nullable): globaltimer stamps of cluster 0 / rank 0 at the phase boundaries of step 5
};

__device__ __forceinline__ size_t pm_poff(int i, int S, int s) {      // flat parameter vector P = [tr S*3 | r6 S*6 | z S*32 | lh S*12 | rh S*12]
    if (i < 3) return (size_t)s * 3 + i;
    if (i < 9) return (size_t)S * 3 + (size_t)s * 6 + (i - 3);
    if (i < 41) return (size_t)S * 9 + (size_t)s * 32 + (i - 9);
    if (i < 53) return (size_t)S * 41 + (size_t)s * 12 + (i - 41);
    return (size_t)S * 53 + (size_t)s * 12 + (i - 53);
}

__device__ __forceinline__ unsigned long long pm_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PM_STAMP(i) do { if (a.tl && blockIdx.x == 0 && tid == 0 && t == 0 && it == 5) a.tl[i] = pm_now(); } while (0)

__global__ void __cluster_dims__(PM_CL, 1, 1) __launch_bounds__(PM_NT, 1) k_perframe_mega(MegaArgs a) {
    cgx::cluster_group cluster = cgx::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncl = gridDim.x / PM_CL, cid = blockIdx.x / PM_CL;
    const int S = a.S, V = a.V, NC = 3 * a.V;
    __shared__ float s_p[65], s_m[65], s_v[65], s_g[65];
    __shared__ __align__(16) float s_vec[512];          // staging: X / dh2 / dh1
    __shared__ __align__(16) float s_h1[512];
    __shared__ __align__(16) float s_h2[512];
    __shared__ float s_red[8][64];
    __shared__ float s_T[PM_VPC][12], s_dT[PM_VPC][12], s_vp[PM_VPC][3], s_gv[PM_VPC][3], s_dvp[3 * PM_VPC];
    __shared__ float s_tgt[201], s_do[128], s_sc[8], s_vt[3 * PM_VPC], s_dh2[64];

    const int v0 = PM_VPC * rank, nv = max(0, min(PM_VPC, V - v0)), c0 = 3 * v0, ncol = 3 * nv;
    // dynamic shared memory: model constants (loaded once per launch) + the body scratch of "frame 0" of this CTA
    extern __shared__ __align__(16) float dyn[];
    float* c_Jd = dyn;                       // [55*3*20]
    float* c_Jt = c_Jd + NJ * 3 * NBETA;     // [165]
    float* c_hl = c_Jt + NJ * 3;             // [npc*45]
    float* c_hr = c_hl + 12 * 45;
    float* c_pm = c_hr + 12 * 45;            // [165]
    int* c_tree = reinterpret_cast<int*>(c_pm + 168);   // [TREE_N] level-ordered tree tables
    float* b_fp = reinterpret_cast<float*>(c_tree + TREE_N);   // full_pose [165]
    float* b_R = b_fp + 168;                 // [55*9]
    float* b_X = b_R + 496;                  // [512]
    float* b_G = b_X + 512;                  // [660]
    float* s_A = b_G + 660;                  // [660]
    float* b_Jr = s_A + 660;                 // [165]
    float* b_Jp = b_Jr + 168;                // [165]
    float* b_dA = b_Jp + 168;                // [660]
    float* b_dX = b_dA + 660;                // [512]
    float* b_dR = b_dX + 512;                // [495]
    float* b_Rg = b_dR + 496;                // [9]
    float* b_Rb = b_Rg + 12;                 // [189]
    float* b_dRg = b_Rb + 192;               // [9]
    float* b_dRb = b_dRg + 12;               // [189]
    float* b_o = b_dRb + 192;                // [126]
    float* b_beta = b_o + 128;               // [10]
    float* c_Wt = b_beta + 16;               // [512][33]: this CTA's 33 blend-shape columns, resident for the whole launch (67.6 KB)
    float* c_wj = c_Wt + XK * 33;            // [55][11]: skinning weights of this CTA's marker rows
    float* c_W1T = c_wj + NJ * PM_VPC + 16;  // [32][512]: first VPoser layer, transposed (fc1 and its adjoint are done by every CTA: no exchange)
    float* c_b1 = c_W1T + 32 * 512;          // [512]
    for (int i = tid; i < NJ * 3 * NBETA; i += PM_NT) c_Jd[i] = a.Jd[i];
    for (int i = tid; i < NJ * 3; i += PM_NT) { c_Jt[i] = a.Jt[i]; c_pm[i] = a.pose_mean[i]; }
    for (int i = tid; i < a.npc * 45; i += PM_NT) { c_hl[i] = a.hand_l[i]; c_hr[i] = a.hand_r[i]; }
    for (int i = tid; i < TREE_N; i += PM_NT) c_tree[i] = a.tree[i];
    for (int i = tid; i < 512 * 32; i += PM_NT) c_W1T[(i & 31) * 512 + (i >> 5)] = a.W1[i];
    for (int i = tid; i < 512; i += PM_NT) c_b1[i] = a.b1[i];
    if (tid < 3 * PM_VPC) s_vt[tid] = tid < ncol ? a.vt[c0 + tid] : 0.f;
    for (int i = tid; i < XK * 33; i += PM_NT) { const int k = i / 33, cc = i - k * 33; c_Wt[i] = cc < ncol ? a.Wt[(size_t)k * NC + c0 + cc] : 0.f; }
    for (int i = tid; i < NJ * PM_VPC; i += PM_NT) { const int j = i / PM_VPC, v = i - j * PM_VPC; c_wj[i] = v < nv ? a.wjm[(size_t)j * V + v0 + v] : 0.f; }
    __syncthreads();

    PoseK pk;
    pk.in = PoseIn();
    pk.in.transl = s_p; pk.in.R_global = b_Rg; pk.in.R_body = b_Rb;
    pk.in.lhand = s_p + 41; pk.in.rhand = s_p + 53;
    pk.in.betas = b_beta; pk.in.betas_stride = 10; pk.in.hand_is_pca = 1;
    pk.hand_l = c_hl; pk.hand_r = c_hr; pk.pose_mean = c_pm; pk.npc = a.npc;
    PoseGrad pg;
    pg.R_global = b_dRg; pg.R_body = b_dRb; pg.lhand = s_g + 41; pg.rhand = s_g + 53;


    for (int s = cid; s < S; s += ncl) {
        if (tid < 65) s_p[tid] = a.P[pm_poff(tid, S, s)];
        if (tid < 10) b_beta[tid] = a.betas[(size_t)s * 10 + tid];
        __syncthreads();
        for (int t = 0; t < a.T; ++t) {
            if (tid < 65) { s_m[tid] = 0.f; s_v[tid] = 0.f; }                      // fresh optim.Adam per frame (:319)
            if (tid < 201) s_tgt[tid] = a.mrec[((size_t)s * a.T + t) * 201 + tid];
            const float lr0 = t == 0 ? 0.1f : 0.01f;                               // :315-318
            for (int it = 0; it < a.n_iters; ++it) {
                const bool last = it == a.n_iters - 1;
                if (tid == 0) {
                    s_sc[0] = it > 80 ? 0.003f : (it > 60 ? 0.01f : lr0);          // `if step > 60` / `if step > 80` (:323-330)
                    const double tt = (double)(it + 1);
                    s_sc[1] = (float)(1.0 - pow(0.9, tt));
                    s_sc[2] = (float)sqrt(1.0 - pow(0.999, tt));
                }
                PM_STAMP(0);
                // ------------------------------------------------ P1 (every CTA): global 6D -> R ; fc1, all 512 rows from shared memory
                if (tid == 0) gs6d_fwd(s_p + 3, b_Rg);
                {
                    float z[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) z[k] = s_p[9 + k];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int n = tid + rr * PM_NT;
                        float acc = 0.f;
#pragma unroll
                        for (int k = 0; k < 32; ++k) acc = fmaf(c_W1T[k * 512 + n], z[k], acc);
                        s_h1[n] = lrelu(acc + c_b1[n]);
                    }
                }
                PM_STAMP(1);
                __syncthreads();
                PM_STAMP(2);
                // ------------------------------------------------ P2: fc2 rows [64 rank, +64), one warp per row
                // every phase below is a stream of L2 reads with ~1 us of latency each: the loops are unrolled so that 16-32 independent
                // loads are in flight per thread (the first version issued them 1-4 at a time and spent 47 us per step waiting)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 wv[4][4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const float4* w = reinterpret_cast<const float4*>(a.W2 + (size_t)(64 * rank + warp * 8 + half * 4 + rr) * 512);
#pragma unroll
                        for (int i = 0; i < 4; ++i) wv[rr][i] = __ldg(w + i * 32 + lane);
                    }
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int n = 64 * rank + warp * 8 + half * 4 + rr;
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 x = *reinterpret_cast<const float4*>(&s_h1[(i * 32 + lane) * 4]);
                            acc = fmaf(wv[rr][i].x, x.x, acc); acc = fmaf(wv[rr][i].y, x.y, acc);
                            acc = fmaf(wv[rr][i].z, x.z, acc); acc = fmaf(wv[rr][i].w, x.w, acc);
                        }
                        acc = warp_sum(acc);
                        if (lane == 0) a.h2[(size_t)s * 512 + n] = lrelu(acc + __ldg(a.b2 + n));
                    }
                }
                PM_STAMP(3);
                cluster.sync();
                PM_STAMP(4);
                // ------------------------------------------------ P3: output rows [16 rank, +16)
                s_h2[tid] = a.h2[(size_t)s * 512 + tid]; s_h2[tid + 256] = a.h2[(size_t)s * 512 + tid + 256];
                __syncthreads();
                {
                    float4 wv[2][4];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int n = min(125, 16 * rank + warp * 2 + rr);
                        const float4* w = reinterpret_cast<const float4*>(a.W3 + (size_t)n * 512);
#pragma unroll
                        for (int i = 0; i < 4; ++i) wv[rr][i] = __ldg(w + i * 32 + lane);
                    }
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int n = 16 * rank + warp * 2 + rr;
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 x = *reinterpret_cast<const float4*>(&s_h2[(i * 32 + lane) * 4]);
                            acc = fmaf(wv[rr][i].x, x.x, acc); acc = fmaf(wv[rr][i].y, x.y, acc);
                            acc = fmaf(wv[rr][i].z, x.z, acc); acc = fmaf(wv[rr][i].w, x.w, acc);
                        }
                        acc = warp_sum(acc);
                        if (lane == 0 && n < 126) a.o[(size_t)s * 126 + n] = acc + __ldg(a.b3 + n);
                    }
                }
                PM_STAMP(5);
                cluster.sync();
                PM_STAMP(6);
                // ------------------------------------------------ P4 (every CTA): Gram-Schmidt, pose -> R -> chain (shared device body)
                if (tid < 126) b_o[tid] = a.o[(size_t)s * 126 + tid];
                __syncthreads();
                if (tid < NBODY) gs6d_fwd(b_o + tid * 6, b_Rb + tid * 9);
                __syncthreads();
                pose_chain_fwd_body<false, true>(pk, c_Jt, c_Jd, c_tree, a.max_depth, b_fp, b_R, b_X, nullptr, b_G, s_A, b_Jr, b_Jp, nullptr, 0, last);
                __syncthreads();
                if (last && rank == 0 && tid < 72) {                               // the [T,72] row the script saves: parameters of the LAST forward
                    float v;
                    if (tid < 3) v = s_p[tid];
                    else if (tid < 6) v = b_fp[tid - 3];
                    else if (tid < 16) v = b_beta[tid - 6];
                    else v = s_p[9 + (tid - 16)];
                    a.p72[((size_t)s * a.T + t) * 72 + tid] = v;
                }
                PM_STAMP(7);
                // ------------------------------------------------ P5: blend + skinning + L1 loss on this CTA's marker rows
                if (tid < 7 * 33) {                                                 // 7 k-groups x 33 columns: consecutive threads, consecutive words
                    const int grp = tid / 33, cc = tid - grp * 33;
                    float acc = 0.f;
#pragma unroll 8
                    for (int k = grp; k < XK; k += 7) acc = fmaf(b_X[k], c_Wt[k * 33 + cc], acc);
                    s_red[grp][cc] = acc;
                }
                if (tid < nv * 12) {
                    const int i = tid / 12, k = tid - i * 12;
                    float acc = 0.f;
#pragma unroll 11
                    for (int j = 0; j < NJ; ++j) acc = fmaf(c_wj[j * PM_VPC + i], s_A[j * 12 + k], acc);
                    s_T[i][k] = acc;
                }
                __syncthreads();
                if (tid < ncol) s_vp[tid / 3][tid % 3] = s_vt[tid] + (((s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid])) + ((s_red[4][tid] + s_red[5][tid]) + s_red[6][tid]));
                __syncthreads();
                float lpart = 0.f;
                if (tid < ncol) {
                    const int i = tid / 3, r = tid - i * 3;
                    const float v = s_T[i][r * 4] * s_vp[i][0] + s_T[i][r * 4 + 1] * s_vp[i][1] + s_T[i][r * 4 + 2] * s_vp[i][2] + s_T[i][r * 4 + 3] + s_p[r];
                    float g = 0.f;
                    if (v0 + i < 67) {                                              // F.l1_loss on the 67 SSM2 markers (:337-340)
                        const float d = v - s_tgt[(v0 + i) * 3 + r];
                        lpart = fabsf(d) * (1.f / 201.f);
                        g = a.w_rec * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * (1.f / 201.f);
                    }
                    s_gv[i][r] = g;
                }
                if (warp < 2) {                                                     // ncol <= 33: warps 0 and 1 hold the loss parts
                    lpart = warp_sum(lpart);
                    if (lane == 0) s_sc[4 + warp] = lpart;
                }
                __syncthreads();
                PM_STAMP(8);
                // ------------------------------------------------ P6: adjoint on this CTA's rows -> partials of dA, dtransl, dX
                if (tid < 3 * PM_VPC) {                                             // (rows past nv: zeros, so the loops below need no predicate)
                    const int i = tid / 3, c = tid - i * 3;
                    s_dvp[tid] = tid < ncol ? s_T[i][c] * s_gv[i][0] + s_T[i][4 + c] * s_gv[i][1] + s_T[i][8 + c] * s_gv[i][2] : 0.f;
                }
                if (tid < PM_VPC * 12) {
                    const int i = tid / 12, k = tid - i * 12, row = k >> 2, col = k & 3;
                    s_dT[i][k] = i < nv ? s_gv[i][row] * (col < 3 ? s_vp[i][col] : 1.f) : 0.f;
                }
                __syncthreads();
                {
                    float* pz = a.dAp + ((size_t)s * PM_CL + rank) * PM_PART;
                    for (int e = tid; e < NJ * 12; e += PM_NT) {
                        const int j = e / 12, k = e - j * 12;
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < PM_VPC; ++i) acc = fmaf(c_wj[j * PM_VPC + i], s_dT[i][k], acc);
                        pz[e] = acc;
                    }
                    if (tid < 3) {
                        float acc = 0.f;
                        for (int i = 0; i < nv; ++i) acc += s_gv[i][tid];
                        pz[NJ * 12 + tid] = acc;
                    }
                    if (tid == 3) pz[NJ * 12 + 3] = s_sc[4] + s_sc[5];
                    float* px = a.dXp + ((size_t)s * PM_CL + rank) * XK;
                    float dv[3 * PM_VPC];
#pragma unroll
                    for (int cc = 0; cc < 3 * PM_VPC; ++cc) dv[cc] = s_dvp[cc];
#pragma unroll
                    for (int kk = 0; kk < XK / PM_NT; ++kk) {
                        const int k = tid + kk * PM_NT;
                        const float* w = c_Wt + k * 33;
                        float acc = 0.f;
#pragma unroll
                        for (int cc = 0; cc < 3 * PM_VPC; ++cc) acc = fmaf(w[cc], dv[cc], acc);
                        px[k] = acc;
                    }
                }
                PM_STAMP(9);
                cluster.sync();
                PM_STAMP(10);
                // ------------------------------------------------ P7 (every CTA): combine partials in rank order, chain adjoint
                // (all partial loads of a thread are independent: issued together, one L2 latency for the lot)
#pragma unroll
                for (int ee = 0; ee < (PM_PART + PM_NT - 1) / PM_NT; ++ee) {
                    const int e = tid + ee * PM_NT;
                    if (e >= PM_PART) break;
                    float pv[PM_CL];
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) pv[r] = __ldcg(a.dAp + ((size_t)s * PM_CL + r) * PM_PART + e);
                    float acc = 0.f;
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) acc += pv[r];
                    if (e < NJ * 12) b_dA[e] = acc;                                  // joint-major [55][B*12] with B = 1
                    else if (e < NJ * 12 + 3) s_g[e - NJ * 12] = acc;               // d loss / d transl
                    else s_sc[3] = acc;                                             // marker loss value
                }
#pragma unroll
                for (int kk = 0; kk < XK / PM_NT; ++kk) {
                    const int k = tid + kk * PM_NT;
                    float pv[PM_CL];
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) pv[r] = __ldcg(a.dXp + ((size_t)s * PM_CL + r) * XK + k);
                    float acc = 0.f;
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) acc += pv[r];
                    b_dX[k] = acc;
                }
                __syncthreads();
                PM_STAMP(11);
                chain_bwd_body<false, false, true>(b_R, b_G, b_Jr, b_dA, nullptr, b_dX, c_Jd, c_tree, a.max_depth, 1, 10, b_dR, nullptr, nullptr, 0);
                __syncthreads();
                PM_STAMP(12);
                // Gram-Schmidt adjoints on warps 2-3, straight from dR, while warps 0-1 run the axis-angle adjoint of the hands
                if (tid == 64) gs6d_bwd(s_p + 3, b_dR, s_g + 3);
                if (tid >= 96 && tid < 96 + NBODY) gs6d_bwd(b_o + (tid - 96) * 6, b_dR + (1 + tid - 96) * 9, s_do + (tid - 96) * 6);
                pose_to_rot_bwd_body(pk, pg, 1, b_fp, b_dR, 0);
                __syncthreads();                                                    // (the hand PCA gradients were written straight into s_g)
                PM_STAMP(13);
                // ------------------------------------------------ P8: dh2 columns [64 rank, +64) = (W3^T d_o) * lrelu'(h2)
                {
                    const int nn = tid & 63, og = tid >> 6, n = 64 * rank + nn;
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int o = og + 4 * i;
                        if (o < 126) acc = fmaf(__ldg(a.W3 + (size_t)o * 512 + n), s_do[o], acc);
                    }
                    s_red[og][nn] = acc;
                }
                __syncthreads();
                if (tid < 64) {
                    const int n = 64 * rank + tid;
                    const float v = (s_red[0][tid] + s_red[1][tid]) + (s_red[2][tid] + s_red[3][tid]);
                    s_dh2[tid] = v * (s_h2[n] > 0.f ? 1.f : 0.2f);
                }
                PM_STAMP(14);
                __syncthreads();                                                    // (no cluster barrier: P9 needs only THIS CTA's 64 columns of dh2)
                PM_STAMP(15);
                // ------------------------------------------------ P9: partial of dh1 = W2^T dh2 over this CTA's rows of W2
                {
                    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 16
                    for (int r = 0; r < 64; ++r) {
                        const int n = 64 * rank + r;
                        const float d = s_dh2[r];
                        acc0 = fmaf(__ldg(a.W2 + (size_t)n * 512 + tid), d, acc0);
                        acc1 = fmaf(__ldg(a.W2 + (size_t)n * 512 + tid + 256), d, acc1);
                    }
                    float* ph = a.dh1p + ((size_t)s * PM_CL + rank) * 512;
                    ph[tid] = acc0; ph[tid + 256] = acc1;
                }
                PM_STAMP(16);
                cluster.sync();
                PM_STAMP(17);
                // ------------------------------------------------ P10 (every CTA): dh1, dz = W1^T dh1, priors, Adam
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < 512 / PM_NT; ++kk) {
                    const int c = tid + kk * PM_NT;
                    float pv[PM_CL];
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) pv[r] = __ldcg(a.dh1p + ((size_t)s * PM_CL + r) * 512 + c);
                    float acc = 0.f;
#pragma unroll
                    for (int r = 0; r < PM_CL; ++r) acc += pv[r];
                    s_vec[c] = acc * (s_h1[c] > 0.f ? 1.f : 0.2f);
                }
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 4; ++q) {                                       // dz[c] = W1^T dh1: one warp per latent, lanes over the 512 rows
                    const int c = warp + 8 * q;
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc = fmaf(c_W1T[c * 512 + lane + 32 * i], s_vec[lane + 32 * i], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) s_g[9 + c] = acc + a.w_vp * 2.f * s_p[9 + c] * (1.f / 32.f);       // + d/dz of w_vposer * mean(z^2)  (:343-346)
                }
                if (tid >= 32 && tid < 32 + 24) {
                    const int c = tid - 32;
                    s_g[41 + c] += a.w_hand * 2.f * s_p[41 + c] * (1.f / 24.f);           // + d of w_hand * mean(hand^2)
                }
                __syncthreads();
                if (last && rank == 0 && tid == 0 && a.acc) {                       // loss terms of the last closure of this frame
                    float pv = 0.f, ph = 0.f, ps = 0.f;
                    for (int k = 0; k < 32; ++k) pv += s_p[9 + k] * s_p[9 + k];
                    for (int k = 0; k < 24; ++k) ph += s_p[41 + k] * s_p[41 + k];
                    for (int k = 0; k < 10; ++k) ps += b_beta[k] * b_beta[k];
                    float* ac = a.acc + (size_t)s * a.acc_n;
                    ac[a.acc_rec] = s_sc[3]; ac[a.acc_vp] = pv / 32.f; ac[a.acc_hand] = ph / 24.f; ac[a.acc_shape] = ps / 10.f;
                }
                if (last) __syncthreads();                                          // the reported terms read s_p before Adam moves it (racecheck)
                PM_STAMP(18);
                if (tid < 65) {                                                     // torch.optim.Adam.step, identical in every CTA
                    const float gi = s_g[tid];
                    const float mi = 0.9f * s_m[tid] + (1.f - 0.9f) * gi;
                    const float vi = 0.999f * s_v[tid] + (1.f - 0.999f) * gi * gi;
                    s_m[tid] = mi; s_v[tid] = vi;
                    s_p[tid] -= (s_sc[0] / s_sc[1]) * (mi / (sqrtf(vi) / s_sc[2] + 1e-8f));
                }
                __syncthreads();
            }
        }
        if (tid < 65) a.P[pm_poff(tid, S, s)] = s_p[tid];                           // state after the last step (lemo_fit_get_state)
        if (tid < 65) a.Gp[pm_poff(tid, S, s)] = s_g[tid];
        __syncthreads();
    }
}

}  // namespace lemo
