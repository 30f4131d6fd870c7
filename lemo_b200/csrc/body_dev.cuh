// Device bodies of the per-frame pose / kinematic-chain kernels: one CTA (any size >= 64 threads, thread j = joint j) per frame b.
// Shared by the stand-alone kernels of body.cu and by the persistent per-frame fitting kernel (perframe_mega.cuh), which calls them
// from inside its iteration loop so that both paths run the SAME arithmetic.  Every thread of the CTA must call them (barriers inside).
#pragma once
#include "body.cuh"

namespace lemo {

// -DLEMO_BODY_TL: globaltimer stamps inside the bodies (CTA 0, thread 0), read back by the per-frame driver's timeline print (debug builds only)
#ifdef LEMO_BODY_TL
static __device__ unsigned long long g_body_tl[32];
#define BODY_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_body_tl[i] = t_; } } while (0)
#else
#define BODY_STAMP(i) do { } while (0)
#endif

// Barrier of the pose / chain bodies.  SUB64 = false: the whole CTA (stand-alone kernels, 64 threads).  SUB64 = true: named barrier 1
// over the first 64 threads only -- the persistent per-frame kernel calls the bodies from its warps 0-1 while warps 2-7 wait at the
// next CTA-wide barrier, so the ~20 level barriers of a chain walk are two-warp barriers instead of eight-warp ones.
template <bool SUB64>
__device__ __forceinline__ void body_sync() {
    if (SUB64) asm volatile("bar.sync 1, 64;" ::: "memory");
    else __syncthreads();
}

// =============================================================================================
// pose -> rotation matrices        (one thread per (frame, joint))
// =============================================================================================
struct PoseK {
    PoseIn in;
    const float* hand_l; const float* hand_r; const float* pose_mean;
    int npc;
};

__device__ __forceinline__ void joint_aa(const PoseK& p, int b, int j, float* aa) {
    aa[0] = aa[1] = aa[2] = 0.f;
    if (j == 0) { if (p.in.global_orient) for (int k = 0; k < 3; ++k) aa[k] = p.in.global_orient[b * 3 + k]; }
    else if (j <= NBODY) { if (p.in.body_pose) for (int k = 0; k < 3; ++k) aa[k] = p.in.body_pose[b * 63 + (j - 1) * 3 + k]; }
    else if (j == 22) { if (p.in.jaw) for (int k = 0; k < 3; ++k) aa[k] = p.in.jaw[b * 3 + k]; }
    else if (j == 23) { if (p.in.leye) for (int k = 0; k < 3; ++k) aa[k] = p.in.leye[b * 3 + k]; }
    else if (j == 24) { if (p.in.reye) for (int k = 0; k < 3; ++k) aa[k] = p.in.reye[b * 3 + k]; }
    else {
        const bool left = j < 40;
        const int h = left ? j - 25 : j - 40;
        const float* src = left ? p.in.lhand : p.in.rhand;
        if (src) {
            if (p.in.hand_is_pca) {
                const float* comp = left ? p.hand_l : p.hand_r;
                for (int c = 0; c < p.npc; ++c) {
                    const float a = src[b * p.npc + c];
                    for (int k = 0; k < 3; ++k) aa[k] = fmaf(a, comp[c * 45 + h * 3 + k], aa[k]);
                }
            } else {
                for (int k = 0; k < 3; ++k) aa[k] = src[b * 45 + h * 3 + k];
            }
        }
    }
    for (int k = 0; k < 3; ++k) aa[k] += p.pose_mean[j * 3 + k];
}

// adjoint: dR -> parameter gradients.  One block (64 threads) per frame.
template <bool SUB64 = false>
__device__ __forceinline__ void pose_to_rot_bwd_body(const PoseK& p, const PoseGrad& g, int B, const float* __restrict__ full_pose,
                                                     const float* __restrict__ dR, int b) {
    __shared__ float s_daa[NJ * 3];
    const int j = threadIdx.x;
    BODY_STAMP(14);
    if (j < NJ) {
        const float* dr = dR + (b * NJ + j) * 9;
        const bool ov = (j == 0 && p.in.R_global) || (j >= 1 && j <= NBODY && p.in.R_body);
        float daa[3] = {0.f, 0.f, 0.f};
        if (ov) {
            float* o = (j == 0) ? (g.R_global ? g.R_global + b * 9 : nullptr)
                                : (g.R_body ? g.R_body + (b * NBODY + (j - 1)) * 9 : nullptr);
            if (o) for (int k = 0; k < 9; ++k) o[k] = dr[k];
        } else {
            float aa[3] = {full_pose[b * 165 + j * 3], full_pose[b * 165 + j * 3 + 1], full_pose[b * 165 + j * 3 + 2]};
            rodrigues_bwd(aa, dr, daa);
        }
        for (int k = 0; k < 3; ++k) s_daa[j * 3 + k] = daa[k];
        if (!ov) {
            float* o = nullptr;
            if (j == 0) o = g.global_orient ? g.global_orient + b * 3 : nullptr;
            else if (j <= NBODY) o = g.body_pose ? g.body_pose + b * 63 + (j - 1) * 3 : nullptr;
            else if (j == 22) o = g.jaw ? g.jaw + b * 3 : nullptr;
            else if (j == 23) o = g.leye ? g.leye + b * 3 : nullptr;
            else if (j == 24) o = g.reye ? g.reye + b * 3 : nullptr;
            else if (!p.in.hand_is_pca) {
                const bool left = j < 40;
                float* base = left ? g.lhand : g.rhand;
                o = base ? base + b * 45 + (left ? j - 25 : j - 40) * 3 : nullptr;
            }
            if (o) for (int k = 0; k < 3; ++k) o[k] = daa[k];
        }
    }
    BODY_STAMP(15);
    body_sync<SUB64>();
    BODY_STAMP(16);
    if (p.in.hand_is_pca && j < 2 * p.npc) {
        const bool left = j < p.npc;
        const int c = left ? j : j - p.npc;
        float* base = left ? g.lhand : g.rhand;
        if (base) {
            const float* comp = (left ? p.hand_l : p.hand_r) + c * 45;
            const float* d = s_daa + (left ? 25 : 40) * 3;
            float acc = 0.f;
            for (int q = 0; q < 45; ++q) acc = fmaf(comp[q], d[q], acc);
            base[b * p.npc + c] = acc;
        }
    }
    BODY_STAMP(17);
}

// =============================================================================================
// pose -> rotations -> kinematic chain, one kernel: one block (64 threads) per frame, thread j = joint j.
//   * axis-angle -> R by Rodrigues (lbs.py:166-193), or rotation-matrix overrides (global 6D rotation / VPoser output) whose
//     axis-angle is produced with the torchgeometry algorithm for the [T,72] result
//   * blend-shape operand X = [R(1..54) - I | betas | expression | 0] and its TF32 split X2 = [Xhi | Xlo] (Xhi = X rounded to
//     TF32's 10 explicit mantissa bits, Xlo = X - Xhi), staged in shared memory and written as coalesced rows
//   * rest joints from the folded regressor, then the chain (lbs.py:196-263) level-synchronously over the tree (<= 10 levels)
// (Two kernels -- one thread per (frame, joint), then this block shape -- cost 5.5 + 8 us per forward at B=120; forking the chain
// onto a side stream beside the blend GEMM was measured too: the fork/join edges cost what the overlap saves.)
// =============================================================================================
// The tree tables are read inside the single-warp walks, where every load is on the critical path: callers that do not already keep them
// in shared memory (the stand-alone kernels: TREE_SMEM = false) stage them here, before the body's first barrier.
template <bool TREE_SMEM>
__device__ __forceinline__ const int* stage_tree(const int* __restrict__ tree, int* s_tree, int nthreads) {
    if (TREE_SMEM) return tree;
    for (int i = threadIdx.x; i < TREE_N; i += nthreads) s_tree[i] = __ldg(tree + i);
    return s_tree;
}

// want_aa = false skips the axis-angle of rotation-matrix overrides (only the saved [T,72] rows need it).
template <bool SUB64 = false, bool TREE_SMEM = false>
__device__ __forceinline__ void pose_chain_fwd_body(const PoseK& p, const float* __restrict__ J_template,
                                                    const float* __restrict__ J_dirs, const int* __restrict__ tree, int max_depth,
                                                    float* __restrict__ full_pose,
                                                    float* __restrict__ R, float* __restrict__ X, float* __restrict__ X2,
                                                    float* __restrict__ G,
                                                    float* __restrict__ A, float* __restrict__ Jrest, float* __restrict__ Jposed,
                                                    float* __restrict__ A2, int b, bool want_aa = true) {
    __shared__ __align__(16) float sG[NJ][12];
    __shared__ float sJ[NJ][3];
    __shared__ float sR[NJ][9];
    __shared__ __align__(16) float sbeta[NBETA];   // 16-byte aligned: the compiler reads it with LDS.128, which otherwise straddles a neighbour array (racecheck)
    __shared__ __align__(16) float sX[XK];
    const int j = threadIdx.x;
    const int nthr = SUB64 ? 64 : blockDim.x;
    const int* T = tree;
    int par = -1, dep = 0;
    if (!TREE_SMEM && j < NJ) { par = __ldg(tree + TREE_PAR + j); dep = __ldg(tree + TREE_DEPTH + j); }
    float rj[9];
    BODY_STAMP(0);
    if (j < NBETA) {
        float v = 0.f;
        if (j < 10) v = p.in.betas ? p.in.betas[(size_t)b * p.in.betas_stride + j] : 0.f;
        else v = p.in.expression ? p.in.expression[b * 10 + (j - 10)] : 0.f;
        sbeta[j] = v;
        sX[NPF + j] = v;
    }
    if (j >= NBETA && j < XK - NPF) sX[NPF + j] = 0.f;
    if (j < NJ) {
        float aa[3], r[9];
        const float* Rov = nullptr;
        if (j == 0 && p.in.R_global) Rov = p.in.R_global + b * 9;
        if (j >= 1 && j <= NBODY && p.in.R_body) Rov = p.in.R_body + (b * NBODY + (j - 1)) * 9;
        if (Rov) {
            for (int k = 0; k < 9; ++k) r[k] = Rov[k];
            if (want_aa) rotmat_to_aa_tgm(r, aa);    // what the reference scripts store in the [T,72] result
            else aa[0] = aa[1] = aa[2] = 0.f;
        } else {
            joint_aa(p, b, j, aa);
            rodrigues_fwd(aa, r);
        }
        for (int k = 0; k < 3; ++k) full_pose[b * 165 + j * 3 + k] = aa[k];
        for (int k = 0; k < 9; ++k) { R[((size_t)b * NJ + j) * 9 + k] = r[k]; sR[j][k] = r[k]; rj[k] = r[k]; }
        if (j >= 1)
            for (int k = 0; k < 9; ++k) sX[(j - 1) * 9 + k] = r[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
    }
    BODY_STAMP(1);
    body_sync<SUB64>();
    BODY_STAMP(2);
    for (int k = threadIdx.x; k < XK; k += nthr) {
        const float x = sX[k];
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(x));       // round to nearest: k_blend_v2 multiplies most columns by Xhi alone
        const float hi = __uint_as_float(t);
        X[(size_t)b * XK + k] = x;
        if (X2) {
            X2[(size_t)b * 2 * XK + k] = hi;
            X2[(size_t)b * 2 * XK + XK + k] = x - hi;
        }
    }
    if (j < NJ) {
        for (int k = 0; k < 3; ++k) {
            float acc = J_template[j * 3 + k];
            const float* jd = J_dirs + (j * 3 + k) * NBETA;
            for (int l = 0; l < NBETA; ++l) acc = fmaf(jd[l], sbeta[l], acc);
            sJ[j][k] = acc;
            Jrest[((size_t)b * NJ + j) * 3 + k] = acc;
        }
    }
    BODY_STAMP(3);
    body_sync<SUB64>();
    BODY_STAMP(4);
    // the chain (lbs.py:196-263): warp 0 walks the tree level by level, lane i = i-th joint of the level (one packed table word per level
    // and lane, the next level's word requested a level ahead); only __syncwarp between levels.  (Fetching the next level's rotation and
    // offset ahead of the barrier as well was measured: slower -- the loads must complete before the barrier, so a level then pays two
    // shared-memory latencies instead of one.)
    if (!TREE_SMEM) {
        // stand-alone kernel (64 threads, tables in global memory): thread j = joint j, parent / depth in registers, a two-warp CTA
        // barrier per level -- cheaper here than staging the level tables for a single-warp walk (ncu at B = 120: 10.9 vs 13.0 us)
        for (int lev = 0; lev <= max_depth; ++lev) {
            if (j < NJ && dep == lev) {
                float g[12];
                if (lev == 0) {
                    for (int q = 0; q < 3; ++q) { g[q * 4] = rj[q * 3]; g[q * 4 + 1] = rj[q * 3 + 1]; g[q * 4 + 2] = rj[q * 3 + 2]; g[q * 4 + 3] = sJ[j][q]; }
                } else {
                    const float* gp = sG[par];
                    const float t[3] = {sJ[j][0] - sJ[par][0], sJ[j][1] - sJ[par][1], sJ[j][2] - sJ[par][2]};
                    for (int q = 0; q < 3; ++q) {
                        for (int c = 0; c < 3; ++c)
                            g[q * 4 + c] = gp[q * 4] * rj[c] + gp[q * 4 + 1] * rj[3 + c] + gp[q * 4 + 2] * rj[6 + c];
                        g[q * 4 + 3] = gp[q * 4] * t[0] + gp[q * 4 + 1] * t[1] + gp[q * 4 + 2] * t[2] + gp[q * 4 + 3];
                    }
                }
                for (int k = 0; k < 12; ++k) sG[j][k] = g[k];
            }
            body_sync<SUB64>();
        }
    } else if (threadIdx.x < 32) {
        int e = T[TREE_LANE + threadIdx.x];
        for (int lev = 0; lev <= max_depth; ++lev) {
            const int e_next = lev < max_depth ? T[TREE_LANE + (lev + 1) * 32 + threadIdx.x] : -1;
            if (e >= 0) {
                const int jj = e & 255, par = ((e >> 8) & 255) - 1;
                const float* r = sR[jj];
                float g[12];
                if (par < 0) {
                    for (int q = 0; q < 3; ++q) { g[q * 4] = r[q * 3]; g[q * 4 + 1] = r[q * 3 + 1]; g[q * 4 + 2] = r[q * 3 + 2]; g[q * 4 + 3] = sJ[jj][q]; }
                } else {
                    const float* gp = sG[par];
                    const float t[3] = {sJ[jj][0] - sJ[par][0], sJ[jj][1] - sJ[par][1], sJ[jj][2] - sJ[par][2]};
                    for (int q = 0; q < 3; ++q) {
                        for (int c = 0; c < 3; ++c)
                            g[q * 4 + c] = gp[q * 4] * r[c] + gp[q * 4 + 1] * r[3 + c] + gp[q * 4 + 2] * r[6 + c];
                        g[q * 4 + 3] = gp[q * 4] * t[0] + gp[q * 4 + 1] * t[1] + gp[q * 4 + 2] * t[2] + gp[q * 4 + 3];
                    }
                }
                for (int k = 0; k < 12; ++k) sG[jj][k] = g[k];
            }
            __syncwarp();
            e = e_next;
        }
    }
    BODY_STAMP(5);
    body_sync<SUB64>();
    BODY_STAMP(6);
    if (j < NJ) {
        const float* g = sG[j];
        float* go = G + ((size_t)b * NJ + j) * 12;
        float* ao = A + ((size_t)b * NJ + j) * 12;
        for (int k = 0; k < 12; ++k) go[k] = g[k];
        for (int i = 0; i < 3; ++i) {
            ao[i * 4] = g[i * 4]; ao[i * 4 + 1] = g[i * 4 + 1]; ao[i * 4 + 2] = g[i * 4 + 2];
            ao[i * 4 + 3] = g[i * 4 + 3] - (g[i * 4] * sJ[j][0] + g[i * 4 + 1] * sJ[j][1] + g[i * 4 + 2] * sJ[j][2]);
            Jposed[((size_t)b * NJ + j) * 3 + i] = g[i * 4 + 3];
        }
        if (A2) {                 // transposed TF32 split for the tensor-core skinning GEMM: row (b, k), column j
            float* chunk = A2 + (size_t)(b / SKIN_TC_FR) * (4 * SKIN_TC_FR * 12 * 32);      // four [96][32] sub-tiles per 8-frame chunk
            const int r0 = (b % SKIN_TC_FR) * 12, sub = j >> 5, col = j & 31;
            for (int k = 0; k < 12; ++k) {
                const float a = ao[k];
                const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
                chunk[((size_t)sub * (SKIN_TC_FR * 12) + r0 + k) * 32 + col] = hi;
                chunk[((size_t)(2 + sub) * (SKIN_TC_FR * 12) + r0 + k) * 32 + col] = a - hi;
            }
        }
    }
    BODY_STAMP(7);
}

// adjoint of k_chain_fwd.  dA is joint-major [55][B*12]; dJp [B,55,3]; dX [B,512].
// Writes dR [B,55,9]; betas/expression grads (per frame, or atomically into one row when betas_stride==0).
// NEED_J = false drops the rest-joint adjoint (it only feeds the betas / expression gradients).
template <bool SUB64 = false, bool NEED_J = true, bool TREE_SMEM = false>
__device__ __forceinline__ void chain_bwd_body(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ Jrest,
                                               const float* __restrict__ dA, const float* __restrict__ dJp, const float* __restrict__ dX,
                                               const float* __restrict__ J_dirs, const int* __restrict__ tree, int max_depth, int B,
                                               int betas_stride,
                                               float* __restrict__ dR, float* __restrict__ dbetas, float* __restrict__ dexpr, int b) {
    __shared__ float sdG[NJ][12];     // dL/dG  (3x4)
    __shared__ float sdJ[NJ][3];      // dL/dJrest
    __shared__ float sC[NJ][15];      // per-joint contribution to its parent: dG (12) + dJrest[parent] (3)
    __shared__ float sG[NJ][12];      // global transforms of this frame
    __shared__ float sJr[NJ][3];
    __shared__ float sR[NJ][9];
    __shared__ float sdR[NJ][9];      // chain part of dR; the blend-shape part dX is added by all threads after the walk
    const int j = threadIdx.x;
    __shared__ int s_tree[TREE_SMEM ? 1 : TREE_N];
    const int* T = stage_tree<TREE_SMEM>(tree, s_tree, SUB64 ? 64 : blockDim.x);
    BODY_STAMP(8);
    if (j < NJ) {
        float gl[12], Jr[3];
        for (int k = 0; k < 9; ++k) sR[j][k] = R[((size_t)b * NJ + j) * 9 + k];
        for (int k = 0; k < 12; ++k) sG[j][k] = gl[k] = G[((size_t)b * NJ + j) * 12 + k];
        for (int k = 0; k < 3; ++k) sJr[j][k] = Jr[k] = Jrest[((size_t)b * NJ + j) * 3 + k];
        const float* da = dA + (size_t)j * B * 12 + (size_t)b * 12;
        float dAt[3] = {da[3], da[7], da[11]};
        // A.R = G.R ; A.t = G.t - G.R J
        for (int i = 0; i < 3; ++i) {
            for (int c = 0; c < 3; ++c) sdG[j][i * 4 + c] = da[i * 4 + c] - dAt[i] * Jr[c];
            sdG[j][i * 4 + 3] = dAt[i] + (dJp ? dJp[((size_t)b * NJ + j) * 3 + i] : 0.f);
        }
        // dJ += -G.R^T dAt
        if (NEED_J)
            for (int c = 0; c < 3; ++c) sdJ[j][c] = -(gl[c] * dAt[0] + gl[4 + c] * dAt[1] + gl[8 + c] * dAt[2]);
    }
    BODY_STAMP(9);
    body_sync<SUB64>();
    BODY_STAMP(10);
    // children -> parent accumulation, deepest level first, by warp 0 alone (lane i = i-th joint of the level).  When a joint's turn comes
    // it first adds what its children left for it, in ascending child index: FIXED order, no shared-memory atomics -- results are bitwise
    // reproducible, which the sequence-sharding contract relies on (same sequence, any slot / GPU -> same parameters).  One __syncwarp
    // per level.
    if (threadIdx.x < 32) {
        int e = T[TREE_LANE + max_depth * 32 + threadIdx.x];
        for (int lev = max_depth; lev >= 0; --lev) {
            const int e_next = lev > 0 ? T[TREE_LANE + (lev - 1) * 32 + threadIdx.x] : -1;
            if (e >= 0) {
                const int jj = e & 255, par = ((e >> 8) & 255) - 1, k0 = (e >> 16) & 255, nk = (e >> 24) & 255;
                float dG[12];
                for (int k = 0; k < 12; ++k) dG[k] = sdG[jj][k];
                for (int q = 0; q < nk; ++q) {
                    const int c = T[TREE_KLIST + k0 + q];
                    for (int k = 0; k < 12; ++k) dG[k] += sC[c][k];
                }
                if (par >= 0) {
                    const float* gp = sG[par];                            // parent's global transform
                    const float gpR[9] = {gp[0], gp[1], gp[2], gp[4], gp[5], gp[6], gp[8], gp[9], gp[10]};
                    const float dGr[9] = {dG[0], dG[1], dG[2], dG[4], dG[5], dG[6], dG[8], dG[9], dG[10]};
                    const float dGt[3] = {dG[3], dG[7], dG[11]};
                    float rr[9];
                    for (int k = 0; k < 9; ++k) rr[k] = sR[jj][k];
                    const float t[3] = {sJr[jj][0] - sJr[par][0], sJr[jj][1] - sJr[par][1], sJr[jj][2] - sJr[par][2]};
                    float dr[9], dpr[9], dt[3];
                    m3_mul_at(gpR, dGr, dr);               // dR_j = Gp.R^T dG_j.R
                    m3_mul_bt(dGr, rr, dpr);               // dGp.R += dG_j.R R_j^T
                    for (int k = 0; k < 9; ++k) sdR[jj][k] = dr[k];
                    for (int q = 0; q < 3; ++q) {
                        for (int c = 0; c < 3; ++c) sC[jj][q * 4 + c] = dpr[q * 3 + c] + dGt[q] * t[c];
                        sC[jj][q * 4 + 3] = dGt[q];
                    }
                    if (NEED_J) {
                        m3t_vec(gpR, dGt, dt);             // dt = Gp.R^T dG_j.t
                        for (int q = 0; q < 3; ++q) { sC[jj][12 + q] = -dt[q]; sdJ[jj][q] += dt[q]; }
                    }
                } else {                                   // root: its rotation gradient is dG.R itself; the translation feeds its rest joint
                    for (int q = 0; q < 3; ++q) {
                        for (int c = 0; c < 3; ++c) sdR[jj][q * 3 + c] = dG[q * 4 + c];
                        if (NEED_J) sdJ[jj][q] += dG[q * 4 + 3];
                    }
                }
            }
            __syncwarp();
            e = e_next;
        }
    }
    BODY_STAMP(11);
    body_sync<SUB64>();
    BODY_STAMP(12);
    if (NEED_J && j < NJ) {                  // children's share of the rest-joint adjoint (ascending child index)
        for (int q = T[TREE_KOFF + j]; q < T[TREE_KOFF + j + 1]; ++q) {
            const int c = T[TREE_KLIST + q];
            for (int k = 0; k < 3; ++k) sdJ[j][k] += sC[c][12 + k];
        }
    }
    if (j < NJ) {
        float* o = dR + ((size_t)b * NJ + j) * 9;
        if (j == 0) for (int k = 0; k < 9; ++k) o[k] = sdR[0][k];
        else {
            const float* dx = dX + (size_t)b * XK + (j - 1) * 9;
            for (int k = 0; k < 9; ++k) o[k] = sdR[j][k] + dx[k];
        }
    }
    if (NEED_J) body_sync<SUB64>();
    // betas / expression: direct (X columns) + through Jrest (skipped when neither gradient is requested: 165 loads per thread)
    if (NEED_J && j < NBETA && (dbetas || dexpr)) {
        float acc = dX[(size_t)b * XK + NPF + j];
        for (int q = 0; q < NJ; ++q)
            for (int k = 0; k < 3; ++k) acc = fmaf(J_dirs[(q * 3 + k) * NBETA + j], sdJ[q][k], acc);
        if (j < 10) {
            if (dbetas) {
                if (betas_stride == 0) atomicAdd(&dbetas[j], acc);
                else dbetas[(size_t)b * 10 + j] = acc;
            }
        } else if (dexpr) dexpr[(size_t)b * 10 + (j - 10)] = acc;
    }
    BODY_STAMP(13);
}

}  // namespace lemo
