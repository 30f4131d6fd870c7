// Device bodies of the per-frame pose / kinematic-chain kernels: one CTA (any size >= 64 threads, thread j = joint j) per frame b.
// Shared by the stand-alone kernels of body.cu and by the persistent per-frame fitting kernel (perframe_mega.cuh), which calls them
// from inside its iteration loop so that both paths run the SAME arithmetic.  Every thread of the CTA must call them (barriers inside).
#pragma once
#include "body.cuh"

namespace lemo {

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
    body_sync<SUB64>();
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
template <bool SUB64 = false>
__device__ __forceinline__ void pose_chain_fwd_body(const PoseK& p, const float* __restrict__ J_template,
                                                    const float* __restrict__ J_dirs, const int* __restrict__ parents,
                                                    const int* __restrict__ depth, int max_depth, float* __restrict__ full_pose,
                                                    float* __restrict__ R, float* __restrict__ X, float* __restrict__ X2,
                                                    float* __restrict__ G,
                                                    float* __restrict__ A, float* __restrict__ Jrest, float* __restrict__ Jposed,
                                                    float* __restrict__ A2, int b) {
    __shared__ float sG[NJ][12];
    __shared__ float sJ[NJ][3];
    __shared__ float sbeta[NBETA];
    __shared__ float sX[XK];
    const int j = threadIdx.x;
    if (j < NBETA) {
        float v = 0.f;
        if (j < 10) v = p.in.betas ? p.in.betas[(size_t)b * p.in.betas_stride + j] : 0.f;
        else v = p.in.expression ? p.in.expression[b * 10 + (j - 10)] : 0.f;
        sbeta[j] = v;
        sX[NPF + j] = v;
    }
    if (j >= NBETA && j < XK - NPF) sX[NPF + j] = 0.f;
    float r[9];
    int par = -1, dep = 0;
    if (j < NJ) {
        float aa[3];
        const float* Rov = nullptr;
        if (j == 0 && p.in.R_global) Rov = p.in.R_global + b * 9;
        if (j >= 1 && j <= NBODY && p.in.R_body) Rov = p.in.R_body + (b * NBODY + (j - 1)) * 9;
        if (Rov) {
            for (int k = 0; k < 9; ++k) r[k] = Rov[k];
            rotmat_to_aa_tgm(r, aa);                 // what the reference scripts store in the [T,72] result
        } else {
            joint_aa(p, b, j, aa);
            rodrigues_fwd(aa, r);
        }
        for (int k = 0; k < 3; ++k) full_pose[b * 165 + j * 3 + k] = aa[k];
        for (int k = 0; k < 9; ++k) R[((size_t)b * NJ + j) * 9 + k] = r[k];
        if (j >= 1)
            for (int k = 0; k < 9; ++k) sX[(j - 1) * 9 + k] = r[k] - ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
    }
    body_sync<SUB64>();
    for (int k = threadIdx.x; k < XK; k += (SUB64 ? 64 : blockDim.x)) {
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
        par = parents[j]; dep = depth[j];
    }
    body_sync<SUB64>();
    for (int lev = 0; lev <= max_depth; ++lev) {
        if (j < NJ && dep == lev) {
            float g[12];
            if (lev == 0) {
                for (int i = 0; i < 3; ++i) { g[i * 4] = r[i * 3]; g[i * 4 + 1] = r[i * 3 + 1]; g[i * 4 + 2] = r[i * 3 + 2]; g[i * 4 + 3] = sJ[j][i]; }
            } else {
                const float* gp = sG[par];
                const float t[3] = {sJ[j][0] - sJ[par][0], sJ[j][1] - sJ[par][1], sJ[j][2] - sJ[par][2]};
                for (int i = 0; i < 3; ++i) {
                    for (int c = 0; c < 3; ++c)
                        g[i * 4 + c] = gp[i * 4] * r[c] + gp[i * 4 + 1] * r[3 + c] + gp[i * 4 + 2] * r[6 + c];
                    g[i * 4 + 3] = gp[i * 4] * t[0] + gp[i * 4 + 1] * t[1] + gp[i * 4 + 2] * t[2] + gp[i * 4 + 3];
                }
            }
            for (int k = 0; k < 12; ++k) sG[j][k] = g[k];
        }
        body_sync<SUB64>();
    }
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
}

// adjoint of k_chain_fwd.  dA is joint-major [55][B*12]; dJp [B,55,3]; dX [B,512].
// Writes dR [B,55,9]; betas/expression grads (per frame, or atomically into one row when betas_stride==0).
template <bool SUB64 = false>
__device__ __forceinline__ void chain_bwd_body(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ Jrest,
                                               const float* __restrict__ dA, const float* __restrict__ dJp, const float* __restrict__ dX,
                                               const float* __restrict__ J_dirs, const int* __restrict__ parents,
                                               const int* __restrict__ depth, int max_depth, int B, int betas_stride,
                                               float* __restrict__ dR, float* __restrict__ dbetas, float* __restrict__ dexpr, int b) {
    __shared__ float sdG[NJ][12];     // dL/dG  (3x4)
    __shared__ float sdJ[NJ][3];      // dL/dJrest
    __shared__ float sC[NJ][15];      // per-joint contribution to its parent: dG (12) + dJrest[parent] (3)
    __shared__ int spar[NJ];
    __shared__ float sG[NJ][12];      // global transforms of this frame (parents are read from here, not from HBM, inside the level loop)
    __shared__ float sJr[NJ][3];
    const int j = threadIdx.x;
    float r[9], gl[12], Jr[3];
    int par = -1, dep = 0;
    if (j < NJ) {
        for (int k = 0; k < 9; ++k) r[k] = R[((size_t)b * NJ + j) * 9 + k];
        for (int k = 0; k < 12; ++k) sG[j][k] = gl[k] = G[((size_t)b * NJ + j) * 12 + k];
        for (int k = 0; k < 3; ++k) sJr[j][k] = Jr[k] = Jrest[((size_t)b * NJ + j) * 3 + k];
        par = parents[j]; dep = depth[j];
        const float* da = dA + (size_t)j * B * 12 + (size_t)b * 12;
        float dAt[3] = {da[3], da[7], da[11]};
        // A.R = G.R ; A.t = G.t - G.R J
        for (int i = 0; i < 3; ++i) {
            for (int c = 0; c < 3; ++c) sdG[j][i * 4 + c] = da[i * 4 + c] - dAt[i] * Jr[c];
            sdG[j][i * 4 + 3] = dAt[i] + (dJp ? dJp[((size_t)b * NJ + j) * 3 + i] : 0.f);
        }
        // dJ += -G.R^T dAt
        for (int c = 0; c < 3; ++c) sdJ[j][c] = -(gl[c] * dAt[0] + gl[4 + c] * dAt[1] + gl[8 + c] * dAt[2]);
    }
    if (j < NJ) spar[j] = par;
    body_sync<SUB64>();
    unsigned long long kids = 0ull;   // children of joint j as a bit mask, walked in ascending order below
    if (j < NJ)
        for (int c = j + 1; c < NJ; ++c)
            if (spar[c] == j) kids |= 1ull << c;
    // children -> parent accumulation in a FIXED order (no shared-memory atomics): results are bitwise reproducible,
    // which the sequence-sharding contract relies on (same sequence, any slot / GPU -> same parameters).
    for (int lev = max_depth; lev >= 1; --lev) {
        if (j < NJ && dep == lev) {
            const float* gp = sG[par];                                // parent's global transform
            float gpR[9] = {gp[0], gp[1], gp[2], gp[4], gp[5], gp[6], gp[8], gp[9], gp[10]};
            float dGr[9] = {sdG[j][0], sdG[j][1], sdG[j][2], sdG[j][4], sdG[j][5], sdG[j][6], sdG[j][8], sdG[j][9], sdG[j][10]};
            float dGt[3] = {sdG[j][3], sdG[j][7], sdG[j][11]};
            const float* Jp = sJr[par];
            const float t[3] = {Jr[0] - Jp[0], Jr[1] - Jp[1], Jr[2] - Jp[2]};
            float dr[9], dpr[9], dt[3];
            m3_mul_at(gpR, dGr, dr);               // dR_j = Gp.R^T dG_j.R
            m3_mul_bt(dGr, r, dpr);                // dGp.R += dG_j.R R_j^T
            m3t_vec(gpR, dGt, dt);                 // dt = Gp.R^T dG_j.t
            float* o = dR + ((size_t)b * NJ + j) * 9;
            const float* dx = dX + (size_t)b * XK + (j - 1) * 9;
            for (int k = 0; k < 9; ++k) o[k] = dr[k] + dx[k];
            for (int i = 0; i < 3; ++i) {
                for (int c = 0; c < 3; ++c) sC[j][i * 4 + c] = dpr[i * 3 + c] + dGt[i] * t[c];
                sC[j][i * 4 + 3] = dGt[i];
                sC[j][12 + i] = -dt[i];
                sdJ[j][i] += dt[i];
            }
        }
        body_sync<SUB64>();
        if (j < NJ && dep == lev - 1) {
            for (unsigned long long m = kids; m; m &= m - 1) {
                const int c = __ffsll((long long)m) - 1;
                for (int k = 0; k < 12; ++k) sdG[j][k] += sC[c][k];
                for (int k = 0; k < 3; ++k) sdJ[j][k] += sC[c][12 + k];
            }
        }
        body_sync<SUB64>();
    }
    if (j == 0) {
        float* o = dR + ((size_t)b * NJ) * 9;
        for (int i = 0; i < 3; ++i) {
            for (int c = 0; c < 3; ++c) o[i * 3 + c] = sdG[0][i * 4 + c];
            sdJ[0][i] += sdG[0][i * 4 + 3];
        }
    }
    body_sync<SUB64>();
    // betas / expression: direct (X columns) + through Jrest (skipped when neither gradient is requested: 165 loads per thread)
    if (j < NBETA && (dbetas || dexpr)) {
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
}


}  // namespace lemo
