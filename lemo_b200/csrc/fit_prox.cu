// Fused PROX stage-2 fitting driver: the closure + Adam loop of the reference's
//   temp_prox/fitting_temp_slide.py:169-313   (FittingMonitor.run_fitting / create_fitting_closure.fitting_func)
//   temp_prox/fitting_temp_slide.py:564-1062  (SMPLifyLoss.forward, the terms cfg_files/PROXD_temp_S2.yaml switches on, plus the
//                                              Chamfer `contact` term BASELINE config 4 names)
// for one B-frame window, entirely on the device: VPoser decode, full-mesh SMPL-X ONCE (the reference evaluates it twice,
// :248-258), 2-D keypoint term through the fixed PerspectiveCamera, L2 / angle priors, camera->world, SDF penetration
// (one shared 256^3 volume; the reference replicates it B times = 6.4 GB), friction, contact (nearest scene point, one
// direction: the loss never reads dist2), Enc smoothness prior on the world markers, the first-15 % gradient erase (:281-288)
// and Adam.  No host synchronisation (the reference has ~14 .item() per step); loss weights, learning rate and the erase
// count live in device memory so that one iteration is ONE replayable CUDA graph across optimisation stages.
// Every reduction has a fixed order (per-CTA partials summed in CTA order, gather-style adjoints instead of float atomics, the body
// adjoint's per-frame partials and dX K-slices added in order -- body.cu), so a window is bitwise reproducible; the one exception is
// the reported smoothness LOSS VALUE (float atomics in the Enc loss kernel), which does not feed back into the parameters.
#include "common.cuh"
#include "body.cuh"
#include "vposer.cuh"
#include "conv.cuh"
#include "gemm.cuh"
#include "fit_common.cuh"
#include "prox_common.cuh"
#include "chamfer.cuh"
#include "../../include/lemo_b200.h"
#include <vector>
#include <algorithm>

namespace lemo {

struct ProxDev {             // device-resident knobs (changed between stages without re-capturing the graph)
    LemoProxWeightsC w;
    int erase_n;             // frames [0, erase_n) keep their parameters (first_batch_flag == False: int(bs * 0.15))
};

enum { PT_JOINT = 0, PT_PPRIOR, PT_SHAPE, PT_ANGLE, PT_HAND, PT_EXPR, PT_JAW, PT_SDF, PT_FRIC_T, PT_FRIC_N, PT_CONTACT, PT_SMOOTH,
       PT_TOTAL = 15, PT_N = 16 };
constexpr int PARTS = 2048;                  // per-CTA partial sums per term
constexpr int PP = 81;                       // optimised parameters per frame
// offsets (x B) inside the flat parameter vector
constexpr int O_TR = 0, O_GO = 3, O_Z = 6, O_LH = 38, O_RH = 50, O_JAW = 62, O_LE = 65, O_RE = 68, O_EX = 71;

struct ProxFit {
    int device = 0, B = 0, Jm = 0, nout = 0;
    const Model* model = nullptr;
    BodyCtx* ctx = nullptr;
    VPoser* vp = nullptr;
    ConvNet* enc = nullptr;
    PlaneGeom geom{};
    Cam cam{}, c2w{};
    Grid grid{};
    const float* sdf = nullptr;              // caller-owned [D,D,D]
    const float* scene = nullptr;            // caller-owned [m,3]
    int n_scene = 0;
    bool has_sdf = false, has_fric = false, has_contact = false, has_smooth = false;
    int n_fric = 0, n_contact = 0, NRW = 0;  // world rows: [81 markers | fric | contact]
    int off_fric = 81, off_contact = 81;
    int n_uniq = 0;
    int *row_ids = nullptr, *uniq_ids = nullptr, *uniq_off = nullptr, *uniq_rows = nullptr;
    int *inv_off = nullptr, *inv_idx = nullptr;      // joint (of nout) -> mapped keypoint slots (CSR)
    float *P = nullptr, *Gp = nullptr, *M1 = nullptr, *M2 = nullptr, *betas = nullptr;
    float *gt = nullptr, *conf = nullptr, *jw = nullptr;
    float *Rb = nullptr, *dRb = nullptr;
    float *verts = nullptr, *joints = nullptr, *d_joints = nullptr;
    float *Wrows = nullptr, *Grows = nullptr, *sdf_fric = nullptr;
    float *cdist = nullptr; int* cidx = nullptr;
    float *xin = nullptr, *gx = nullptr, *gv = nullptr, *canon = nullptr, *stats = nullptr;
    float *part = nullptr, *losses = nullptr, *fpart = nullptr;
    ProxDev* dev = nullptr;
    Sched* sched = nullptr;
    SceneGrid* sgrid = nullptr;              // uniform grid over the scene points (exact NN, chamfer.cu)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t gstream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    int use_graph = 1;
    float smooth_w_host = -1.f;               // motion_prior_smooth_weight as baked into the captured Enc loss kernel
    long long launches = 0, launches_per_iter = 0;
    float* p(int off) const { return P + (size_t)B * off; }
    float* g(int off) const { return Gp + (size_t)B * off; }
};

// ------------------------------------------------------------------------------------------------ keypoints
// joint_loss = mean(weights^2 * |gt - camera(joints[:, map])|) * data_weight      (fitting_temp_slide.py:573-581, camera.py:93-116)
// One thread per (frame, model joint): projects once, serves every keypoint slot mapped to that joint (the OpenPose map lists the
// wrists twice), and WRITES d_joints -- no atomics.
__global__ void __launch_bounds__(128) k_prox_keypoints(const float* __restrict__ joints, int B, int nout, int Jm, const int* __restrict__ inv_off,
                                                        const int* __restrict__ inv_idx, const float* __restrict__ gt,
                                                        const float* __restrict__ conf, const float* __restrict__ jw, Cam c,
                                                        const ProxDev* __restrict__ dv, float* __restrict__ d_joints, float* __restrict__ part) {
    __shared__ float sred[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f;
    if (i < B * nout) {
        const int b = i / nout, j = i - b * nout;
        const float* p = joints + (size_t)i * 3;
        float q[3];
        cam_apply(c, p, q);
        const float u = c.fx * (q[0] / q[2]) + c.cx, v = c.fy * (q[1] / q[2]) + c.cy;
        float gu = 0.f, gv = 0.f;
        const bool use_conf = dv->w.use_joints_conf != 0;
        for (int e = inv_off[j]; e < inv_off[j + 1]; ++e) {
            const size_t s = (size_t)b * Jm + inv_idx[e];
            float w = jw[s];
            if (use_conf) w *= conf[s];
            const float w2 = w * w;
            const float du = gt[s * 2] - u, dw = gt[s * 2 + 1] - v;
            loss += w2 * (fabsf(du) + fabsf(dw));
            gu -= w2 * (du > 0.f ? 1.f : (du < 0.f ? -1.f : 0.f));
            gv -= w2 * (dw > 0.f ? 1.f : (dw < 0.f ? -1.f : 0.f));
        }
        const float scale = dv->w.data_weight / ((float)B * (float)Jm * 2.f);
        gu *= scale * c.fx; gv *= scale * c.fy;
        const float dq[3] = {gu / q[2], gv / q[2], -(gu * q[0] + gv * q[1]) / (q[2] * q[2])};
        float dp[3];
        cam_apply_t(c, dq, dp);
        d_joints[(size_t)i * 3] = dp[0]; d_joints[(size_t)i * 3 + 1] = dp[1]; d_joints[(size_t)i * 3 + 2] = dp[2];
    }
    loss = block_sum(loss, sred);
    if (threadIdx.x == 0) part[PT_JOINT * PARTS + blockIdx.x] = loss * (dv->w.data_weight / ((float)B * (float)Jm * 2.f));
}

// ------------------------------------------------------------------------------------------------ scene: SDF penetration on every vertex
// vertices_world = R v + t; body_sdf = grid_sample(sdf, ...); loss = w * sum_{sdf<0} |sdf|      (fitting_temp_slide.py:673-694)
// Writes the dense vertex gradient (camera coordinates) for ALL vertices: it is the buffer the skinning adjoint consumes.
__global__ void __launch_bounds__(256) k_prox_scene(const float* __restrict__ verts, long long n, Cam c2w, const float* __restrict__ sdf, Grid g,
                                                    int has_sdf, const ProxDev* __restrict__ dv, float* __restrict__ Gv, float* __restrict__ part) {
    __shared__ float sred[32];
    const float w = dv->w.sdf_penetration_weight;
    float pen = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gc[3] = {0.f, 0.f, 0.f};
        if (has_sdf && w > 0.f) {
            const float p[3] = {verts[i * 3], verts[i * 3 + 1], verts[i * 3 + 2]};
            float pw[3], d[3];
            cam_apply(c2w, p, pw);
            const float s = sdf_eval<true>(sdf, g, pw, d);
            if (s < 0.f) {
                pen -= s;
                const float gw[3] = {-w * d[0], -w * d[1], -w * d[2]};
                cam_apply_t(c2w, gw, gc);
            }
        }
        Gv[i * 3] = gc[0]; Gv[i * 3 + 1] = gc[1]; Gv[i * 3 + 2] = gc[2];
    }
    pen = block_sum(pen, sred);
    if (threadIdx.x == 0) part[PT_SDF * PARTS + blockIdx.x] = pen * w;
}

// world coordinates of the loss rows [81 smoothness markers | friction vertices | contact vertices] + SDF at the friction vertices
__global__ void k_prox_rows(const float* __restrict__ verts, int B, int V, const int* __restrict__ row_ids, int NRW, int off_fric, int n_fric,
                            Cam c2w, const float* __restrict__ sdf, Grid g, float* __restrict__ Wrows, float* __restrict__ sdf_fric) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NRW) return;
    const int b = i / NRW, r = i - b * NRW;
    const float* p = verts + ((size_t)b * V + row_ids[r]) * 3;
    float pw[3];
    cam_apply(c2w, p, pw);
    Wrows[(size_t)i * 3] = pw[0]; Wrows[(size_t)i * 3 + 1] = pw[1]; Wrows[(size_t)i * 3 + 2] = pw[2];
    if (sdf && r >= off_fric && r < off_fric + n_fric) sdf_fric[(size_t)b * n_fric + (r - off_fric)] = sdf_eval<false>(sdf, g, pw, nullptr);
}

// friction (fitting_temp_slide.py:699-739), scene normal = +z: among the (frame, vertex) pairs with sdf < 0.01,
//   tangent: mean |v_xy| over those with |v_xy| > 1e-4   * w_t ;  normal: mean |v_z| over those with v_z < 0   * w_n
// Pass 1 (FR_G CTAs): counts / sums as per-CTA partials.  Pass 2: every CTA adds the partials in CTA order (fixed order: bitwise
// reproducible) and writes the world-row gradient gather-style (row b gets + from velocity b-1 and - from velocity b: no atomics).
constexpr int FR_G = 32;
__global__ void __launch_bounds__(256) k_prox_fric_count(const float* __restrict__ Wrows, const float* __restrict__ sdf_fric, int B, int NRW, int off,
                                                         int nf, float* __restrict__ fpart /*[4][FR_G]*/) {
    __shared__ float sred[32];
    float ct = 0.f, st = 0.f, cn = 0.f, sn = 0.f;
    const int n = (B - 1) * nf;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int b = i / nf, r = i - b * nf;
        if (!(sdf_fric[(size_t)b * nf + r] < 0.01f)) continue;
        const float* a0 = Wrows + ((size_t)b * NRW + off + r) * 3;
        const float* a1 = Wrows + ((size_t)(b + 1) * NRW + off + r) * 3;
        const float vx = a1[0] - a0[0], vy = a1[1] - a0[1], vz = a1[2] - a0[2];
        const float vt = sqrtf(vx * vx + vy * vy);
        if (vt > 1e-4f) { ct += 1.f; st += vt; }
        if (vz < 0.f) { cn += 1.f; sn -= vz; }
    }
    ct = block_sum(ct, sred); if (threadIdx.x == 0) fpart[0 * FR_G + blockIdx.x] = ct;
    st = block_sum(st, sred); if (threadIdx.x == 0) fpart[1 * FR_G + blockIdx.x] = st;
    cn = block_sum(cn, sred); if (threadIdx.x == 0) fpart[2 * FR_G + blockIdx.x] = cn;
    sn = block_sum(sn, sred); if (threadIdx.x == 0) fpart[3 * FR_G + blockIdx.x] = sn;
}
__global__ void __launch_bounds__(256) k_prox_fric_grad(const float* __restrict__ Wrows, const float* __restrict__ sdf_fric, int B, int NRW, int off,
                                                        int nf, const float* __restrict__ fpart, const ProxDev* __restrict__ dv,
                                                        float* __restrict__ Grows, float* __restrict__ part) {
    __shared__ float s_val[4];
    if (threadIdx.x < 4) {
        float a = 0.f;
        for (int g = 0; g < FR_G; ++g) a += fpart[threadIdx.x * FR_G + g];
        s_val[threadIdx.x] = a;
    }
    __syncthreads();
    const float wt = dv->w.friction_tangent_weight, wn = dv->w.friction_normal_weight;
    const float it = s_val[0] >= 1.f ? wt / s_val[0] : 0.f, in = s_val[2] >= 1.f ? wn / s_val[2] : 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        part[PT_FRIC_T * PARTS] = s_val[0] >= 1.f ? wt * s_val[1] / s_val[0] : 0.f;
        part[PT_FRIC_N * PARTS] = s_val[2] >= 1.f ? wn * s_val[3] / s_val[2] : 0.f;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * nf) return;
    const int b = i / nf, r = i - b * nf;
    float gsum[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 2; ++e) {                     // e = 0: velocity b-1 (this row is its head, +) ; e = 1: velocity b (tail, -)
        const int vb = b - 1 + e;
        if (vb < 0 || vb > B - 2) continue;
        if (!(sdf_fric[(size_t)vb * nf + r] < 0.01f)) continue;
        const float* a0 = Wrows + ((size_t)vb * NRW + off + r) * 3;
        const float* a1 = Wrows + ((size_t)(vb + 1) * NRW + off + r) * 3;
        const float vx = a1[0] - a0[0], vy = a1[1] - a0[1], vz = a1[2] - a0[2];
        const float vt = sqrtf(vx * vx + vy * vy);
        const float sg = e == 0 ? 1.f : -1.f;
        if (vt > 1e-4f) { gsum[0] += sg * it * vx / vt; gsum[1] += sg * it * vy / vt; }
        if (vz < 0.f) gsum[2] -= sg * in;
    }
    float* o = Grows + ((size_t)b * NRW + off + r) * 3;
    o[0] = gsum[0]; o[1] = gsum[1]; o[2] = gsum[2];
}

// contact (fitting_temp_slide.py:743-753): d = squared distance to the nearest scene point; loss = w * mean(r / (r + 1)), r = sqrt(d + 1e-4)
__global__ void __launch_bounds__(256) k_prox_contact(const float* __restrict__ Wrows, const float* __restrict__ scene, const float* __restrict__ cdist,
                                                      const int* __restrict__ cidx, int B, int NRW, int off, int nc,
                                                      const ProxDev* __restrict__ dv, float* __restrict__ Grows, float* __restrict__ part) {
    __shared__ float sred[32];
    float acc = 0.f;
    const float w = dv->w.contact_loss_weight / ((float)B * (float)nc);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * nc; i += gridDim.x * blockDim.x) {
        const int b = i / nc, r = i - b * nc;
        const float rr = sqrtf(cdist[i] + 1e-4f);
        acc += rr / (rr + 1.f);
        const float coef = w / ((rr + 1.f) * (rr + 1.f)) / (2.f * rr) * 2.f;       // dL/dd * d(d)/d(v - s)
        const float* v = Wrows + ((size_t)b * NRW + off + r) * 3;
        const float* s = scene + (size_t)cidx[i] * 3;
        float* o = Grows + ((size_t)b * NRW + off + r) * 3;
        o[0] = coef * (v[0] - s[0]); o[1] = coef * (v[1] - s[1]); o[2] = coef * (v[2] - s[2]);
    }
    acc = block_sum(acc, sred);
    if (threadIdx.x == 0) part[PT_CONTACT * PARTS + blockIdx.x] = acc * w;
}

// canonical frame of the smoothness prior in WORLD coordinates (fitting_temp_slide.py:1003-1014), detached
__global__ void k_prox_canon(const float* __restrict__ joints, const float* __restrict__ Wrows, Cam c2w, float* __restrict__ canon) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float d[3] = {joints[2 * 3] - joints[1 * 3], joints[2 * 3 + 1] - joints[1 * 3 + 1], joints[2 * 3 + 2] - joints[1 * 3 + 2]};
    float x0 = c2w.R[0] * d[0] + c2w.R[1] * d[1] + c2w.R[2] * d[2];
    float x1 = c2w.R[3] * d[0] + c2w.R[4] * d[1] + c2w.R[5] * d[2];
    const float nx = sqrtf(x0 * x0 + x1 * x1);
    x0 /= nx; x1 /= nx;
    float y0 = -x1, y1 = x0;
    const float ny = sqrtf(y0 * y0 + y1 * y1);
    y0 /= ny; y1 /= ny;
    canon[0] = x0; canon[1] = y0; canon[2] = 0.f;
    canon[3] = x1; canon[4] = y1; canon[5] = 0.f;
    canon[6] = 0.f; canon[7] = 0.f; canon[8] = 1.f;
    canon[9] = Wrows[0]; canon[10] = Wrows[1]; canon[11] = Wrows[2];
}

// world-row gradients -> dense vertex gradient (camera coordinates): one thread per (frame, distinct vertex) sums the rows that
// refer to it in table order (a marker vertex may also be a contact vertex) and is the only writer of that vertex
__global__ void k_prox_rows_bwd(const float* __restrict__ Grows, int B, int V, int NRW, int n_uniq, const int* __restrict__ uniq_ids,
                                const int* __restrict__ uniq_off, const int* __restrict__ uniq_rows, Cam c2w, float* __restrict__ Gv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n_uniq) return;
    const int b = i / n_uniq, u = i - b * n_uniq;
    float gw[3] = {0.f, 0.f, 0.f};
    for (int e = uniq_off[u]; e < uniq_off[u + 1]; ++e) {
        const float* g = Grows + ((size_t)b * NRW + uniq_rows[e]) * 3;
        gw[0] += g[0]; gw[1] += g[1]; gw[2] += g[2];
    }
    float gc[3];
    cam_apply_t(c2w, gw, gc);
    float* o = Gv + ((size_t)b * V + uniq_ids[u]) * 3;
    o[0] += gc[0]; o[1] += gc[1]; o[2] += gc[2];
}

// angle prior on the elbows / knees (prior.py:53-89, fitting_temp_slide.py:594-596): sum exp(sign * full_pose[:, {55,58,12,15}]) * bending^2.
// The body pose enters the model as rotation matrices, so the gradient is pulled back through tgm's R -> aa (rotmat_to_aa_tgm_bwd).
__global__ void __launch_bounds__(256) k_prox_angle(const float* __restrict__ full_pose, const float* __restrict__ Rb, int B,
                                                    const ProxDev* __restrict__ dv, float* __restrict__ dRb, float* __restrict__ part) {
    __shared__ float sred[32];
    const float w2 = dv->w.bending_prior_weight * dv->w.bending_prior_weight;
    float acc = 0.f;
    for (int i = threadIdx.x; i < B * 4; i += blockDim.x) {
        const int b = i >> 2, k = i & 3;
        const int col = k == 0 ? 55 : (k == 1 ? 58 : (k == 2 ? 12 : 15));
        const float sg = k == 0 ? 1.f : -1.f;
        const int jf = col / 3, comp = col - jf * 3;
        const float e = expf(full_pose[(size_t)b * 165 + col] * sg);
        acc += e;
        float daa[3] = {0.f, 0.f, 0.f}, dR[9];
        daa[comp] = sg * e * w2;
        const float* R = Rb + ((size_t)b * NBODY + (jf - 1)) * 9;
        rotmat_to_aa_tgm_bwd(R, daa, dR);
        float* o = dRb + ((size_t)b * NBODY + (jf - 1)) * 9;
#pragma unroll
        for (int q = 0; q < 9; ++q) o[q] += dR[q];
    }
    acc = block_sum(acc, sred);
    if (threadIdx.x == 0) part[PT_ANGLE * PARTS] = acc * w2;
}

// L2 priors (fitting_temp_slide.py:585-616): sum(z^2) w_bp^2, sum(betas^2) w_shape^2 (value only: betas are not optimised),
// sum(lh^2 + rh^2) w_hand^2, sum(expr^2) w_expr^2, sum((jaw w_jaw)^2); adds their gradients.
__global__ void __launch_bounds__(1024) k_prox_priors(const float* __restrict__ P, float* __restrict__ Gp, const float* __restrict__ betas, int B,
                                                      const ProxDev* __restrict__ dv, float* __restrict__ part) {
    __shared__ float sred[32];
    const LemoProxWeightsC w = dv->w;
    float sz = 0.f, sh = 0.f, se = 0.f, sj = 0.f, sb = 0.f;
    const float wz = w.body_pose_weight * w.body_pose_weight, wh = w.hand_prior_weight * w.hand_prior_weight,
                we = w.expr_prior_weight * w.expr_prior_weight, wj = w.jaw_prior_weight * w.jaw_prior_weight;
    for (int i = threadIdx.x; i < B * 32; i += blockDim.x) { const size_t j = (size_t)B * O_Z + i; const float x = P[j]; sz += x * x; Gp[j] += 2.f * x * wz; }
    for (int i = threadIdx.x; i < B * 24; i += blockDim.x) { const size_t j = (size_t)B * O_LH + i; const float x = P[j]; sh += x * x; Gp[j] += 2.f * x * wh; }
    for (int i = threadIdx.x; i < B * 10; i += blockDim.x) { const size_t j = (size_t)B * O_EX + i; const float x = P[j]; se += x * x; Gp[j] += 2.f * x * we; }
    for (int i = threadIdx.x; i < B * 3; i += blockDim.x) { const size_t j = (size_t)B * O_JAW + i; const float x = P[j]; sj += x * x; Gp[j] += 2.f * x * wj; }
    for (int i = threadIdx.x; i < B * 10; i += blockDim.x) { const float x = betas[i]; sb += x * x; }
    sz = block_sum(sz, sred); if (threadIdx.x == 0) part[PT_PPRIOR * PARTS] = sz * wz;
    sh = block_sum(sh, sred); if (threadIdx.x == 0) part[PT_HAND * PARTS] = sh * wh;
    se = block_sum(se, sred); if (threadIdx.x == 0) part[PT_EXPR * PARTS] = se * we;
    sj = block_sum(sj, sred); if (threadIdx.x == 0) part[PT_JAW * PARTS] = sj * wj;
    sb = block_sum(sb, sred); if (threadIdx.x == 0) part[PT_SHAPE * PARTS] = sb * w.shape_weight * w.shape_weight;
}

// `body_param.grad[0:erase_n, :] = 0` for every optimised tensor (fitting_temp_slide.py:281-288)
__global__ void k_prox_erase(float* __restrict__ Gp, int B, const ProxDev* __restrict__ dv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * PP) return;
    const int offs[10] = {O_TR, O_GO, O_Z, O_LH, O_RH, O_JAW, O_LE, O_RE, O_EX, PP};
    int k = 0;
    while (i >= B * offs[k + 1]) ++k;
    const int width = offs[k + 1] - offs[k];
    const int b = (i - B * offs[k]) / width;
    if (b < dv->erase_n) Gp[i] = 0.f;
}

// fixed-order sum of the per-CTA partials -> losses[16]
struct PartCount { int n[PT_N]; };
__global__ void __launch_bounds__(256) k_prox_finish(const float* __restrict__ part, PartCount pc, const ProxDev* __restrict__ dv,
                                                     float* __restrict__ losses) {
    __shared__ float sred[32];
    float total = 0.f;
    for (int t = 0; t < PT_TOTAL; ++t) {
        float a = 0.f;
        for (int i = threadIdx.x; i < pc.n[t]; i += blockDim.x) a += part[t * PARTS + i];
        a = block_sum(a, sred);
        if (t == PT_SMOOTH) a *= dv->w.motion_prior_smooth_weight;      // the Enc loss kernel accumulates the unweighted mean
        if (threadIdx.x == 0) { losses[t] = a; total += a; }
    }
    if (threadIdx.x == 0) losses[PT_TOTAL] = total;
}

// ------------------------------------------------------------------------------------------------ one closure step
static int prox_iteration(ProxFit* f, bool with_adam, cudaStream_t st) {
    const int B = f->B, V = f->model->V, NRW = f->NRW;
    long long nl = 0;
    PartCount pc{};
    k_sched<<<1, 1, 0, st>>>(f->sched); nl++;
    LEMO_CUDA(cudaMemsetAsync(f->Grows, 0, (size_t)B * NRW * 3 * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(f->part, 0, (size_t)PT_N * PARTS * sizeof(float), st));
    // ---------------- forward: VPoser -> SMPL-X (full mesh), once
    LEMO_TRY(vposer_decode(f->vp, f->p(O_Z), B, f->Rb, nullptr, st)); nl += 4;
    PoseIn in;
    in.transl = f->p(O_TR); in.global_orient = f->p(O_GO); in.R_body = f->Rb; in.jaw = f->p(O_JAW); in.leye = f->p(O_LE); in.reye = f->p(O_RE);
    in.lhand = f->p(O_LH); in.rhand = f->p(O_RH); in.betas = f->betas; in.betas_stride = 10; in.expression = f->p(O_EX); in.hand_is_pca = 1;
    LEMO_TRY(body_pose_forward(f->ctx, in, B, st)); nl++;
    LEMO_TRY(body_skin_forward(f->ctx, f->ctx, in, B, f->verts, f->joints, st)); nl += 3;
    // ---------------- loss terms and their gradients on vertices / joints
    const int kp_blocks = cdiv(B * f->nout, 128);
    LEMO_CHECK(kp_blocks <= PARTS, "window too large for the partial-sum scratch");
    k_prox_keypoints<<<kp_blocks, 128, 0, st>>>(f->joints, B, f->nout, f->Jm, f->inv_off, f->inv_idx, f->gt, f->conf, f->jw, f->cam, f->dev,
                                                 f->d_joints, f->part); nl++;
    pc.n[PT_JOINT] = kp_blocks;
    const int sc_blocks = std::min(PARTS, cdiv((long long)B * V, 256));
    k_prox_scene<<<sc_blocks, 256, 0, st>>>(f->verts, (long long)B * V, f->c2w, f->sdf, f->grid, f->has_sdf ? 1 : 0, f->dev, f->ctx->Gv, f->part); nl++;
    pc.n[PT_SDF] = sc_blocks;
    k_prox_rows<<<cdiv(B * NRW, 128), 128, 0, st>>>(f->verts, B, V, f->row_ids, NRW, f->off_fric, f->has_fric ? f->n_fric : 0, f->c2w,
                                                     f->has_fric ? f->sdf : nullptr, f->grid, f->Wrows, f->sdf_fric); nl++;
    if (f->has_fric) {
        k_prox_fric_count<<<FR_G, 256, 0, st>>>(f->Wrows, f->sdf_fric, B, NRW, f->off_fric, f->n_fric, f->fpart); nl++;
        k_prox_fric_grad<<<cdiv(B * f->n_fric, 256), 256, 0, st>>>(f->Wrows, f->sdf_fric, B, NRW, f->off_fric, f->n_fric, f->fpart, f->dev,
                                                                    f->Grows, f->part); nl++;
        pc.n[PT_FRIC_T] = pc.n[PT_FRIC_N] = 1;
    }
    if (f->has_contact) {
        LEMO_TRY(scene_grid_query(f->sgrid, f->Wrows + (size_t)f->off_contact * 3, (long long)NRW * 3, f->n_contact, B, f->cdist, f->cidx, st)); nl++;
        const int cb = std::min(PARTS, cdiv(B * f->n_contact, 256));
        k_prox_contact<<<cb, 256, 0, st>>>(f->Wrows, f->scene, f->cdist, f->cidx, B, NRW, f->off_contact, f->n_contact, f->dev, f->Grows, f->part); nl++;
        pc.n[PT_CONTACT] = cb;
    }
    if (f->has_smooth) {
        const PlaneGeom& g = f->geom;
        k_prox_canon<<<1, 32, 0, st>>>(f->joints, f->Wrows, f->c2w, f->canon); nl++;
        k_smooth_input<<<dim3(cdiv(g.W, 128), g.H, 1), 128, 0, st>>>(f->Wrows, f->canon, f->stats, B, NRW, g.H, g.W, g.Wp, g.PS, f->xin); nl++;
        LEMO_TRY(enc_forward_planes(f->enc, f->xin, 1, st)); nl += 10;
        // the smoothness weight is read on the host when the graph is (re)captured: set_weights invalidates the graph
        if (enc_uses_tc(f->enc)) LEMO_TRY(enc_tc_smooth_loss(f->enc, 1, f->smooth_w_host, PARTS, 0, f->part + PT_SMOOTH * PARTS, st));
        else k_smooth_loss<<<dim3(8, 64, 1), 256, 0, st>>>(enc_z_planes(f->enc), 64, g.H, g.W, g.Wp, g.PS, f->smooth_w_host, enc_gz_planes(f->enc),
                                                           f->part + PT_SMOOTH * PARTS, PARTS, 0);
        nl++;
        pc.n[PT_SMOOTH] = 1;
        LEMO_TRY(enc_backward_planes(f->enc, 1, f->gx, st)); nl += 10;
        k_smooth_bwd_a<<<dim3(cdiv(B - 1, 128), 243, 1), 128, 0, st>>>(f->gx, B, g.H, g.W, g.Wp, g.PS, f->gv); nl++;
        k_smooth_bwd_b<<<cdiv(B * 81, 256), 256, 0, st>>>(f->gv, f->canon, f->stats, B, NRW, 1, f->Grows); nl++;
    }
    k_prox_rows_bwd<<<cdiv(B * f->n_uniq, 128), 128, 0, st>>>(f->Grows, B, V, NRW, f->n_uniq, f->uniq_ids, f->uniq_off, f->uniq_rows, f->c2w,
                                                               f->ctx->Gv); nl++;
    LEMO_CUDA(cudaGetLastError());
    // ---------------- backward through the body model
    LEMO_TRY(body_grad_begin(f->ctx, B, st));
    LEMO_TRY(body_skin_backward(f->ctx, f->ctx, B, f->ctx->Gv, f->d_joints, st)); nl += 3;
    PoseGrad pg;
    pg.transl = f->g(O_TR); pg.global_orient = f->g(O_GO); pg.R_body = f->dRb; pg.jaw = f->g(O_JAW); pg.leye = f->g(O_LE); pg.reye = f->g(O_RE);
    pg.lhand = f->g(O_LH); pg.rhand = f->g(O_RH); pg.expression = f->g(O_EX);
    LEMO_TRY(body_pose_backward(f->ctx, in, B, pg, st)); nl += 3;
    k_prox_angle<<<1, 256, 0, st>>>(f->ctx->full_pose, f->Rb, B, f->dev, f->dRb, f->part); nl++;
    pc.n[PT_ANGLE] = 1;
    LEMO_TRY(vposer_decode_backward(f->vp, f->p(O_Z), B, f->dRb, f->g(O_Z), st)); nl += 4;
    k_prox_priors<<<1, 1024, 0, st>>>(f->P, f->Gp, f->betas, B, f->dev, f->part); nl++;
    pc.n[PT_PPRIOR] = pc.n[PT_SHAPE] = pc.n[PT_HAND] = pc.n[PT_EXPR] = pc.n[PT_JAW] = 1;
    k_prox_finish<<<1, 256, 0, st>>>(f->part, pc, f->dev, f->losses); nl++;
    k_prox_erase<<<cdiv(B * PP, 256), 256, 0, st>>>(f->Gp, B, f->dev); nl++;
    if (with_adam) { k_adam_dev<<<cdiv(B * PP, 256), 256, 0, st>>>(f->P, f->Gp, f->M1, f->M2, B * PP, f->sched); nl++; }
    LEMO_CUDA(cudaGetLastError());
    f->launches_per_iter = nl;
    return 0;
}

static void prox_drop_graph(ProxFit* f) {
    if (f->gexec) { cudaGraphExecDestroy(f->gexec); f->gexec = nullptr; }
    if (f->graph) { cudaGraphDestroy(f->graph); f->graph = nullptr; }
}

static int upload_ints(int** dst, const std::vector<int>& v) {
    LEMO_CUDA(cudaMalloc((void**)dst, std::max<size_t>(1, v.size()) * sizeof(int)));
    if (!v.empty()) LEMO_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

}  // namespace lemo

using namespace lemo;
#include "handles.cuh"
struct LemoProxFit { ProxFit f; };

extern "C" {

int lemo_fit_prox_create(const LemoModel* model, LemoVPoser* vposer, LemoConvNet* enc, const LemoProxConfigC* cfg, int device, LemoProxFit** out) {
    LEMO_CHECK(model && vposer && cfg && out, "null argument");
    LEMO_CHECK(cfg->n_frames >= 2, "a window needs at least two frames");
    LEMO_CHECK(!model->m->is_sub, "the PROX driver needs the full mesh (SDF penetration reads every vertex)");
    LEMO_CUDA(cudaSetDevice(device));
    LemoProxFit* h = new LemoProxFit();
    ProxFit* f = &h->f;
    const int B = cfg->n_frames, V = model->m->V;
    f->device = device; f->B = B; f->model = model->m; f->vp = vposer->v; f->enc = enc ? enc->n : nullptr;
    f->nout = NJ + model->m->n_extra + model->m->n_lmk;
    f->use_graph = cfg->use_cuda_graph;
    LEMO_CHECK(f->vp->maxB >= B, "VPoser handle batch too small for this window");
    // keypoint map (JointMapper of smplx_to_openpose, misc_utils.py) -> CSR inverse
    f->Jm = cfg->h_joint_map ? cfg->n_joints_mapped : f->nout;
    std::vector<std::vector<int>> inv(f->nout);
    for (int s = 0; s < f->Jm; ++s) {
        const int j = cfg->h_joint_map ? cfg->h_joint_map[s] : s;
        LEMO_CHECK(j >= 0 && j < f->nout, "joint map entry out of range");
        inv[j].push_back(s);
    }
    std::vector<int> ioff(1, 0), iidx;
    for (auto& l : inv) { for (int s : l) iidx.push_back(s); ioff.push_back((int)iidx.size()); }
    LEMO_TRY(upload_ints(&f->inv_off, ioff)); LEMO_TRY(upload_ints(&f->inv_idx, iidx));
    // cameras
    for (int i = 0; i < 9; ++i) { f->cam.R[i] = cfg->cam_R[i]; f->c2w.R[i] = cfg->R[i]; }
    for (int i = 0; i < 3; ++i) { f->cam.t[i] = cfg->cam_t[i]; f->c2w.t[i] = cfg->t[i]; }
    f->cam.fx = cfg->fx; f->cam.fy = cfg->fy; f->cam.cx = cfg->cx; f->cam.cy = cfg->cy;
    f->c2w.fx = f->c2w.fy = 1.f; f->c2w.cx = f->c2w.cy = 0.f;
    // scene
    f->has_sdf = cfg->sdf != nullptr && cfg->sdf_penetration;
    f->has_fric = cfg->sdf != nullptr && cfg->use_friction && cfg->n_fric > 0;
    if (cfg->sdf) {
        LEMO_CHECK(cfg->sdf_dim > 1, "sdf_dim");
        f->sdf = cfg->sdf; f->grid.dim = cfg->sdf_dim;
        for (int i = 0; i < 3; ++i) { f->grid.gmin[i] = cfg->grid_min[i]; f->grid.gmax[i] = cfg->grid_max[i]; }
    }
    f->has_contact = cfg->contact && cfg->scene_v && cfg->n_scene > 0 && cfg->n_contact > 0;
    f->has_smooth = cfg->use_motion_smooth_prior && f->enc != nullptr;
    f->n_fric = f->has_fric ? cfg->n_fric : 0;
    f->n_contact = f->has_contact ? cfg->n_contact : 0;
    // world rows [81 markers | friction | contact] and the distinct-vertex CSR used by the adjoint scatter
    std::vector<int> rows;
    LEMO_CHECK(cfg->h_markers81, "the 81 smoothness markers are required (loader/SSM2_withhand.json)");
    for (int i = 0; i < 81; ++i) rows.push_back(cfg->h_markers81[i]);
    f->off_fric = (int)rows.size();
    for (int i = 0; i < f->n_fric; ++i) rows.push_back(cfg->h_fric_ids[i]);
    f->off_contact = (int)rows.size();
    for (int i = 0; i < f->n_contact; ++i) rows.push_back(cfg->h_contact_ids[i]);
    f->NRW = (int)rows.size();
    for (int r : rows) LEMO_CHECK(r >= 0 && r < V, "vertex id out of range");
    std::vector<int> order(rows.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows[a] < rows[b]; });
    std::vector<int> uid, uoff(1, 0), urow;
    for (size_t i = 0; i < order.size(); ++i) {
        if (i == 0 || rows[order[i]] != rows[order[i - 1]]) { if (i) uoff.push_back((int)urow.size()); uid.push_back(rows[order[i]]); }
        urow.push_back(order[i]);
    }
    uoff.push_back((int)urow.size());
    f->n_uniq = (int)uid.size();
    LEMO_TRY(upload_ints(&f->row_ids, rows)); LEMO_TRY(upload_ints(&f->uniq_ids, uid)); LEMO_TRY(upload_ints(&f->uniq_off, uoff));
    LEMO_TRY(upload_ints(&f->uniq_rows, urow));
    if (f->has_contact) {
        f->scene = cfg->scene_v; f->n_scene = cfg->n_scene;
        LEMO_TRY(scene_grid_create(cfg->scene_v, cfg->n_scene, &f->sgrid));
    }
    LEMO_TRY(bodyctx_create(f->model, B, true, &f->ctx));
    const size_t Bz = B;
    LEMO_TRY(dalloc(&f->P, Bz * PP)); LEMO_TRY(dalloc(&f->Gp, Bz * PP)); LEMO_TRY(dalloc(&f->M1, Bz * PP)); LEMO_TRY(dalloc(&f->M2, Bz * PP));
    LEMO_TRY(dalloc(&f->betas, Bz * 10));
    LEMO_TRY(dalloc(&f->gt, Bz * f->Jm * 2)); LEMO_TRY(dalloc(&f->conf, Bz * f->Jm)); LEMO_TRY(dalloc(&f->jw, Bz * f->Jm));
    LEMO_TRY(dalloc(&f->Rb, Bz * NBODY * 9)); LEMO_TRY(dalloc(&f->dRb, Bz * NBODY * 9));
    LEMO_TRY(dalloc(&f->verts, Bz * V * 3)); LEMO_TRY(dalloc(&f->joints, Bz * f->nout * 3)); LEMO_TRY(dalloc(&f->d_joints, Bz * f->nout * 3));
    LEMO_TRY(dalloc(&f->Wrows, Bz * f->NRW * 3)); LEMO_TRY(dalloc(&f->Grows, Bz * f->NRW * 3));
    LEMO_TRY(dalloc(&f->sdf_fric, Bz * std::max(1, f->n_fric)));
    LEMO_TRY(dalloc(&f->cdist, Bz * std::max(1, f->n_contact))); LEMO_TRY(dalloc(&f->cidx, Bz * std::max(1, f->n_contact)));
    LEMO_TRY(dalloc(&f->part, (size_t)PT_N * PARTS)); LEMO_TRY(dalloc(&f->losses, PT_N)); LEMO_TRY(dalloc(&f->fpart, 4 * FR_G));
    LEMO_TRY(dalloc(&f->dev, 1)); LEMO_TRY(dalloc(&f->sched, 1));
    LEMO_TRY(dalloc(&f->canon, 12)); LEMO_TRY(dalloc(&f->stats, 486));
    if (f->has_smooth) {
        LEMO_CHECK(cfg->h_smooth_mean && cfg->h_smooth_std, "smoothness prior needs its normalisation statistics");
        LEMO_CHECK(B - 1 > 8, "the smoothness prior reflect-pads 8 frames: need n_frames >= 10");
        LEMO_CUDA(cudaMemcpy(f->stats, cfg->h_smooth_mean, 243 * sizeof(float), cudaMemcpyHostToDevice));
        LEMO_CUDA(cudaMemcpy(f->stats + 243, cfg->h_smooth_std, 243 * sizeof(float), cudaMemcpyHostToDevice));
        f->geom = f->enc->geom[0];
        LEMO_CHECK(f->geom.H == 245 && f->geom.W == B - 1 + 16, "Enc handle must be created for H=245, W=B-1+16");
        LEMO_CHECK(f->enc->maxN >= 1 && f->enc->with_backward, "Enc handle without backward");
        LEMO_TRY(dalloc(&f->xin, (size_t)f->geom.PS)); LEMO_TRY(dalloc(&f->gx, (size_t)f->geom.PS));
        LEMO_TRY(dalloc(&f->gv, (size_t)243 * (B - 1)));
    }
    *out = h;
    return lemo_fit_prox_set_weights(h, &cfg->weights, 0, nullptr);
}

int lemo_fit_prox_destroy(LemoProxFit* h) {
    if (!h) return 0;
    ProxFit* f = &h->f;
    cudaSetDevice(f->device);
    prox_drop_graph(f);
    if (f->gstream) { cudaStreamDestroy(f->gstream); cudaEventDestroy(f->ev_in); cudaEventDestroy(f->ev_out); }
    float* ps[] = {f->P, f->Gp, f->M1, f->M2, f->betas, f->gt, f->conf, f->jw, f->Rb, f->dRb, f->verts, f->joints, f->d_joints, f->Wrows, f->Grows,
                   f->sdf_fric, f->cdist, f->xin, f->gx, f->gv, f->canon, f->stats, f->part, f->losses, f->fpart};
    for (float* p : ps) cudaFree(p);
    int* is[] = {f->row_ids, f->uniq_ids, f->uniq_off, f->uniq_rows, f->inv_off, f->inv_idx, f->cidx};
    for (int* p : is) cudaFree(p);
    cudaFree(f->dev); cudaFree(f->sched);
    scene_grid_free(f->sgrid);
    bodyctx_free(f->ctx);
    delete h;
    return 0;
}

int lemo_fit_prox_set_weights(LemoProxFit* h, const LemoProxWeightsC* w, int32_t erase_n, void* stream) {
    LEMO_CHECK(h && w, "null argument");
    ProxFit* f = &h->f;
    LEMO_CHECK(erase_n >= 0 && erase_n <= f->B, "erase_n out of range");
    ProxDev d;
    d.w = *w; d.erase_n = erase_n;
    if (f->smooth_w_host != w->motion_prior_smooth_weight) prox_drop_graph(f);     // baked into the Enc loss kernel's arguments
    f->smooth_w_host = w->motion_prior_smooth_weight;
    LEMO_CUDA(cudaMemcpyAsync(f->dev, &d, sizeof(ProxDev), cudaMemcpyHostToDevice, (cudaStream_t)stream));   // pageable source: staged before return
    return 0;
}

static int copy_block(float* dst, const float* src, size_t n, cudaStream_t st) {
    if (src) LEMO_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else LEMO_CUDA(cudaMemsetAsync(dst, 0, n * sizeof(float), st));
    return 0;
}

int lemo_fit_prox_set_window(LemoProxFit* h, const LemoProxWindowC* w, void* stream) {
    LEMO_NVTX("lemo_fit_prox_set_window");
    LEMO_CHECK(h && w, "null argument");
    LEMO_CHECK(w->gt_joints && w->joint_weights, "gt_joints and joint_weights are required");
    ProxFit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = f->B;
    LEMO_TRY(copy_block(f->p(O_TR), w->transl, B * 3, st)); LEMO_TRY(copy_block(f->p(O_GO), w->global_orient, B * 3, st));
    LEMO_TRY(copy_block(f->p(O_Z), w->pose_embedding, B * 32, st));
    LEMO_TRY(copy_block(f->p(O_LH), w->left_hand_pose, B * 12, st)); LEMO_TRY(copy_block(f->p(O_RH), w->right_hand_pose, B * 12, st));
    LEMO_TRY(copy_block(f->p(O_JAW), w->jaw_pose, B * 3, st)); LEMO_TRY(copy_block(f->p(O_LE), w->leye_pose, B * 3, st));
    LEMO_TRY(copy_block(f->p(O_RE), w->reye_pose, B * 3, st)); LEMO_TRY(copy_block(f->p(O_EX), w->expression, B * 10, st));
    LEMO_TRY(copy_block(f->betas, w->betas, B * 10, st));
    LEMO_TRY(copy_block(f->gt, w->gt_joints, B * f->Jm * 2, st)); LEMO_TRY(copy_block(f->jw, w->joint_weights, B * f->Jm, st));
    if (w->joints_conf) LEMO_TRY(copy_block(f->conf, w->joints_conf, B * f->Jm, st));
    else {          // no confidences: all ones (use_joints_conf False gives the same arithmetic)
        std::vector<float> ones(B * f->Jm, 1.f);
        LEMO_CUDA(cudaMemcpyAsync(f->conf, ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice, st));
        LEMO_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

static int prox_begin(ProxFit* f, float lr, bool reset_moments, cudaStream_t st) {
    if (reset_moments) {
        LEMO_CUDA(cudaMemsetAsync(f->M1, 0, (size_t)f->B * PP * sizeof(float), st));
        LEMO_CUDA(cudaMemsetAsync(f->M2, 0, (size_t)f->B * PP * sizeof(float), st));
    }
    Sched s{};
    s.it = 0; s.lr0 = s.lr1 = s.lr2 = lr; s.sw1 = s.sw2 = 1 << 30;
    LEMO_CUDA(cudaMemcpyAsync(f->sched, &s, sizeof(Sched), cudaMemcpyHostToDevice, st));
    return 0;
}

int lemo_fit_prox_run(LemoProxFit* h, int32_t n_iters, float lr, int32_t resume, void* stream) {
    LEMO_NVTX("lemo_fit_prox_run");
    LEMO_CHECK(h && n_iters >= 0, "bad arguments");
    ProxFit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    // a fresh optimiser per stage (optim_factory.create_optimizer, fit_temp_loadprox_slide.py:519); resume keeps moments, step count and lr
    if (!resume) LEMO_TRY(prox_begin(f, lr, true, st));
    if (f->use_graph && n_iters > 0) {
        if (!f->gstream) {
            LEMO_CUDA(cudaStreamCreateWithFlags(&f->gstream, cudaStreamNonBlocking));
            LEMO_CUDA(cudaEventCreateWithFlags(&f->ev_in, cudaEventDisableTiming));
            LEMO_CUDA(cudaEventCreateWithFlags(&f->ev_out, cudaEventDisableTiming));
        }
        LEMO_CUDA(cudaEventRecord(f->ev_in, st));
        LEMO_CUDA(cudaStreamWaitEvent(f->gstream, f->ev_in, 0));
        if (!f->gexec) {
            LEMO_CUDA(cudaStreamBeginCapture(f->gstream, cudaStreamCaptureModeThreadLocal));
            const int r = prox_iteration(f, true, f->gstream);
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(f->gstream, &g);
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            LEMO_CUDA(e);
            f->graph = g;
            LEMO_CUDA(cudaGraphInstantiate(&f->gexec, g, 0));
        }
        for (int i = 0; i < n_iters; ++i) LEMO_CUDA(cudaGraphLaunch(f->gexec, f->gstream));
        LEMO_CUDA(cudaEventRecord(f->ev_out, f->gstream));
        LEMO_CUDA(cudaStreamWaitEvent(st, f->ev_out, 0));
    } else {
        for (int i = 0; i < n_iters; ++i) LEMO_TRY(prox_iteration(f, true, st));
    }
    f->launches += (long long)n_iters * f->launches_per_iter;
    return 0;
}

int lemo_fit_prox_eval(LemoProxFit* h, void* stream) {
    LEMO_NVTX("lemo_fit_prox_eval");
    LEMO_CHECK(h, "null handle");
    ProxFit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_TRY(prox_begin(f, 0.f, false, st));
    LEMO_TRY(prox_iteration(f, false, st));
    f->launches += f->launches_per_iter;
    return 0;
}

static int out_block(float* dst, const float* src, size_t n, cudaStream_t st) {
    if (dst) LEMO_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

static int prox_export(ProxFit* f, const float* base, const LemoProxParamsOutC* o, cudaStream_t st) {
    const size_t B = f->B;
    LEMO_TRY(out_block(o->transl, base + B * O_TR, B * 3, st)); LEMO_TRY(out_block(o->global_orient, base + B * O_GO, B * 3, st));
    LEMO_TRY(out_block(o->pose_embedding, base + B * O_Z, B * 32, st)); LEMO_TRY(out_block(o->left_hand_pose, base + B * O_LH, B * 12, st));
    LEMO_TRY(out_block(o->right_hand_pose, base + B * O_RH, B * 12, st)); LEMO_TRY(out_block(o->jaw_pose, base + B * O_JAW, B * 3, st));
    LEMO_TRY(out_block(o->leye_pose, base + B * O_LE, B * 3, st)); LEMO_TRY(out_block(o->reye_pose, base + B * O_RE, B * 3, st));
    LEMO_TRY(out_block(o->expression, base + B * O_EX, B * 10, st));
    return 0;
}

int lemo_fit_prox_get(LemoProxFit* h, const LemoProxParamsOutC* params, const LemoProxParamsOutC* grads, float* losses16, void* stream) {
    LEMO_CHECK(h, "null handle");
    ProxFit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    if (params) LEMO_TRY(prox_export(f, f->P, params, st));
    if (grads) LEMO_TRY(prox_export(f, f->Gp, grads, st));
    if (losses16) LEMO_CUDA(cudaMemcpyAsync(losses16, f->losses, PT_N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int64_t lemo_fit_prox_kernel_launches(const LemoProxFit* h) { return h ? h->f.launches : 0; }
}
