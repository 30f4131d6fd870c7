// Fused fitting drivers: the Adam inner loops of the reference's
//   opt_amass_temp.py:343-455      (temporal stage: B = T frames, marker L1 + smoothness prior +
//                                   foot-contact velocity + L2 priors)
//   opt_amass_perframe.py:293-361  (per-frame stage: T sequential B=1 problems, warm start)
// run entirely on the device for S independent sequences side by side: no host round trip inside an
// iteration (the reference has 4 .item() syncs per step), SMPL-X evaluated once instead of twice,
// only the 253 loss rows of the mesh are skinned (81 markers + 172 heel/toe vertices; identical loss,
// SURVEY.md section 0.10), LR schedule and Adam bias corrections computed on device so one iteration is a
// replayable CUDA graph.
#include "common.cuh"
#include "body.cuh"
#include "vposer.cuh"
#include "conv.cuh"
#include "gemm.cuh"
#include "fit_common.cuh"
#include "perframe_mega.cuh"
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include "../../include/lemo_b200.h"
#include <vector>

namespace lemo {

struct Fit {
    int device = 0, mode = 0, S = 0, T = 0, B = 0;   // B = rows in flight (S*T temporal, S per-frame)
    LemoFitConfigC cfg{};
    const Model* model = nullptr;
    Model* sub = nullptr;
    BodyCtx* ctx = nullptr;
    VPoser* vp = nullptr;
    ConvNet* enc = nullptr;
    int NR = 0, foot_off[4] = {0, 0, 0, 0}, foot_n[4] = {0, 0, 0, 0};
    PlaneGeom geom{};
    // optimisation state: P = [transl B*3 | rot6d B*6 | z B*32 | lh B*12 | rh B*12]
    float *P = nullptr, *Gp = nullptr, *M1 = nullptr, *M2 = nullptr;
    float *betas = nullptr;                  // [B,10]
    float *mrec = nullptr, *contact = nullptr;   // [S*T,67,3], [S*T,4]   (always full sequences)
    float *Rg = nullptr, *Rb = nullptr, *aa_body = nullptr, *dRg = nullptr, *dRb = nullptr;
    float *Vr = nullptr, *Grows = nullptr;   // [B,NR,3]
    float *xin = nullptr, *gx = nullptr;     // Enc input planes / their gradient [S][1][PS]
    float *gv = nullptr;                     // [S,243,T-1]
    float *canon = nullptr;                  // [S,12] Rt(9) + origin(3)
    float *stats = nullptr;                  // Xmean[243] Xstd[243]
    float *acc = nullptr;                    // [S,16] loss accumulators
    float *p72 = nullptr;                    // [S,T,72] snapshot of the last forward
    float *pf_ws = nullptr;                  // per-frame persistent kernel: per-CTA partials [S][8][512 + 512 + 664]
    Sched* sched = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaStream_t gstream = nullptr;          // graphs are captured/replayed on a private stream (the caller may be on the
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;   // legacy default stream, which cannot be captured) and fork/joined with events
    long long launches = 0, launches_per_iter = 0;
    float* tr() const { return P; }
    float* r6() const { return P + (size_t)B * 3; }
    float* zz() const { return P + (size_t)B * 9; }
    float* lh() const { return P + (size_t)B * 41; }
    float* rh() const { return P + (size_t)B * 53; }
    float* g_tr() const { return Gp; }
    float* g_r6() const { return Gp + (size_t)B * 3; }
    float* g_zz() const { return Gp + (size_t)B * 9; }
    float* g_lh() const { return Gp + (size_t)B * 41; }
    float* g_rh() const { return Gp + (size_t)B * 53; }
};

enum { ACC_REC = 0, ACC_VP, ACC_SHAPE, ACC_HAND, ACC_SMOOTH, ACC_CNT0, ACC_SUM0 = ACC_CNT0 + 4, ACC_N = 16 };


// marker reconstruction loss  F.l1_loss(markers_opt, markers_rec)  (opt_amass_temp.py:395)
__global__ void __launch_bounds__(256) k_marker_l1(const float* __restrict__ Vr, const float* __restrict__ mrec, int Tb, int NR, float w,
                                                   float* __restrict__ Grows, float* __restrict__ acc) {
    __shared__ float sred[32];
    const int s = blockIdx.y;
    const int n = Tb * 67 * 3;
    float part = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % 3, mk = (i / 3) % 67, t = i / 201;
        const size_t b = (size_t)s * Tb + t;
        const float d = Vr[(b * NR + mk) * 3 + k] - mrec[(b * 67 + mk) * 3 + k];
        part += fabsf(d);
        const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        Grows[(b * NR + mk) * 3 + k] += w * sg / (float)n;
    }
    part = block_sum(part, sred);
    if (threadIdx.x == 0) atomicAdd(&acc[s * ACC_N + ACC_REC], part / (float)n);
}

// L2 priors (opt_amass_temp.py:397-404): mean(z^2), mean(betas^2), mean(hand^2); adds their grads
__global__ void __launch_bounds__(1024) k_priors(const float* __restrict__ z, const float* __restrict__ lh, const float* __restrict__ rh,
                                                const float* __restrict__ betas, int Tb, float w_vp, float w_hand,
                                                float* __restrict__ gz, float* __restrict__ glh, float* __restrict__ grh,
                                                float* __restrict__ acc) {
    __shared__ float sred[32];
    const int s = blockIdx.x;
    float pv = 0.f, ph = 0.f, ps = 0.f;
    const int nz = Tb * 32, nh = Tb * 12, nb = Tb * 10;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) {
        const size_t j = (size_t)s * nz + i;
        const float x = z[j]; pv += x * x; gz[j] += w_vp * 2.f * x / (float)nz;
    }
    for (int i = threadIdx.x; i < nh; i += blockDim.x) {
        const size_t j = (size_t)s * nh + i;
        const float a = lh[j], b = rh[j]; ph += a * a + b * b;
        glh[j] += w_hand * 2.f * a / (float)(2 * nh); grh[j] += w_hand * 2.f * b / (float)(2 * nh);
    }
    for (int i = threadIdx.x; i < nb; i += blockDim.x) { const float x = betas[(size_t)s * nb + i]; ps += x * x; }
    pv = block_sum(pv, sred); if (threadIdx.x == 0) acc[s * ACC_N + ACC_VP] = pv / (float)nz;
    ph = block_sum(ph, sred); if (threadIdx.x == 0) acc[s * ACC_N + ACC_HAND] = ph / (float)(2 * nh);
    ps = block_sum(ps, sred); if (threadIdx.x == 0) acc[s * ACC_N + ACC_SHAPE] = ps / (float)nb;
}

// foot-contact velocity loss (opt_amass_temp.py:407-447).  grid (4 parts, S): one CTA per (part, sequence) does the
// count+sum pass, reduces in-block, then the gradient pass -- the data-dependent masked mean with its empty-set guard
// (`if (...).sum().item() >= 1`, four host syncs per iteration in the reference) never leaves the device.
struct FootTab { int off[4]; int n[4]; };
__global__ void __launch_bounds__(1024) k_contact(const float* __restrict__ Vr, const float* __restrict__ contact, int T, int NR, FootTab ft,
                                                 float fps, float thres, float w, float* __restrict__ acc, float* __restrict__ Grows) {
    __shared__ float sred[32];
    __shared__ float s_cnt;
    const int part = blockIdx.x, s = blockIdx.y;
    const int off = ft.off[part], cnt_rows = ft.n[part];
    const int n = (T - 1) * cnt_rows;
    float cnt = 0.f, sum = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i % cnt_rows, t = i / cnt_rows;
        const size_t b = (size_t)s * T + t;
        if (contact[b * 4 + part] != 1.f) continue;
        const float* a0 = Vr + (b * NR + off + r) * 3;
        const float* a1 = Vr + ((b + 1) * NR + off + r) * 3;
        const float vx = (a1[0] - a0[0]) * fps, vy = (a1[1] - a0[1]) * fps, vz = (a1[2] - a0[2]) * fps;
        const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
        if (nrm > thres) { cnt += 1.f; sum += nrm; }
    }
    cnt = block_sum(cnt, sred);
    if (threadIdx.x == 0) { s_cnt = cnt; acc[s * ACC_N + ACC_CNT0 + part] = cnt; }
    sum = block_sum(sum, sred);
    if (threadIdx.x == 0) acc[s * ACC_N + ACC_SUM0 + part] = sum;
    __syncthreads();
    if (s_cnt < 1.f) return;
    const float inv = w * fps / s_cnt;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = i % cnt_rows, t = i / cnt_rows;
        const size_t b = (size_t)s * T + t;
        if (contact[b * 4 + part] != 1.f) continue;
        const float* a0 = Vr + (b * NR + off + r) * 3;
        const float* a1 = Vr + ((b + 1) * NR + off + r) * 3;
        const float vx = (a1[0] - a0[0]) * fps, vy = (a1[1] - a0[1]) * fps, vz = (a1[2] - a0[2]) * fps;
        const float nrm = sqrtf(vx * vx + vy * vy + vz * vz);
        if (nrm > thres) {
            const float c = inv / nrm;
            float* g0 = Grows + (b * NR + off + r) * 3;
            float* g1 = Grows + ((b + 1) * NR + off + r) * 3;
            atomicAdd(&g1[0], c * vx); atomicAdd(&g1[1], c * vy); atomicAdd(&g1[2], c * vz);
            atomicAdd(&g0[0], -c * vx); atomicAdd(&g0[1], -c * vy); atomicAdd(&g0[2], -c * vz);
        }
    }
}

// canonical frame of the smoothness prior (opt_amass_temp.py:368-377), detached: Rt (3x3) + origin (marker 0 of frame 0)
__global__ void k_canon(const float* __restrict__ Jposed, const float* __restrict__ Vr, int T, int NR, int S, float* __restrict__ canon) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float* J = Jposed + (size_t)s * T * NJ * 3;
    float x0 = J[2 * 3] - J[1 * 3], x1 = J[2 * 3 + 1] - J[1 * 3 + 1];
    const float nx = sqrtf(x0 * x0 + x1 * x1);
    x0 /= nx; x1 /= nx;
    float y0 = -x1, y1 = x0;                       // cross((0,0,1), x)
    const float ny = sqrtf(y0 * y0 + y1 * y1);
    y0 /= ny; y1 /= ny;
    float* c = canon + s * 12;
    c[0] = x0; c[1] = y0; c[2] = 0.f;              // Rt[k][c] row-major, columns = x,y,z axes
    c[3] = x1; c[4] = y1; c[5] = 0.f;
    c[6] = 0.f; c[7] = 0.f; c[8] = 1.f;
    const float* o = Vr + (size_t)s * T * NR * 3;
    c[9] = o[0]; c[10] = o[1]; c[11] = o[2];
}

// [B,72] result vector of the current forward (transl, aa(global), betas, z, lh, rh)
__global__ void k_snapshot(const float* __restrict__ tr, const float* __restrict__ full_pose, const float* __restrict__ betas,
                           const float* __restrict__ z, const float* __restrict__ lh, const float* __restrict__ rh, int B,
                           float* __restrict__ out, int out_row_stride, const Sched* __restrict__ sc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 72) return;
    const int out_row0 = out_row_stride > 1 ? sc->frame : 0;
    const int b = i / 72, c = i - b * 72;
    float v;
    if (c < 3) v = tr[b * 3 + c];
    else if (c < 6) v = full_pose[b * 165 + (c - 3)];
    else if (c < 16) v = betas[b * 10 + (c - 6)];
    else if (c < 48) v = z[b * 32 + (c - 16)];
    else if (c < 60) v = lh[b * 12 + (c - 48)];
    else v = rh[b * 12 + (c - 60)];
    out[((size_t)b * out_row_stride + out_row0) * 72 + c] = v;
}

__global__ void k_split72(const float* __restrict__ p72, int B, float* tr, float* r6, float* betas, float* z, float* lh, float* rh) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* p = p72 + (size_t)b * 72;
    for (int k = 0; k < 3; ++k) tr[b * 3 + k] = p[k];
    float aa[3] = {p[3], p[4], p[5]}, r[9];
    aa_to_rotmat_tgm(aa, r);                              // convert_to_6D_all (utils/utils.py:127-130)
    r6[b * 6 + 0] = r[0]; r6[b * 6 + 1] = r[1]; r6[b * 6 + 2] = r[3];
    r6[b * 6 + 3] = r[4]; r6[b * 6 + 4] = r[6]; r6[b * 6 + 5] = r[7];
    for (int k = 0; k < 10; ++k) betas[b * 10 + k] = p[6 + k];
    for (int k = 0; k < 32; ++k) z[b * 32 + k] = p[16 + k];
    for (int k = 0; k < 12; ++k) { lh[b * 12 + k] = p[48 + k]; rh[b * 12 + k] = p[60 + k]; }
}

__global__ void k_losses_out(const float* __restrict__ acc, const LemoFitConfigC cfg, int S, float* __restrict__ out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float* a = acc + s * ACC_N;
    float con = 0.f;
    for (int p = 0; p < 4; ++p) if (a[ACC_CNT0 + p] >= 1.f) con += a[ACC_SUM0 + p] / a[ACC_CNT0 + p];
    float* o = out + s * 8;
    o[1] = a[ACC_REC]; o[2] = a[ACC_VP]; o[3] = a[ACC_SHAPE]; o[4] = a[ACC_HAND]; o[5] = con; o[6] = a[ACC_SMOOTH]; o[7] = 0.f;
    o[0] = cfg.w_rec * o[1] + cfg.w_vposer * o[2] + cfg.w_shape * o[3] + cfg.w_hand * o[4] + cfg.w_contact * o[5] + cfg.w_smooth * o[6];
}

static int fit_iteration(Fit* f, cudaStream_t st) {
    const int B = f->B, S = f->S, NR = f->NR;
    const int Tb = f->mode == 0 ? f->T : 1;         // frames per sequence inside the batch
    const LemoFitConfigC& c = f->cfg;
    long long nl = 0;
    k_sched<<<1, 1, 0, st>>>(f->sched); nl++;
    LEMO_CUDA(cudaMemsetAsync(f->Grows, 0, (size_t)B * NR * 3 * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(f->acc, 0, (size_t)S * ACC_N * sizeof(float), st));
    // ---------------- forward
    LEMO_TRY(lemo_rot6d_to_rotmat(f->r6(), B, f->Rg, st)); nl++;
    LEMO_TRY(vposer_decode(f->vp, f->zz(), B, f->Rb, nullptr, st)); nl += 4;
    PoseIn in;
    in.transl = f->tr(); in.R_global = f->Rg; in.R_body = f->Rb; in.lhand = f->lh(); in.rhand = f->rh();
    in.betas = f->betas; in.betas_stride = 10; in.hand_is_pca = 1;
    LEMO_TRY(body_pose_forward(f->ctx, in, B, st)); nl += 1;
    LEMO_TRY(body_skin_forward(f->ctx, f->ctx, in, B, f->Vr, nullptr, st)); nl += 2;
    // snapshot of the parameters this forward used (what the scripts save after the loop)
    const float* contact = f->contact;
    k_snapshot<<<cdiv(B * 72, 256), 256, 0, st>>>(f->tr(), f->ctx->full_pose, f->betas, f->zz(), f->lh(), f->rh(), B, f->p72,
                                                   f->mode == 0 ? 1 : f->T, f->sched); nl++;
    // ---------------- losses + their gradients on the loss rows
    // (per-frame mode: the current frame's targets of every sequence are staged in f->gv by lemo_fit_run_perframe)
    k_marker_l1<<<dim3(cdiv(Tb * 201, 256), S), 256, 0, st>>>(f->Vr, f->mode == 0 ? f->mrec : f->gv, Tb, NR, c.w_rec, f->Grows, f->acc); nl++;
    const bool smooth = f->mode == 0 && f->enc && c.w_smooth > 0.f;
    const bool con = f->mode == 0 && c.w_contact > 0.f;
    if (con) {
        FootTab ft;
        for (int p = 0; p < 4; ++p) { ft.off[p] = f->foot_off[p]; ft.n[p] = f->foot_n[p]; }
        k_contact<<<dim3(4, S), 1024, 0, st>>>(f->Vr, contact, f->T, NR, ft, c.fps, c.vel_thres, c.w_contact, f->acc, f->Grows); nl++;
    }
    if (smooth) {
        const PlaneGeom& g = f->geom;
        k_canon<<<cdiv(S, 32), 32, 0, st>>>(f->ctx->Jposed, f->Vr, f->T, NR, S, f->canon); nl++;
        k_smooth_input<<<dim3(cdiv(g.W, 128), g.H, S), 128, 0, st>>>(f->Vr, f->canon, f->stats, f->T, NR, g.H, g.W, g.Wp, g.PS, f->xin); nl++;
        LEMO_TRY(enc_forward_planes(f->enc, f->xin, S, st)); nl += 10;
        if (enc_uses_tc(f->enc)) LEMO_TRY(enc_tc_smooth_loss(f->enc, S, c.w_smooth, ACC_N, ACC_SMOOTH, f->acc, st));
        else k_smooth_loss<<<dim3(8, 64, S), 256, 0, st>>>(enc_z_planes(f->enc), 64, g.H, g.W, g.Wp, g.PS, c.w_smooth, enc_gz_planes(f->enc), f->acc, ACC_N, ACC_SMOOTH);
        nl++;
        LEMO_TRY(enc_backward_planes(f->enc, S, f->gx, st)); nl += 10;
        k_smooth_bwd_a<<<dim3(cdiv(f->T - 1, 128), 243, S), 128, 0, st>>>(f->gx, f->T, g.H, g.W, g.Wp, g.PS, f->gv); nl++;
        k_smooth_bwd_b<<<cdiv(S * f->T * 81, 256), 256, 0, st>>>(f->gv, f->canon, f->stats, f->T, NR, S, f->Grows); nl++;
    }
    LEMO_CUDA(cudaGetLastError());
    // ---------------- backward through the body model
    LEMO_TRY(body_grad_begin(f->ctx, B, st));
    LEMO_TRY(body_skin_backward(f->ctx, f->ctx, B, f->Grows, nullptr, st)); nl += 2;
    PoseGrad pg;
    pg.transl = f->g_tr(); pg.R_global = f->dRg; pg.R_body = f->dRb; pg.lhand = f->g_lh(); pg.rhand = f->g_rh();
    LEMO_TRY(body_pose_backward(f->ctx, in, B, pg, st)); nl += 3;
    LEMO_TRY(lemo_rot6d_to_rotmat_backward(f->r6(), f->dRg, B, f->g_r6(), st)); nl++;
    LEMO_TRY(vposer_decode_backward(f->vp, f->zz(), B, f->dRb, f->g_zz(), st)); nl += 4;
    k_priors<<<S, 1024, 0, st>>>(f->zz(), f->lh(), f->rh(), f->betas, Tb, c.w_vposer, c.w_hand, f->g_zz(), f->g_lh(), f->g_rh(), f->acc); nl++;
    // ---------------- Adam
    k_adam_dev<<<cdiv(B * 65, 256), 256, 0, st>>>(f->P, f->Gp, f->M1, f->M2, B * 65, f->sched); nl++;
    LEMO_CUDA(cudaGetLastError());
    f->launches_per_iter = nl;
    return 0;
}

static int fit_run_iters(Fit* f, int n_iters, cudaStream_t st) {
    if (f->cfg.use_cuda_graph && n_iters > 0) {
        if (!f->gstream) {
            LEMO_CUDA(cudaStreamCreateWithFlags(&f->gstream, cudaStreamNonBlocking));
            LEMO_CUDA(cudaEventCreateWithFlags(&f->ev_in, cudaEventDisableTiming));
            LEMO_CUDA(cudaEventCreateWithFlags(&f->ev_out, cudaEventDisableTiming));
        }
        LEMO_CUDA(cudaEventRecord(f->ev_in, st));
        LEMO_CUDA(cudaStreamWaitEvent(f->gstream, f->ev_in, 0));
        if (!f->gexec) {
            LEMO_CUDA(cudaStreamBeginCapture(f->gstream, cudaStreamCaptureModeThreadLocal));
            const int r = fit_iteration(f, f->gstream);
            cudaGraph_t g = nullptr;
            const cudaError_t e = cudaStreamEndCapture(f->gstream, &g);     // always close the capture, even on error
            if (r) { if (g) cudaGraphDestroy(g); return r; }
            LEMO_CUDA(e);
            f->graph = g;
            LEMO_CUDA(cudaGraphInstantiate(&f->gexec, g, 0));
        }
        for (int i = 0; i < n_iters; ++i) LEMO_CUDA(cudaGraphLaunch(f->gexec, f->gstream));
        LEMO_CUDA(cudaEventRecord(f->ev_out, f->gstream));
        LEMO_CUDA(cudaStreamWaitEvent(st, f->ev_out, 0));
    } else {
        for (int i = 0; i < n_iters; ++i) LEMO_TRY(fit_iteration(f, st));
    }
    f->launches += (long long)n_iters * f->launches_per_iter;
    return 0;
}

}  // namespace lemo

using namespace lemo;
#include "handles.cuh"
struct LemoFit { Fit f; };

extern "C" {

int lemo_fit_create(const LemoModel* model, LemoVPoser* vposer, const float* unused, LemoConvNet* enc, const LemoFitConfigC* cfg,
                    int device, LemoFit** out) {
    (void)unused;
    LEMO_CHECK(model && vposer && cfg && out, "null argument");
    LEMO_CHECK(cfg->n_seq > 0 && cfg->n_frames > 1, "need at least one sequence of two frames");
    LEMO_CHECK(cfg->mode == 0 || cfg->mode == 1, "mode must be 0 (temporal) or 1 (per-frame)");
    LEMO_CUDA(cudaSetDevice(device));
    LemoFit* h = new LemoFit();
    Fit* f = &h->f;
    f->device = device; f->mode = cfg->mode; f->S = cfg->n_seq; f->T = cfg->n_frames;
    f->B = cfg->mode == 0 ? f->S * f->T : f->S;
    f->cfg = *cfg;
    f->model = model->m; f->vp = vposer->v; f->enc = enc ? enc->n : nullptr;
    LEMO_CHECK(f->vp->maxB >= f->B, "VPoser handle batch too small for this fit");
    // loss rows: 81 markers (first 67 == SSM2, opt_amass_temp.py:237-241) then the four foot sets
    std::vector<int> rows(cfg->h_markers81, cfg->h_markers81 + 81);
    for (int i = 0; i < 67; ++i) LEMO_CHECK(cfg->h_markers67[i] == cfg->h_markers81[i], "markers81 must start with markers67");
    for (int p = 0; p < 4; ++p) {
        f->foot_off[p] = (int)rows.size(); f->foot_n[p] = cfg->mode == 0 ? cfg->n_foot[p] : 0;
        for (int i = 0; i < f->foot_n[p]; ++i) rows.push_back(cfg->h_foot_ids[p][i]);
    }
    f->NR = (int)rows.size();
    LEMO_TRY(model_select_rows(f->model, rows.data(), f->NR, &f->sub));
    LEMO_TRY(bodyctx_create(f->sub, f->B, true, &f->ctx));
    const size_t B = f->B, ST = (size_t)f->S * f->T;
    LEMO_TRY(dalloc(&f->P, B * 65)); LEMO_TRY(dalloc(&f->Gp, B * 65)); LEMO_TRY(dalloc(&f->M1, B * 65)); LEMO_TRY(dalloc(&f->M2, B * 65));
    LEMO_TRY(dalloc(&f->betas, B * 10));
    LEMO_TRY(dalloc(&f->mrec, ST * 67 * 3)); LEMO_TRY(dalloc(&f->contact, ST * 4));
    LEMO_TRY(dalloc(&f->Rg, B * 9)); LEMO_TRY(dalloc(&f->Rb, B * NBODY * 9)); LEMO_TRY(dalloc(&f->dRg, B * 9)); LEMO_TRY(dalloc(&f->dRb, B * NBODY * 9));
    LEMO_TRY(dalloc(&f->Vr, B * f->NR * 3)); LEMO_TRY(dalloc(&f->Grows, B * f->NR * 3));
    LEMO_TRY(dalloc(&f->acc, (size_t)f->S * ACC_N));
    LEMO_TRY(dalloc(&f->p72, ST * 72));
    LEMO_TRY(dalloc(&f->sched, 1));
    LEMO_TRY(dalloc(&f->canon, (size_t)f->S * 12));
    LEMO_TRY(dalloc(&f->stats, 486));
    if (cfg->h_smooth_mean && cfg->h_smooth_std) {
        LEMO_CUDA(cudaMemcpy(f->stats, cfg->h_smooth_mean, 243 * sizeof(float), cudaMemcpyHostToDevice));
        LEMO_CUDA(cudaMemcpy(f->stats + 243, cfg->h_smooth_std, 243 * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (cfg->mode == 0 && f->enc) {
        // F.pad(..., (8,8,1,1), 'reflect') needs more than 8 velocity frames (torch raises otherwise, opt_amass_temp.py:385-387)
        LEMO_CHECK(f->T - 1 > 8, "the smoothness prior reflect-pads 8 frames: need n_frames >= 10");
        f->geom = f->enc->geom[0];
        LEMO_CHECK(f->geom.H == 245 && f->geom.W == f->T - 1 + 16, "Enc handle must be created for H=245, W=T-1+16");
        LEMO_CHECK(f->enc->maxN >= f->S && f->enc->with_backward, "Enc handle too small / without backward");
        LEMO_TRY(dalloc(&f->xin, (size_t)f->S * f->geom.PS)); LEMO_TRY(dalloc(&f->gx, (size_t)f->S * f->geom.PS));
        LEMO_TRY(dalloc(&f->gv, (size_t)f->S * 243 * (f->T - 1)));
    } else {
        LEMO_TRY(dalloc(&f->gv, (size_t)f->S * 67 * 3));      // per-frame: current-frame targets
    }
    *out = h;
    return 0;
}

int lemo_fit_destroy(LemoFit* h) {
    if (!h) return 0;
    Fit* f = &h->f;
    cudaSetDevice(f->device);
    if (f->gexec) cudaGraphExecDestroy(f->gexec);
    if (f->graph) cudaGraphDestroy(f->graph);
    if (f->gstream) { cudaStreamDestroy(f->gstream); cudaEventDestroy(f->ev_in); cudaEventDestroy(f->ev_out); }
    float* ps[] = {f->P, f->Gp, f->M1, f->M2, f->betas, f->mrec, f->contact, f->Rg, f->Rb, f->dRg, f->dRb, f->Vr, f->Grows, f->xin, f->gx,
                   f->gv, f->canon, f->stats, f->acc, f->p72, f->pf_ws};
    for (float* p : ps) cudaFree(p);
    cudaFree(f->sched);
    bodyctx_free(f->ctx);
    model_free(f->sub);
    delete h;
    return 0;
}

__global__ void k_init_perframe(const float* __restrict__ init72, float* tr, float* r6, float* betas, float* z, float* lh, float* rh) {
    const int k = threadIdx.x;
    if (k == 0) {
        tr[0] = 0.f; tr[1] = 0.4f; tr[2] = 1.0f;
        const float aa[3] = {0.f, 1.6f, 3.14f};
        float r[9];
        aa_to_rotmat_tgm(aa, r);
        r6[0] = r[0]; r6[1] = r[1]; r6[2] = r[3]; r6[3] = r[4]; r6[4] = r[6]; r6[5] = r[7];
    }
    if (k < 10) betas[k] = init72 ? init72[6 + k] : 0.f;
    if (k < 32) z[k] = 0.f;
    if (k < 12) { lh[k] = 0.f; rh[k] = 0.f; }
}

__global__ void k_copy_rows(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols, int src_stride_rows, int src_row0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i - r * cols;
    dst[i] = src[((size_t)r * src_stride_rows + src_row0) * cols + c];
}

int lemo_fit_set_sequence(LemoFit* h, int32_t s, const float* init72, const float* markers_rec, const float* contact, void* stream) {
    LEMO_CHECK(h && s >= 0 && s < h->f.S && markers_rec, "bad arguments");
    Fit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    const int T = f->T;
    LEMO_CUDA(cudaMemcpyAsync(f->mrec + (size_t)s * T * 201, markers_rec, (size_t)T * 201 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (contact) LEMO_CUDA(cudaMemcpyAsync(f->contact + (size_t)s * T * 4, contact, (size_t)T * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (f->mode == 0) {
        LEMO_CHECK(init72, "temporal mode needs the per-frame initialisation [T,72]");
        const size_t o = (size_t)s * T;
        k_split72<<<cdiv(T, 64), 64, 0, st>>>(init72, T, f->tr() + o * 3, f->r6() + o * 6, f->betas + o * 10, f->zz() + o * 32,
                                              f->lh() + o * 12, f->rh() + o * 12);
    } else {
        // per-frame: only betas (columns 6:16 of row 0) are taken from init72 (opt_amass_perframe.py:295);
        // the pose starts from transl (0,.4,1), aa (0,1.6,3.14), zeros (opt_amass_perframe.py:298-312)
        k_init_perframe<<<1, 64, 0, st>>>(init72, f->tr() + s * 3, f->r6() + s * 6, f->betas + s * 10, f->zz() + s * 32, f->lh() + s * 12,
                                          f->rh() + s * 12);
    }
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int lemo_fit_set_sequences(LemoFit* h, const float* init72, const float* markers_rec, const float* contact, void* stream) {
    LEMO_CHECK(h && init72 && markers_rec && contact, "bad arguments");
    Fit* f = &h->f;
    LEMO_CHECK(f->mode == 0, "lemo_fit_set_sequences is the temporal-mode loader");
    cudaStream_t st = (cudaStream_t)stream;
    const int B = f->S * f->T;                       // sequence s owns rows [s T, (s+1) T) of every per-frame array
    LEMO_CUDA(cudaMemcpyAsync(f->mrec, markers_rec, (size_t)B * 201 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    LEMO_CUDA(cudaMemcpyAsync(f->contact, contact, (size_t)B * 4 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    k_split72<<<cdiv(B, 64), 64, 0, st>>>(init72, B, f->tr(), f->r6(), f->betas, f->zz(), f->lh(), f->rh());
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

static int fit_begin_run(Fit* f, float lr0, float lr1, float lr2, int sw1, int sw2, int frame, cudaStream_t st) {
    // fresh optimiser (optim.Adam(final_params, lr=init_lr): opt_amass_temp.py:344-345, opt_amass_perframe.py:319)
    LEMO_CUDA(cudaMemsetAsync(f->M1, 0, (size_t)f->B * 65 * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(f->M2, 0, (size_t)f->B * 65 * sizeof(float), st));
    Sched h{};
    h.it = 0; h.lr0 = lr0; h.lr1 = lr1; h.lr2 = lr2; h.sw1 = sw1; h.sw2 = sw2; h.frame = frame;
    LEMO_CUDA(cudaMemcpyAsync(f->sched, &h, sizeof(Sched), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    return 0;
}

int lemo_fit_run(LemoFit* h, int32_t n_iters, float lr0, float lr1, int32_t lr_switch, void* stream) {
    LEMO_NVTX("lemo_fit_run");
    LEMO_CHECK(h && h->f.mode == 0 && n_iters >= 0, "lemo_fit_run is the temporal-mode driver");
    Fit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    LEMO_TRY(fit_begin_run(f, lr0, lr1, lr1, lr_switch, 1 << 30, 0, st));
    return fit_run_iters(f, n_iters, st);
}

// LEMO_PERFRAME=graph selects the round-1 path (one CUDA graph of ~23 kernels per step); default = the persistent cluster kernel
static int g_perframe_mode = -1;
static bool perframe_mega_enabled() {
    if (g_perframe_mode < 0) { const char* e = getenv("LEMO_PERFRAME"); g_perframe_mode = (e && strcmp(e, "graph") == 0) ? 0 : 1; }
    return g_perframe_mode != 0;
}

static int perframe_mega_run(Fit* f, int n_iters, cudaStream_t st) {
    const int S = f->S;
    if (!f->pf_ws) LEMO_TRY(dalloc(&f->pf_ws, (size_t)S * PM_CL * (512 + 512 + PM_PART)));
    LEMO_CHECK(f->sub->V <= PM_CL * PM_VPC, "per-frame kernel: more loss rows than the cluster covers");
    MegaArgs a{};
    const Model* m = f->sub;
    a.vt = m->v_template; a.Wt = m->Wt; a.wjm = m->w_jm; a.Jt = m->J_template; a.Jd = m->J_dirs; a.hand_l = m->hand_l; a.hand_r = m->hand_r;
    a.pose_mean = m->pose_mean; a.tree = m->tree; a.max_depth = m->max_depth; a.V = m->V; a.npc = m->npc;
    VPoser* v = f->vp;
    a.W1 = v->W1; a.b1 = v->b1; a.W2 = v->W2; a.b2 = v->b2; a.W3 = v->W3; a.b3 = v->b3;
    a.h2 = v->h2; a.o = v->o;
    a.dh1p = f->pf_ws; a.dXp = a.dh1p + (size_t)S * PM_CL * 512; a.dAp = a.dXp + (size_t)S * PM_CL * 512;
    a.P = f->P; a.Gp = f->Gp; a.betas = f->betas; a.mrec = f->mrec; a.p72 = f->p72; a.acc = f->acc;
    a.acc_n = ACC_N; a.acc_rec = ACC_REC; a.acc_vp = ACC_VP; a.acc_shape = ACC_SHAPE; a.acc_hand = ACC_HAND;
    a.S = S; a.T = f->T; a.n_iters = n_iters;
    a.w_rec = f->cfg.w_rec; a.w_vp = f->cfg.w_vposer; a.w_shape = f->cfg.w_shape; a.w_hand = f->cfg.w_hand;
    LEMO_CUDA(cudaMemsetAsync(f->acc, 0, (size_t)S * ACC_N * sizeof(float), st));
    static unsigned long long* d_tl = nullptr;           // LEMO_PERFRAME_TL=1: phase timeline of one step (debug; printed after a sync)
    static int want_tl = -1;
    if (want_tl < 0) { const char* e = getenv("LEMO_PERFRAME_TL"); want_tl = (e && e[0] == '1') ? 1 : 0; }
    if (want_tl && !d_tl) { LEMO_CUDA(cudaMalloc((void**)&d_tl, 32 * sizeof(unsigned long long))); LEMO_CUDA(cudaMemset(d_tl, 0, 32 * 8)); }
    a.tl = want_tl ? d_tl : nullptr;
    const int nclusters = std::min(S, 16);           // 8-CTA clusters: two per GPC; more sequences than that are walked in turn
    static bool configured = false;
    if (!configured) {
        LEMO_CUDA(cudaFuncSetAttribute(k_perframe_mega, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PM_DYN_BYTES));
        configured = true;
    }
    k_perframe_mega<<<PM_CL * nclusters, PM_NT, PM_DYN_BYTES, st>>>(a);
    LEMO_CUDA(cudaGetLastError());
    if (want_tl && n_iters > 5) {
        unsigned long long h[32];
        LEMO_CUDA(cudaStreamSynchronize(st));
        LEMO_CUDA(cudaMemcpy(h, d_tl, sizeof(h), cudaMemcpyDeviceToHost));
        const char* names[19] = {"P1 fc1 (all rows)", "cta barrier", "P2 fc2", "sync2", "P3 out", "sync3", "P4 gs + pose/chain fwd", "P5 blend/skin/loss",
                                 "P6 adjoint partials", "sync4", "P7 combine", "chain_bwd", "pose_to_rot_bwd + gs_bwd", "P8 dh2", "cta barrier", "P9 dh1 partial",
                                 "sync6", "P10 dh1/dz/priors", "adam"};
        printf("[perframe timeline] step 5 of frame 0, cluster 0 rank 0 (ns):");
        for (int i = 0; i < 18; ++i) printf(" %s %llu |", names[i], h[i + 1] - h[i]);
        printf(" total to adam %llu\n", h[18] - h[0]);
#ifdef LEMO_BODY_TL
        unsigned long long bt[32];
        LEMO_CUDA(cudaMemcpyFromSymbol(bt, g_body_tl, sizeof(bt)));
        const char* bn[17] = {"fwd: betas/rodrigues", "sync", "X + rest joints", "sync", "tree walk", "sync", "G/A/Jposed out", "(gap)", "bwd: load + dG init",
                              "sync", "tree walk", "sync", "dR out", "(gap)", "p2r: rodrigues adjoint", "sync", "hand PCA"};
        printf("[body timeline, last step] (ns):");
        for (int i = 0; i < 17; ++i) if (i != 7 && i != 13) printf(" %s %lld |", bn[i], (long long)(bt[i + 1] - bt[i]));
        printf("\n");
#endif
    }
    f->launches += 1;
    return 0;
}

int lemo_debug_set_perframe(int32_t mode) { g_perframe_mode = mode; return 0; }

int lemo_fit_run_perframe(LemoFit* h, int32_t n_iters, void* stream) {
    LEMO_NVTX("lemo_fit_run_perframe");
    LEMO_CHECK(h && h->f.mode == 1 && n_iters >= 0, "lemo_fit_run_perframe is the per-frame-mode driver");
    Fit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    if (perframe_mega_enabled() && n_iters > 0) return perframe_mega_run(f, n_iters, st);
    for (int t = 0; t < f->T; ++t) {
        // lr .1 for frame 0 else .01; ->.01 @step>60, ->.003 @step>80 (opt_amass_perframe.py:315-330); warm start = P carried over
        LEMO_TRY(fit_begin_run(f, t == 0 ? 0.1f : 0.01f, 0.01f, 0.003f, 60, 80, t, st));
        k_copy_rows<<<cdiv(f->S * 201, 256), 256, 0, st>>>(f->mrec, f->gv, f->S, 201, f->T, t);
        LEMO_CUDA(cudaGetLastError());
        LEMO_TRY(fit_run_iters(f, n_iters, st));
    }
    return 0;
}

int lemo_fit_get(LemoFit* h, float* params72, float* losses, void* stream) {
    LEMO_CHECK(h, "null handle");
    Fit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    if (params72) LEMO_CUDA(cudaMemcpyAsync(params72, f->p72, (size_t)f->S * f->T * 72 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (losses) { k_losses_out<<<cdiv(f->S, 32), 32, 0, st>>>(f->acc, f->cfg, f->S, losses); LEMO_CUDA(cudaGetLastError()); }
    return 0;
}

__global__ void k_join_other(const float* __restrict__ z, const float* __restrict__ lh, const float* __restrict__ rh, int B, float* __restrict__ o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 56) return;
    const int b = i / 56, c = i - b * 56;
    o[i] = c < 32 ? z[b * 32 + c] : (c < 44 ? lh[b * 12 + c - 32] : rh[b * 12 + c - 44]);
}

int lemo_fit_get_state(LemoFit* h, float* transl, float* rot6d, float* other, float* g_transl, float* g_rot6d, float* g_other, void* stream) {
    LEMO_CHECK(h, "null handle");
    Fit* f = &h->f;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = f->B;
    if (transl) LEMO_CUDA(cudaMemcpyAsync(transl, f->tr(), B * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (rot6d) LEMO_CUDA(cudaMemcpyAsync(rot6d, f->r6(), B * 6 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (other) k_join_other<<<cdiv(B * 56, 256), 256, 0, st>>>(f->zz(), f->lh(), f->rh(), (int)B, other);
    if (g_transl) LEMO_CUDA(cudaMemcpyAsync(g_transl, f->g_tr(), B * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g_rot6d) LEMO_CUDA(cudaMemcpyAsync(g_rot6d, f->g_r6(), B * 6 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (g_other) k_join_other<<<cdiv(B * 56, 256), 256, 0, st>>>(f->g_zz(), f->g_lh(), f->g_rh(), (int)B, g_other);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int64_t lemo_fit_kernel_launches(const LemoFit* h) { return h ? h->f.launches : 0; }
}
