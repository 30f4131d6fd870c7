// SMPL-X body-model path on sm_100a: pose -> rotations, kinematic chain, fused blend-shape contraction,
// linear blend skinning, output joints, and the adjoint of all of it.
//
// Reference semantics reproduced (paths under the reference tree):
//   human_body_prior/body_model/lbs.py:34-263   (lbs, blend_shapes, vertices2joints, batch_rodrigues,
//                                                batch_rigid_transform)
//   smplx==0.1.26 SMPLX.forward                 (hand PCA, +pose_mean, betas|expression, vertex joints,
//                                                landmarks, +transl)
// B200-first restructuring:
//   * shape and pose blend shapes are ONE contraction  v_posed = v_template + X[B,512] . Wt[512,3V]
//     with X = [R(1..54)-I | betas | expression | 0]; the joint regressor is folded into
//     J_template/J_dirs at model-create time, so nothing of size V is touched per frame except Wt.
//   * rigid transforms are 3x4 (the reference carries 4x4 with a constant last row).
//   * no W.repeat(B) (274 MB at B=119 in the reference), no [B,V,4,4] T tensor.
#include "body.cuh"
#include "gemm.cuh"
#include "../../include/lemo_b200.h"
#include <vector>
#include <cstring>
#include <algorithm>
#include <cstdlib>

namespace lemo {

constexpr int SKB_TV = 256;                 // vertices per tile of k_skin_bwd
constexpr int SKB_PART = NJ * 12 + 3;       // per-CTA partial: dA[55][12] + dtransl[3]
static inline int skin_bwd_tiles(int V, int B) {
    const int ntile = cdiv(V, SKB_TV);
    // full meshes: aim at ~16 CTAs per SM in total (the per-CTA partials are combined in CTA order, so the count does not affect the result's
    // reproducibility, only the summation grouping)
    return ntile <= 2 ? ntile : std::max(1, std::min(ntile, (int)((long long)ntile * B / 2368)));
}
static inline int skin_bwd_ctas(int V, int B) { return cdiv(cdiv(V, SKB_TV), skin_bwd_tiles(V, B)); }
static inline int dx_slices(int V) { return std::max(1, std::min(80, (3 * V) / 416)); }     // full mesh: 75 K-slices x 4 column blocks = 300 CTAs, two per SM


// =============================================================================================
// model
// =============================================================================================
template <typename T>
static int dev_upload(T** dst, const T* src, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    LEMO_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
template <typename T>
static int dev_alloc(T** dst, size_t n) {
    LEMO_CUDA(cudaMalloc((void**)dst, n * sizeof(T)));
    LEMO_CUDA(cudaMemset(*dst, 0, n * sizeof(T)));
    return 0;
}

// Level-ordered tables of the kinematic tree (layout: TREE_* in body.cuh), pure host code -- exported as lemo_host_tree_tables for the CPU tests.
// parents[0] is taken as the root (-1).  Returns 0, or 1 (parents[j] >= j or < 0), 2 (deeper than TREE_MAX_DEPTH), 3 (a level wider than 32).
static int compute_depth(const int* parents, int* depth, int* max_depth);
int build_tree_tables(const int* parents_in, int* t, int* max_depth_out) {
    int parents[NJ], depth[NJ], max_depth = 0;
    for (int j = 0; j < NJ; ++j) parents[j] = parents_in[j];
    parents[0] = -1;
    if (compute_depth(parents, depth, &max_depth)) return 1;
    if (max_depth > TREE_MAX_DEPTH) return 2;
    for (int i = 0; i < TREE_N; ++i) t[i] = 0;
    int pos = 0, k = 0;
    for (int lev = 0; lev <= max_depth + 1; ++lev) {
        t[TREE_OFF + lev] = pos;
        for (int j = 0; j < NJ; ++j)
            if (depth[j] == lev) t[TREE_ORDER + pos++] = j;
    }
    for (int j = 0; j < NJ; ++j) {
        t[TREE_PAR + j] = parents[j];
        t[TREE_DEPTH + j] = depth[j];
        t[TREE_KOFF + j] = k;
        for (int c = j + 1; c < NJ; ++c)
            if (parents[c] == j) t[TREE_KLIST + k++] = c;
    }
    t[TREE_KOFF + NJ] = k;
    for (int i = 0; i < (TREE_MAX_DEPTH + 1) * 32; ++i) t[TREE_LANE + i] = -1;
    for (int lev = 0; lev <= max_depth; ++lev) {
        const int o0 = t[TREE_OFF + lev], o1 = t[TREE_OFF + lev + 1];
        if (o1 - o0 > 32) return 3;
        for (int i = o0; i < o1; ++i) {
            const int j = t[TREE_ORDER + i];
            t[TREE_LANE + lev * 32 + (i - o0)] = j | ((parents[j] + 1) << 8) | (t[TREE_KOFF + j] << 16) | ((t[TREE_KOFF + j + 1] - t[TREE_KOFF + j]) << 24);
        }
    }
    *max_depth_out = max_depth;
    return 0;
}

static int compute_depth(const int* parents, int* depth, int* max_depth) {
    *max_depth = 0;
    for (int j = 0; j < NJ; ++j) {
        if (j == 0) { depth[j] = 0; continue; }
        if (parents[j] < 0 || parents[j] >= j) return 1;
        depth[j] = depth[parents[j]] + 1;
        if (depth[j] > *max_depth) *max_depth = depth[j];
    }
    return 0;
}

static int g_blend_tc = -1;
bool blend_tc_enabled() {
    if (g_blend_tc < 0) { const char* e = getenv("LEMO_BLEND"); g_blend_tc = (e && strcmp(e, "simt") == 0) ? 0 : 1; }
    return g_blend_tc == 1;
}
void blend_tc_set(int on) { g_blend_tc = on ? 1 : 0; }
static int g_skin_tc = -1;
bool skin_tc_enabled() {
    if (g_skin_tc < 0) { const char* e = getenv("LEMO_SKIN"); g_skin_tc = (e && strcmp(e, "simt") == 0) ? 0 : 1; }
    return g_skin_tc == 1;
}
void skin_tc_set(int on) { g_skin_tc = on ? 1 : 0; }
// sparse full-mesh skinning adjoint (Model::sk_*): on whenever the model's weights are sparse; LEMO_SKIN_ADJ=dense / lemo_debug_set_skin_sparse(0)
// force the dense kernel (parity tests compare the two)
static int g_skin_sparse = []() { const char* e = getenv("LEMO_SKIN_ADJ"); return (e && strcmp(e, "dense") == 0) ? 0 : 1; }();
void skin_sparse_set(int on) { g_skin_sparse = on ? 1 : 0; }
// LEMO_SKIN_SMALL=0: loss-row sub-models through the general skinning kernels (A/B measurements)
static int g_skin_small = []() { const char* e = getenv("LEMO_SKIN_SMALL"); return (e && e[0] == '0') ? 0 : 1; }();
// LEMO_DX=gemm routes the full-mesh dX product through the generic SGEMM instead of k_dx_tallk (A/B measurements)
static int g_dx_tallk = []() { const char* e = getenv("LEMO_DX"); return (e && strcmp(e, "gemm") == 0) ? 0 : 1; }();

// K-major transposed copy of Wt + its TMA descriptor for the tensor-core blend GEMM
static int model_setup_tc(Model* m) {
    m->has_tc = false;
    LEMO_TRY(dev_alloc(&m->WtT, (size_t)blend_tc_wtt_floats(3 * m->V)));
    LEMO_CUDA(cudaMemset(m->WtT, 0, (size_t)blend_tc_wtt_floats(3 * m->V) * sizeof(float)));     // rows past 3V of the last tile
    LEMO_TRY(blend_tc_transpose(m->Wt, m->WtT, 3 * m->V));
    LEMO_CUDA(cudaDeviceSynchronize());
    LEMO_TRY(blend_tc_map_w(m->WtT, 3 * m->V, m->map_w));
    m->has_tc = true;
    // tensor-core skinning operand (full meshes only: the loss-row sub-models are a single CTA wave on the CUDA-core kernel)
    m->has_skin_tc = false;
    if (m->V >= 2048) {
        LEMO_TRY(dev_alloc(&m->W2, (size_t)skin_tc_vpad(m->V) * 128));
        LEMO_TRY(skin_tc_prep_w(m->w_jm, m->V, m->W2, m->map_w2));
        LEMO_CUDA(cudaDeviceSynchronize());
        m->has_skin_tc = true;
    }
    return 0;
}

int model_create_from_host(const LemoModelDescC* d, int device, Model** out) {
    LEMO_CHECK(d && out, "null argument");
    LEMO_CHECK(d->n_verts > 0 && d->num_pca_comps > 0 && d->num_pca_comps <= 45, "bad model sizes");
    LEMO_CUDA(cudaSetDevice(device));
    Model* m = new Model();
    m->device = device;
    const int V = m->V = d->n_verts;
    m->npc = d->num_pca_comps;
    m->n_extra = d->h_extra_joint_vids ? d->n_extra_joints : 0;
    m->n_lmk = (d->h_lmk_faces_idx && d->h_faces && d->h_lmk_bary) ? d->n_landmarks : 0;
    memcpy(m->h_parents, d->h_parents, NJ * sizeof(int));
    m->h_parents[0] = -1;
    LEMO_CHECK(compute_depth(m->h_parents, m->h_depth, &m->max_depth) == 0, "parents must satisfy parents[j] < j");

    LEMO_TRY(dev_upload(&m->v_template, d->h_v_template, (size_t)V * 3));
    // Wt = [posedirs ; shapedirs^T ; 0]
    {
        std::vector<float> wt((size_t)XK * 3 * V, 0.f);
        memcpy(wt.data(), d->h_posedirs, (size_t)NPF * 3 * V * sizeof(float));
        for (int v = 0; v < V; ++v)
            for (int k = 0; k < 3; ++k)
                for (int l = 0; l < NBETA; ++l)
                    wt[(size_t)(NPF + l) * 3 * V + 3 * v + k] = d->h_shapedirs[((size_t)v * 3 + k) * NBETA + l];
        LEMO_TRY(dev_upload(&m->Wt, wt.data(), wt.size()));
    }
    {
        std::vector<float> wjm((size_t)NJ * V);
        for (int v = 0; v < V; ++v)
            for (int j = 0; j < NJ; ++j) wjm[(size_t)j * V + v] = d->h_lbs_weights[(size_t)v * NJ + j];
        LEMO_TRY(dev_upload(&m->w_jm, wjm.data(), wjm.size()));
        const int ntile = cdiv(V, SKB_TV);
        if (ntile > 2) {                                              // compact adjoint tables (see Model::sk_*)
            std::vector<int> aoff(ntile + 1), aj;
            for (int t = 0; t < ntile; ++t) {
                aoff[t] = (int)aj.size();
                for (int j = 0; j < NJ; ++j) {
                    bool any = false;
                    for (int v = t * SKB_TV; v < std::min(V, (t + 1) * SKB_TV) && !any; ++v) any = wjm[(size_t)j * V + v] != 0.f;
                    if (any) aj.push_back(j);
                }
            }
            aoff[ntile] = (int)aj.size();
            if (aj.size() * 5 < (size_t)ntile * NJ * 2) {
                const int ns = (int)aj.size();
                std::vector<float> w((size_t)ns * SKB_TV, 0.f);
                std::vector<int> joff(NJ + 1), jslot;
                for (int t = 0; t < ntile; ++t)
                    for (int sl = aoff[t]; sl < aoff[t + 1]; ++sl)
                        for (int u = 0; u < SKB_TV && t * SKB_TV + u < V; ++u) w[(size_t)sl * SKB_TV + u] = wjm[(size_t)aj[sl] * V + t * SKB_TV + u];
                for (int j = 0; j < NJ; ++j) {
                    joff[j] = (int)jslot.size();
                    for (int sl = 0; sl < ns; ++sl) if (aj[sl] == j) jslot.push_back(sl);
                }
                joff[NJ] = (int)jslot.size();
                m->sk_ntile = ntile; m->sk_nslot = ns;
                LEMO_TRY(dev_upload(&m->sk_aoff, aoff.data(), aoff.size()));
                LEMO_TRY(dev_upload(&m->sk_aj, aj.data(), aj.size()));
                LEMO_TRY(dev_upload(&m->sk_w, w.data(), w.size()));
                LEMO_TRY(dev_upload(&m->sk_joff, joff.data(), joff.size()));
                LEMO_TRY(dev_upload(&m->sk_jslot, jslot.data(), jslot.size()));
            }
        }
    }
    {   // fold the joint regressor: J = Jreg.(v_template + shapedirs.beta) = J_template + J_dirs.beta   (double accum)
        std::vector<float> jt(NJ * 3), jd(NJ * 3 * NBETA);
        std::vector<double> acc(3 + 3 * NBETA);
        for (int j = 0; j < NJ; ++j) {
            std::fill(acc.begin(), acc.end(), 0.0);
            const float* jr = d->h_J_regressor + (size_t)j * V;
            for (int v = 0; v < V; ++v) {
                const double w = jr[v];
                if (w == 0.0) continue;
                for (int k = 0; k < 3; ++k) {
                    acc[k] += w * d->h_v_template[(size_t)v * 3 + k];
                    const float* sd = d->h_shapedirs + ((size_t)v * 3 + k) * NBETA;
                    for (int l = 0; l < NBETA; ++l) acc[3 + k * NBETA + l] += w * sd[l];
                }
            }
            for (int k = 0; k < 3; ++k) {
                jt[j * 3 + k] = (float)acc[k];
                for (int l = 0; l < NBETA; ++l) jd[(j * 3 + k) * NBETA + l] = (float)acc[3 + k * NBETA + l];
            }
        }
        LEMO_TRY(dev_upload(&m->J_template, jt.data(), jt.size()));
        LEMO_TRY(dev_upload(&m->J_dirs, jd.data(), jd.size()));
    }
    LEMO_TRY(model_setup_tc(m));
    LEMO_TRY(dev_upload(&m->parents, m->h_parents, NJ));
    LEMO_TRY(dev_upload(&m->depth, m->h_depth, NJ));
    {
        int t[TREE_N], md = 0;
        const int rc = build_tree_tables(m->h_parents, t, &md);
        LEMO_CHECK(rc != 2, "kinematic tree deeper than the level table");
        LEMO_CHECK(rc != 3, "more than 32 joints on one level of the kinematic tree");
        LEMO_CHECK(rc == 0 && md == m->max_depth, "parents must satisfy parents[j] < j");
        LEMO_TRY(dev_upload(&m->tree, t, TREE_N));
    }
    LEMO_TRY(dev_upload(&m->hand_l, d->h_hand_comp_l, (size_t)m->npc * 45));
    LEMO_TRY(dev_upload(&m->hand_r, d->h_hand_comp_r, (size_t)m->npc * 45));
    LEMO_TRY(dev_upload(&m->pose_mean, d->h_pose_mean, 165));
    if (m->n_extra) LEMO_TRY(dev_upload(&m->extra_vids, d->h_extra_joint_vids, m->n_extra));
    if (m->n_lmk) {
        std::vector<int> tri(m->n_lmk * 3);
        for (int l = 0; l < m->n_lmk; ++l) {
            const int f = d->h_lmk_faces_idx[l];
            LEMO_CHECK(f >= 0 && f < d->n_faces, "landmark face index out of range");
            for (int k = 0; k < 3; ++k) tri[l * 3 + k] = d->h_faces[(size_t)f * 3 + k];
        }
        LEMO_TRY(dev_upload(&m->lmk_tri, tri.data(), tri.size()));
        LEMO_TRY(dev_upload(&m->lmk_bary, d->h_lmk_bary, (size_t)m->n_lmk * 3));
    }
    {   // inverse map vertex -> (output joint, weight) for the deterministic joint adjoint (k_joints_bwd)
        std::vector<std::pair<int, std::pair<int, float>>> ent;          // (vertex, (q, w))
        for (int e = 0; e < m->n_extra; ++e) ent.push_back({d->h_extra_joint_vids[e], {NJ + e, 1.f}});
        for (int l = 0; l < m->n_lmk; ++l)
            for (int k = 0; k < 3; ++k)
                ent.push_back({d->h_faces[(size_t)d->h_lmk_faces_idx[l] * 3 + k], {NJ + m->n_extra + l, d->h_lmk_bary[l * 3 + k]}});
        std::stable_sort(ent.begin(), ent.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        std::vector<int> vid, off(1, 0), q;
        std::vector<float> w;
        for (size_t i = 0; i < ent.size(); ++i) {
            if (i == 0 || ent[i].first != ent[i - 1].first) { if (i) off.push_back((int)q.size()); vid.push_back(ent[i].first); }
            q.push_back(ent[i].second.first); w.push_back(ent[i].second.second);
        }
        off.push_back((int)q.size());
        m->n_jv = (int)vid.size();
        if (m->n_jv) {
            LEMO_TRY(dev_upload(&m->jv_vid, vid.data(), vid.size())); LEMO_TRY(dev_upload(&m->jv_off, off.data(), off.size()));
            LEMO_TRY(dev_upload(&m->jv_q, q.data(), q.size())); LEMO_TRY(dev_upload(&m->jv_w, w.data(), w.size()));
        }
    }
    *out = m;
    return 0;
}

__global__ void k_select_rows(const float* __restrict__ vt, const float* __restrict__ Wt, const float* __restrict__ wjm,
                              const int* __restrict__ rows, int V, int n, float* vt_o, float* Wt_o, float* wjm_o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over n*3 columns
    if (i >= n * 3) return;
    const int r = i / 3, k = i - r * 3;
    const int v = rows[r];
    vt_o[i] = vt[v * 3 + k];
    for (int p = 0; p < XK; ++p) Wt_o[(size_t)p * 3 * n + i] = Wt[(size_t)p * 3 * V + 3 * v + k];
    if (k == 0)
        for (int j = 0; j < NJ; ++j) wjm_o[(size_t)j * n + r] = wjm[(size_t)j * V + v];
}

int model_select_rows(const Model* m, const int* rows_host, int n, Model** out) {
    LEMO_CHECK(m && rows_host && out && n > 0, "bad arguments");
    for (int i = 0; i < n; ++i) LEMO_CHECK(rows_host[i] >= 0 && rows_host[i] < m->V, "row index out of range");
    LEMO_CUDA(cudaSetDevice(m->device));
    Model* s = new Model(*m);
    s->is_sub = true;
    s->V = n;
    s->n_extra = 0; s->n_lmk = 0; s->extra_vids = nullptr; s->lmk_tri = nullptr; s->lmk_bary = nullptr;
    s->n_jv = 0; s->jv_vid = nullptr; s->jv_off = nullptr; s->jv_q = nullptr; s->jv_w = nullptr;
    s->sk_ntile = 0; s->sk_nslot = 0; s->sk_aoff = nullptr; s->sk_aj = nullptr; s->sk_w = nullptr; s->sk_joff = nullptr; s->sk_jslot = nullptr;
    int* rows_dev = nullptr;
    LEMO_TRY(dev_upload(&rows_dev, rows_host, n));
    LEMO_TRY(dev_alloc(&s->v_template, (size_t)n * 3));
    LEMO_TRY(dev_alloc(&s->Wt, (size_t)XK * 3 * n));
    LEMO_TRY(dev_alloc(&s->w_jm, (size_t)NJ * n));
    k_select_rows<<<cdiv(n * 3, 128), 128>>>(m->v_template, m->Wt, m->w_jm, rows_dev, m->V, n, s->v_template, s->Wt, s->w_jm);
    LEMO_CUDA(cudaGetLastError());
    LEMO_CUDA(cudaDeviceSynchronize());
    cudaFree(rows_dev);
    s->WtT = nullptr;
    s->W2 = nullptr;          // the copy above must not alias the parent's tensor-core operands (model_free would free them twice)
    s->has_skin_tc = false;
    LEMO_TRY(model_setup_tc(s));
    *out = s;      // shares J_template/J_dirs/parents/hand/pose_mean pointers with the parent model
    return 0;
}

void model_free(Model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->v_template); cudaFree(m->Wt); cudaFree(m->WtT); cudaFree(m->W2); cudaFree(m->w_jm);
    if (!m->is_sub) {
        cudaFree(m->J_template); cudaFree(m->J_dirs); cudaFree(m->parents); cudaFree(m->depth); cudaFree(m->tree);
        cudaFree(m->hand_l); cudaFree(m->hand_r); cudaFree(m->pose_mean);
        cudaFree(m->extra_vids); cudaFree(m->lmk_tri); cudaFree(m->lmk_bary);
        cudaFree(m->jv_vid); cudaFree(m->jv_off); cudaFree(m->jv_q); cudaFree(m->jv_w);
        cudaFree(m->sk_aoff); cudaFree(m->sk_aj); cudaFree(m->sk_w); cudaFree(m->sk_joff); cudaFree(m->sk_jslot);
    }
    delete m;
}

int bodyctx_create(const Model* m, int maxB, bool with_backward, BodyCtx** out) {
    LEMO_CHECK(m && out && maxB > 0, "bad arguments");
    LEMO_CUDA(cudaSetDevice(m->device));
    BodyCtx* c = new BodyCtx();
    c->m = m; c->maxB = maxB; c->device = m->device;
    const size_t B = maxB, V = m->V;
    LEMO_TRY(dev_alloc(&c->full_pose, B * 165));
    LEMO_TRY(dev_alloc(&c->R, B * NJ * 9));
    LEMO_TRY(dev_alloc(&c->X, B * XK));
    LEMO_TRY(dev_alloc(&c->X2, B * 2 * XK));
    LEMO_TRY(blend_tc_map_x(c->X2, maxB, c->map_x));
    LEMO_TRY(dev_alloc(&c->G, B * NJ * 12));
    LEMO_TRY(dev_alloc(&c->A, B * NJ * 12));
    if (m->has_skin_tc) {
        LEMO_TRY(dev_alloc(&c->A2, skin_tc_a2_floats(maxB)));
        LEMO_CUDA(cudaMemset(c->A2, 0, skin_tc_a2_floats(maxB) * sizeof(float)));      // joints 55..63 and frames past B stay zero
        LEMO_TRY(skin_tc_map_a(c->A2, maxB, c->map_a2));
    }
    LEMO_TRY(dev_alloc(&c->Jrest, B * NJ * 3));
    LEMO_TRY(dev_alloc(&c->Jposed, B * NJ * 3));
    LEMO_TRY(dev_alloc(&c->VP, B * 3 * V));
    if (with_backward) {
        LEMO_TRY(dev_alloc(&c->Gv, B * 3 * V));
        LEMO_TRY(dev_alloc(&c->DVP, B * 3 * V));
        LEMO_TRY(dev_alloc(&c->dA, (size_t)NJ * B * 12));
        LEMO_TRY(dev_alloc(&c->dX, B * XK));
        LEMO_TRY(dev_alloc(&c->dR, B * NJ * 9));
        LEMO_TRY(dev_alloc(&c->dJp, B * NJ * 3));
        LEMO_TRY(dev_alloc(&c->dtr, B * 3));
        const int ntile = cdiv(m->V, SKB_TV);
        if (ntile > 2) {          // full meshes: several CTAs per frame in k_skin_bwd and a sliced dX GEMM, combined in a fixed order
            c->part_floats = std::max((size_t)skin_bwd_ctas(m->V, maxB) * B * SKB_PART, (size_t)dx_slices(m->V) * B * XK);
            LEMO_TRY(dev_alloc(&c->part, c->part_floats));
        }
    }
    *out = c;
    return 0;
}

void bodyctx_free(BodyCtx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    float* ptrs[] = {c->full_pose, c->R, c->X, c->X2, c->G, c->A, c->A2, c->Jrest, c->Jposed, c->VP, c->Gv, c->DVP, c->dA, c->dX, c->dR, c->dJp, c->dtr, c->part};
    for (float* p : ptrs) cudaFree(p);
    delete c;
}

}  // namespace lemo
#include "body_dev.cuh"
namespace lemo {

// stand-alone kernels: one 64-thread CTA per frame around the shared device bodies (body_dev.cuh)
__global__ void __launch_bounds__(64) k_pose_to_rot_bwd(PoseK p, PoseGrad g, int B, const float* __restrict__ full_pose, const float* __restrict__ dR) {
    pose_to_rot_bwd_body(p, g, B, full_pose, dR, blockIdx.x);
}
__global__ void __launch_bounds__(64) k_pose_chain_fwd(PoseK p, const float* __restrict__ J_template, const float* __restrict__ J_dirs,
                                                       const int* __restrict__ tree, int max_depth,
                                                       float* __restrict__ full_pose, float* __restrict__ R, float* __restrict__ X,
                                                       float* __restrict__ X2, float* __restrict__ G, float* __restrict__ A,
                                                       float* __restrict__ Jrest, float* __restrict__ Jposed, float* __restrict__ A2) {
    pose_chain_fwd_body(p, J_template, J_dirs, tree, max_depth, full_pose, R, X, X2, G, A, Jrest, Jposed, A2, blockIdx.x);
}
__global__ void __launch_bounds__(64) k_chain_bwd(const float* __restrict__ R, const float* __restrict__ G, const float* __restrict__ Jrest,
                                                  const float* __restrict__ dA, const float* __restrict__ dJp, const float* __restrict__ dX,
                                                  const float* __restrict__ J_dirs, const int* __restrict__ tree, int max_depth, int B,
                                                  int betas_stride,
                                                  float* __restrict__ dR, float* __restrict__ dbetas, float* __restrict__ dexpr) {
    chain_bwd_body(R, G, Jrest, dA, dJp, dX, J_dirs, tree, max_depth, B, betas_stride, dR, dbetas, dexpr, blockIdx.x);
}


// =============================================================================================
// skinning (lbs.py:106-117): one thread per (frame, vertex); blockIdx.y = frame
// VP holds X.Wt on entry, v_posed (= + v_template) on exit (kept for the backward pass).
// =============================================================================================
// (a 4-frames-per-thread variant with float4 smem reads measured 15 % SLOWER on B200: 145 vs 127 us for the full forward)
__global__ void __launch_bounds__(256) k_skin_fwd(const float* __restrict__ A, const float* __restrict__ w_jm,
                                                  const float* __restrict__ v_template, const float* __restrict__ transl,
                                                  int V, float* __restrict__ VP, float* __restrict__ verts) {
    __shared__ float sA[NJ * 12];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < NJ * 12; i += blockDim.x) sA[i] = A[(size_t)b * NJ * 12 + i];
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = 0.f;
    for (int j = 0; j < NJ; ++j) {
        const float w = w_jm[(size_t)j * V + v];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = fmaf(w, sA[j * 12 + k], T[k]);
    }
    float* vp = VP + ((size_t)b * V + v) * 3;
    const float p0 = vp[0] + v_template[v * 3], p1 = vp[1] + v_template[v * 3 + 1], p2 = vp[2] + v_template[v * 3 + 2];
    vp[0] = p0; vp[1] = p1; vp[2] = p2;
    const float t0 = transl ? transl[b * 3] : 0.f, t1 = transl ? transl[b * 3 + 1] : 0.f, t2 = transl ? transl[b * 3 + 2] : 0.f;
    float* o = verts + ((size_t)b * V + v) * 3;
    o[0] = T[0] * p0 + T[1] * p1 + T[2] * p2 + T[3] + t0;
    o[1] = T[4] * p0 + T[5] * p1 + T[6] * p2 + T[7] + t1;
    o[2] = T[8] * p0 + T[9] * p1 + T[10] * p2 + T[11] + t2;
}

// adjoint per (frame, vertex): dvp = T.R^T g ; dT = g (x) [vp;1] ; dtransl += g, fused with the contraction over vertices
// dA[j][b][12] += sum_v w[j][v] dT[v][b][12]  (the first version wrote dT [V, B*12] to HBM -- 48 B per (frame, vertex) at a 46 KB stride --
// and ran a separate split-K GEMM over it: 56 + 28 us per fitting iteration).  One CTA = one frame x `tiles` tiles of 256 vertices:
// per tile, phase A computes dvp and parks dT in shared memory, phase B lets thread t accumulate outputs t, t+256, t+512 of the 660
// (joint, 3x4 entry) pairs over the tile.  A CTA that owns all vertices of its frame (the loss-row sub-models: V = 253) adds in a fixed
// order => bitwise reproducible; several CTAs per frame (full mesh) combine with atomicAdd like the split-K GEMM did.
__global__ void __launch_bounds__(256) k_skin_bwd(const float* __restrict__ A, const float* __restrict__ w_jm,
                                                  const float* __restrict__ VP, const float* __restrict__ Gv, int V, int B, int tiles,
                                                  float* __restrict__ DVP, float* __restrict__ dA, float* __restrict__ dtr,
                                                  float* __restrict__ part) {
    __shared__ float sA[NJ * 12];
    __shared__ __align__(16) float s_dt[SKB_TV * 12];
    __shared__ float s_w[NJ][33];
    __shared__ float sred[32];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < NJ * 12; i += blockDim.x) sA[i] = A[(size_t)b * NJ * 12 + i];
    const int jB = threadIdx.x % NJ, sB = threadIdx.x / NJ;       // phase-B role (threads < 220)
    float acc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    float gs0 = 0.f, gs1 = 0.f, gs2 = 0.f;
    const int t0 = blockIdx.x * tiles;
    for (int tl = 0; tl < tiles; ++tl) {
        const int v0 = (t0 + tl) * SKB_TV;
        if (v0 >= V) break;
        __syncthreads();                                   // sA ready / previous tile's s_dt consumed
        const int v = v0 + threadIdx.x;
        float dt[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) dt[k] = 0.f;
        if (v < V) {
            float T[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) T[k] = 0.f;
            for (int j = 0; j < NJ; ++j) {
                const float w = w_jm[(size_t)j * V + v];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) T[i * 3 + c] = fmaf(w, sA[j * 12 + i * 4 + c], T[i * 3 + c]);
            }
            const float* g = Gv + ((size_t)b * V + v) * 3;
            const float g0 = g[0], g1 = g[1], g2 = g[2];
            gs0 += g0; gs1 += g1; gs2 += g2;
            const float* vp = VP + ((size_t)b * V + v) * 3;
            const float p[4] = {vp[0], vp[1], vp[2], 1.f};
            float* dvp = DVP + ((size_t)b * V + v) * 3;
            dvp[0] = T[0] * g0 + T[3] * g1 + T[6] * g2;
            dvp[1] = T[1] * g0 + T[4] * g1 + T[7] * g2;
            dvp[2] = T[2] * g0 + T[5] * g1 + T[8] * g2;
            const float gg[3] = {g0, g1, g2};
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) dt[i * 4 + c] = gg[i] * p[c];
        }
#pragma unroll
        for (int k4 = 0; k4 < 3; ++k4)
            *reinterpret_cast<float4*>(&s_dt[threadIdx.x * 12 + k4 * 4]) = make_float4(dt[k4 * 4], dt[k4 * 4 + 1], dt[k4 * 4 + 2], dt[k4 * 4 + 3]);
        __syncthreads();
        // phase B, register tiled: thread (joint jB, vertex residue sB) keeps the 12 entries of dA[jB] for vertices u = sB mod 4.
        // The weight chunk [55 joints][32 vertices] is staged through shared memory (coalesced global read, conflict-free pitch 33),
        // dT[u] is three broadcast 128-bit reads: 4 shared-memory wavefronts per 12 FMAs.  (One thread per (joint, entry) with a
        // global weight load and a shared dT load per FMA was load-issue bound: 65 us per fitting step at B = 960.)
        const int nv = min(SKB_TV, V - v0);
        for (int c0 = 0; c0 < nv; c0 += 32) {
            for (int idx = threadIdx.x; idx < NJ * 32; idx += 256) {
                const int j = idx >> 5, u = idx & 31;
                s_w[j][u] = (c0 + u < nv) ? __ldg(w_jm + (size_t)j * V + v0 + c0 + u) : 0.f;
            }
            __syncthreads();
            if (threadIdx.x < NJ * 4) {
#pragma unroll
                for (int uu = sB; uu < 32; uu += 4) {
                    const float w = s_w[jB][uu];
                    const float4* d = reinterpret_cast<const float4*>(&s_dt[(c0 + uu) * 12]);
                    const float4 d0 = d[0], d1 = d[1], d2 = d[2];
                    acc[0] = fmaf(w, d0.x, acc[0]); acc[1] = fmaf(w, d0.y, acc[1]); acc[2] = fmaf(w, d0.z, acc[2]); acc[3] = fmaf(w, d0.w, acc[3]);
                    acc[4] = fmaf(w, d1.x, acc[4]); acc[5] = fmaf(w, d1.y, acc[5]); acc[6] = fmaf(w, d1.z, acc[6]); acc[7] = fmaf(w, d1.w, acc[7]);
                    acc[8] = fmaf(w, d2.x, acc[8]); acc[9] = fmaf(w, d2.y, acc[9]); acc[10] = fmaf(w, d2.z, acc[10]); acc[11] = fmaf(w, d2.w, acc[11]);
                }
            }
            __syncthreads();
        }
    }
    // combine the four vertex residues in a fixed order (s_dt is free now) and add into dA
    __syncthreads();
    if (threadIdx.x < NJ * 4) {
#pragma unroll
        for (int k = 0; k < 12; ++k) s_dt[sB * (NJ * 12) + jB * 12 + k] = acc[k];
    }
    __syncthreads();
    // one CTA per frame (loss-row sub-models): add straight into dA / dtr (single contribution per address: deterministic).
    // several CTAs per frame (full mesh): park this CTA's 660 + 3 sums in part[blockIdx.x][b][.]; k_skin_bwd_reduce adds them in CTA order.
    float* pz = part ? part + ((size_t)blockIdx.x * B + b) * SKB_PART : nullptr;
    for (int t = threadIdx.x; t < NJ * 12; t += 256) {
        const float v = (s_dt[t] + s_dt[NJ * 12 + t]) + (s_dt[2 * NJ * 12 + t] + s_dt[3 * NJ * 12 + t]);
        const int j = t / 12, k = t - j * 12;
        if (pz) pz[t] = v;
        else atomicAdd(&dA[(size_t)j * B * 12 + (size_t)b * 12 + k], v);
    }
    float s;
    s = block_sum(gs0, sred); if (threadIdx.x == 0) { if (pz) pz[NJ * 12] = s; else atomicAdd(&dtr[b * 3], s); }
    s = block_sum(gs1, sred); if (threadIdx.x == 0) { if (pz) pz[NJ * 12 + 1] = s; else atomicAdd(&dtr[b * 3 + 1], s); }
    s = block_sum(gs2, sred); if (threadIdx.x == 0) { if (pz) pz[NJ * 12 + 2] = s; else atomicAdd(&dtr[b * 3 + 2], s); }
}
// ---- loss-row sub-models (V <= SKS_V rows: the 253 marker + foot rows of the AMASS fits, one CTA per frame in the general kernels = 960
// CTAs whose whole life is latency: 55 dependent-ish global weight loads per thread, staging loops, 46 us for 90 MFLOP).  Here a CTA keeps
// the sub-model's weights [55][V] in (dynamic) shared memory (pitch 257: conflict-free by row and by column) and walks frames b = blockIdx.x,
// blockIdx.x + gridDim.x, ...  Same summation orders as k_skin_fwd / k_skin_bwd: bit-identical results.
constexpr int SKS_V = 256, SKS_P = 257;
constexpr size_t SKS_FWD_SMEM = (size_t)(NJ * SKS_P + NJ * 12 + 4) * sizeof(float);
constexpr size_t SKS_BWD_SMEM = (size_t)(NJ * SKS_P + 3 + NJ * 12 + SKS_V * 12 + 4 * NJ * 12 + 32) * sizeof(float);
__global__ void __launch_bounds__(256) k_skin_fwd_small(const float* __restrict__ A, const float* __restrict__ w_jm,
                                                        const float* __restrict__ v_template, const float* __restrict__ transl,
                                                        int V, int B, float* __restrict__ VP, float* __restrict__ verts) {
    extern __shared__ __align__(16) float sks_smem[];
    float* sA = sks_smem;                                  // [660], 16-byte aligned
    float* s_w = sks_smem + NJ * 12 + 4;                   // [55][SKS_P]
    for (int i = threadIdx.x; i < NJ * V; i += 256) { const int j = i / V, v = i - j * V; s_w[j * SKS_P + v] = w_jm[i]; }
    const int v = threadIdx.x;
    float vt0 = 0.f, vt1 = 0.f, vt2 = 0.f;
    if (v < V) { vt0 = v_template[v * 3]; vt1 = v_template[v * 3 + 1]; vt2 = v_template[v * 3 + 2]; }
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < NJ * 12; i += 256) sA[i] = A[(size_t)b * NJ * 12 + i];
        __syncthreads();
        if (v >= V) continue;
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) T[k] = 0.f;
#pragma unroll 5
        for (int j = 0; j < NJ; ++j) {
            const float w = s_w[j * SKS_P + v];
            const float4* a = reinterpret_cast<const float4*>(sA + j * 12);
            const float4 a0 = a[0], a1 = a[1], a2 = a[2];
            T[0] = fmaf(w, a0.x, T[0]); T[1] = fmaf(w, a0.y, T[1]); T[2] = fmaf(w, a0.z, T[2]); T[3] = fmaf(w, a0.w, T[3]);
            T[4] = fmaf(w, a1.x, T[4]); T[5] = fmaf(w, a1.y, T[5]); T[6] = fmaf(w, a1.z, T[6]); T[7] = fmaf(w, a1.w, T[7]);
            T[8] = fmaf(w, a2.x, T[8]); T[9] = fmaf(w, a2.y, T[9]); T[10] = fmaf(w, a2.z, T[10]); T[11] = fmaf(w, a2.w, T[11]);
        }
        float* vp = VP + ((size_t)b * V + v) * 3;
        const float p0 = vp[0] + vt0, p1 = vp[1] + vt1, p2 = vp[2] + vt2;
        vp[0] = p0; vp[1] = p1; vp[2] = p2;
        const float t0 = transl ? transl[b * 3] : 0.f, t1 = transl ? transl[b * 3 + 1] : 0.f, t2 = transl ? transl[b * 3 + 2] : 0.f;
        float* o = verts + ((size_t)b * V + v) * 3;
        o[0] = T[0] * p0 + T[1] * p1 + T[2] * p2 + T[3] + t0;
        o[1] = T[4] * p0 + T[5] * p1 + T[6] * p2 + T[7] + t1;
        o[2] = T[8] * p0 + T[9] * p1 + T[10] * p2 + T[11] + t2;
    }
}
__global__ void __launch_bounds__(256) k_skin_bwd_small(const float* __restrict__ A, const float* __restrict__ w_jm,
                                                        const float* __restrict__ VP, const float* __restrict__ Gv, int V, int B,
                                                        float* __restrict__ DVP, float* __restrict__ dA, float* __restrict__ dtr) {
    extern __shared__ __align__(16) float sks_smem[];
    float* s_dt = sks_smem;                                // [SKS_V][12], 16-byte aligned
    float* sA = s_dt + SKS_V * 12;                         // [660]
    float (*s_acc)[NJ * 12] = reinterpret_cast<float (*)[NJ * 12]>(sA + NJ * 12);
    float* sred = sA + NJ * 12 + 4 * NJ * 12;              // [32]
    float* s_w = sred + 32;                                // [55][SKS_P]
    for (int i = threadIdx.x; i < NJ * V; i += 256) { const int j = i / V, v = i - j * V; s_w[j * SKS_P + v] = w_jm[i]; }
    const int v = threadIdx.x;
    const int jB = threadIdx.x % NJ, sB = threadIdx.x / NJ;       // phase-B role (threads < 220)
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = threadIdx.x; i < NJ * 12; i += 256) sA[i] = A[(size_t)b * NJ * 12 + i];
        __syncthreads();
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        if (v < V) {
            float T[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) T[k] = 0.f;
#pragma unroll 5
            for (int j = 0; j < NJ; ++j) {
                const float w = s_w[j * SKS_P + v];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c = 0; c < 3; ++c) T[i * 3 + c] = fmaf(w, sA[j * 12 + i * 4 + c], T[i * 3 + c]);
            }
            const float* g = Gv + ((size_t)b * V + v) * 3;
            g0 = g[0]; g1 = g[1]; g2 = g[2];
            const float* vp = VP + ((size_t)b * V + v) * 3;
            const float p[4] = {vp[0], vp[1], vp[2], 1.f};
            float* dvp = DVP + ((size_t)b * V + v) * 3;
            dvp[0] = T[0] * g0 + T[3] * g1 + T[6] * g2;
            dvp[1] = T[1] * g0 + T[4] * g1 + T[7] * g2;
            dvp[2] = T[2] * g0 + T[5] * g1 + T[8] * g2;
            const float gg[3] = {g0, g1, g2};
#pragma unroll
            for (int i = 0; i < 3; ++i)
                *reinterpret_cast<float4*>(&s_dt[v * 12 + i * 4]) = make_float4(gg[i] * p[0], gg[i] * p[1], gg[i] * p[2], gg[i] * p[3]);
        }
        __syncthreads();
        if (threadIdx.x < NJ * 4) {
            float acc[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) acc[k] = 0.f;
            for (int u = sB; u < V; u += 4) {
                const float w = s_w[jB * SKS_P + u];
                const float4* d = reinterpret_cast<const float4*>(&s_dt[u * 12]);
                const float4 d0 = d[0], d1 = d[1], d2 = d[2];
                acc[0] = fmaf(w, d0.x, acc[0]); acc[1] = fmaf(w, d0.y, acc[1]); acc[2] = fmaf(w, d0.z, acc[2]); acc[3] = fmaf(w, d0.w, acc[3]);
                acc[4] = fmaf(w, d1.x, acc[4]); acc[5] = fmaf(w, d1.y, acc[5]); acc[6] = fmaf(w, d1.z, acc[6]); acc[7] = fmaf(w, d1.w, acc[7]);
                acc[8] = fmaf(w, d2.x, acc[8]); acc[9] = fmaf(w, d2.y, acc[9]); acc[10] = fmaf(w, d2.z, acc[10]); acc[11] = fmaf(w, d2.w, acc[11]);
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) s_acc[sB][jB * 12 + k] = acc[k];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < NJ * 12; t += 256) {        // this CTA is the only writer of frame b's entries: plain read-modify-write
            const float a = (s_acc[0][t] + s_acc[1][t]) + (s_acc[2][t] + s_acc[3][t]);
            const int j = t / 12, k = t - j * 12;
            dA[(size_t)j * B * 12 + (size_t)b * 12 + k] += a;
        }
        float q;
        q = block_sum(g0, sred); if (threadIdx.x == 0) dtr[b * 3] += q;
        q = block_sum(g1, sred); if (threadIdx.x == 0) dtr[b * 3 + 1] += q;
        q = block_sum(g2, sred); if (threadIdx.x == 0) dtr[b * 3 + 2] += q;
    }
}

// ---- compact skinning adjoint (full meshes with sparse weights; same result as k_skin_bwd up to summation order, fixed order throughout).
// One CTA per (tile of 256 vertices, frame).  Phase A, thread = vertex: T = sum over the tile's ACTIVE joints of w A_j.R, d v_posed = T^T g,
// dT = g (x) [v_posed, 1] parked in shared memory.  Phase B, thread = (active-joint slot a, vertex residue r of 16): the 12 entries of
// dA[joint(a)] over the vertices u = r mod 16, then the 16 residues are added in order and the slot's 12 sums are written to
// partc[frame][slot]; k_skin_bwd_act_reduce adds the slots of a joint in tile order.
__global__ void __launch_bounds__(256) k_skin_bwd_act(const float* __restrict__ A, const int* __restrict__ aoff, const int* __restrict__ aj,
                                                      const float* __restrict__ wact, const float* __restrict__ VP, const float* __restrict__ Gv,
                                                      int V, int B, int nslot, float* __restrict__ DVP, float* __restrict__ partc,
                                                      float* __restrict__ part_tr) {
    __shared__ float sA[NJ * 12];
    __shared__ __align__(16) float s_dt[SKB_TV * 12];
    __shared__ float s_part[16][16][12];
    __shared__ float sred[32];
    __shared__ int s_j[NJ];
    const int b = blockIdx.y, tile = blockIdx.x;
    const int s0 = aoff[tile], na = aoff[tile + 1] - s0;
    for (int i = threadIdx.x; i < NJ * 12; i += 256) sA[i] = A[(size_t)b * NJ * 12 + i];
    if (threadIdx.x < na) s_j[threadIdx.x] = aj[s0 + threadIdx.x];
    __syncthreads();
    const int u = threadIdx.x, v = tile * SKB_TV + u;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    float dt[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) dt[k] = 0.f;
    if (v < V) {
        float T[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) T[k] = 0.f;
        for (int a = 0; a < na; ++a) {
            const float w = __ldg(wact + (size_t)(s0 + a) * SKB_TV + u);
            const float* aa = sA + s_j[a] * 12;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c) T[i * 3 + c] = fmaf(w, aa[i * 4 + c], T[i * 3 + c]);
        }
        const float* g = Gv + ((size_t)b * V + v) * 3;
        g0 = g[0]; g1 = g[1]; g2 = g[2];
        const float* vp = VP + ((size_t)b * V + v) * 3;
        const float p[4] = {vp[0], vp[1], vp[2], 1.f};
        float* dvp = DVP + ((size_t)b * V + v) * 3;
        dvp[0] = T[0] * g0 + T[3] * g1 + T[6] * g2;
        dvp[1] = T[1] * g0 + T[4] * g1 + T[7] * g2;
        dvp[2] = T[2] * g0 + T[5] * g1 + T[8] * g2;
        const float gg[3] = {g0, g1, g2};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) dt[i * 4 + c] = gg[i] * p[c];
    }
#pragma unroll
    for (int k4 = 0; k4 < 3; ++k4)
        *reinterpret_cast<float4*>(&s_dt[u * 12 + k4 * 4]) = make_float4(dt[k4 * 4], dt[k4 * 4 + 1], dt[k4 * 4 + 2], dt[k4 * 4 + 3]);
    __syncthreads();
    const int a = threadIdx.x >> 4, r = threadIdx.x & 15;
    for (int a0 = 0; a0 < na; a0 += 16) {
        float acc[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) acc[k] = 0.f;
        if (a0 + a < na) {
            const float* wrow = wact + (size_t)(s0 + a0 + a) * SKB_TV;
#pragma unroll 4
            for (int uu = r; uu < SKB_TV; uu += 16) {
                const float w = __ldg(wrow + uu);
                const float4* d = reinterpret_cast<const float4*>(&s_dt[uu * 12]);
                const float4 d0 = d[0], d1 = d[1], d2 = d[2];
                acc[0] = fmaf(w, d0.x, acc[0]); acc[1] = fmaf(w, d0.y, acc[1]); acc[2] = fmaf(w, d0.z, acc[2]); acc[3] = fmaf(w, d0.w, acc[3]);
                acc[4] = fmaf(w, d1.x, acc[4]); acc[5] = fmaf(w, d1.y, acc[5]); acc[6] = fmaf(w, d1.z, acc[6]); acc[7] = fmaf(w, d1.w, acc[7]);
                acc[8] = fmaf(w, d2.x, acc[8]); acc[9] = fmaf(w, d2.y, acc[9]); acc[10] = fmaf(w, d2.z, acc[10]); acc[11] = fmaf(w, d2.w, acc[11]);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) s_part[a][r][k] = acc[k];
        __syncthreads();
        if (threadIdx.x < 16 * 12) {
            const int aa = threadIdx.x / 12, k = threadIdx.x - aa * 12;
            if (a0 + aa < na) {
                float sum = 0.f;
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) sum += s_part[aa][rr][k];
                partc[((size_t)b * nslot + s0 + a0 + aa) * 12 + k] = sum;
            }
        }
        __syncthreads();
    }
    float* pt = part_tr + ((size_t)tile * B + b) * 3;
    float s;
    s = block_sum(g0, sred); if (threadIdx.x == 0) pt[0] = s;
    s = block_sum(g1, sred); if (threadIdx.x == 0) pt[1] = s;
    s = block_sum(g2, sred); if (threadIdx.x == 0) pt[2] = s;
}
__global__ void k_skin_bwd_act_reduce(const float* __restrict__ partc, const int* __restrict__ joff, const int* __restrict__ jslot, int nslot,
                                      const float* __restrict__ part_tr, int ntile, int B, float* __restrict__ dA, float* __restrict__ dtr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * SKB_PART) return;
    const int b = i / SKB_PART, t = i - b * SKB_PART;
    float a = 0.f;
    if (t < NJ * 12) {
        const int j = t / 12, k = t - j * 12;
        for (int q = joff[j]; q < joff[j + 1]; ++q) a += partc[((size_t)b * nslot + jslot[q]) * 12 + k];
        dA[(size_t)j * B * 12 + (size_t)b * 12 + k] += a;
    } else {
        for (int p = 0; p < ntile; ++p) a += part_tr[((size_t)p * B + b) * 3 + (t - NJ * 12)];
        dtr[b * 3 + (t - NJ * 12)] += a;
    }
}

__global__ void k_skin_bwd_reduce(const float* __restrict__ part, int nparts, int B, float* __restrict__ dA, float* __restrict__ dtr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * SKB_PART) return;
    const int b = i / SKB_PART, t = i - b * SKB_PART;
    float a = 0.f;
    for (int p = 0; p < nparts; ++p) a += part[((size_t)p * B + b) * SKB_PART + t];
    if (t < NJ * 12) { const int j = t / 12, k = t - j * 12; dA[(size_t)j * B * 12 + (size_t)b * 12 + k] += a; }
    else dtr[b * 3 + (t - NJ * 12)] += a;
}
// C[M,N] += sum over slices of the partial products parked by a splitk == 2 GEMM, in slice order
// dX partials of the full mesh: part[z][M][512] = DVP[M][Kz] . Wt[512][Kz]^T for K slice z (M = frames <= 128, K = 3V = 31425).
// Both operands are K-fast in memory (a dot-product GEMM), M is small and K huge, so the tile is the whole M x 128 columns and the grid is
// (4 column blocks, K slices).  8 x 8 outputs per thread, as two 4-wide halves 64 apart so that the shared-memory fragments are read
// with conflict-free 128-bit loads: 4 LDS.128 per 64 FMAs (the generic 4 x 4 kernel needs 2 per 16 and is bound by shared-memory
// bandwidth: 161 us for this product at B = 100).  Global loads are register-prefetched one K block ahead.
constexpr int DXK = 16, DXP = 132, DXS = 3;
constexpr size_t DX_SMEM = (size_t)DXS * 2 * DXK * DXP * sizeof(float);
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = pred ? 4 : 0;                     // src-size 0: the 4 bytes are zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
// Operand blocks travel global -> shared with 4-byte cp.async (the transposed [k][row] layout rules out wider copies; the row pitch of
// both operands, 3V floats, is not 16-byte aligned anyway), three stages deep, so no registers are spent on staging and two CTAs fit an SM.
__global__ void __launch_bounds__(256, 2) k_dx_tallk(const float* __restrict__ A, const float* __restrict__ Bm, int M, int K, int kslice,
                                                     float* __restrict__ part) {
    extern __shared__ __align__(16) float dx_smem[];
    float (*As)[DXK][DXP] = reinterpret_cast<float (*)[DXK][DXP]>(dx_smem);
    float (*Bs)[DXK][DXP] = reinterpret_cast<float (*)[DXK][DXP]>(dx_smem + DXS * DXK * DXP);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * 128, z = blockIdx.y;
    const int k_begin = z * kslice, k_end = min(K, k_begin + kslice);
    const int nblk = (k_end - k_begin + DXK - 1) / DXK;
    const int lk = tid & 15, lr = tid >> 4;            // loader role: column lk of the K block, rows lr + 16 i
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    auto issue = [&](int blk) {
        if (blk < nblk) {
            const int st = blk % DXS, gk = k_begin + blk * DXK + lk;
            const bool kin = gk < k_end;
            const int gkc = kin ? gk : k_begin;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = lr + 16 * i;
                cp_async4(&As[st][lk][r], A + (size_t)min(r, M - 1) * K + gkc, kin && r < M);
                cp_async4(&Bs[st][lk][r], Bm + (size_t)(n0 + r) * K + gkc, kin);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");      // (an empty group past the end keeps the wait count uniform)
    };
    issue(0);
    issue(1);
    for (int blk = 0; blk < nblk; ++blk) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");      // block `blk` has landed (this thread's copies)
        __syncthreads();                                          // ... everybody's; and stage (blk + 2) % 3 is no longer being read
        issue(blk + 2);
        const int st = blk % DXS;
#pragma unroll
        for (int k = 0; k < DXK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[st][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[st][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float* out = part + (size_t)z * M * XK;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = (i < 4 ? 0 : 64) + ty * 4 + (i & 3);
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh)
            *reinterpret_cast<float4*>(out + (size_t)m * XK + n0 + jh * 64 + tx * 4) =
                make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
    }
}

__global__ void k_slices_reduce(const float* __restrict__ part, int nz, long long mn, float* __restrict__ C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn) return;
    float a = 0.f;
    for (int z = 0; z < nz; ++z) a += part[(size_t)z * mn + i];
    C[i] += a;
}

// =============================================================================================
// output joints (smplx SMPLX.forward): 55 posed joints, 21 vertex joints, 51 barycentric landmarks
// =============================================================================================
__global__ void k_joints_fwd(const float* __restrict__ Jposed, const float* __restrict__ transl, const float* __restrict__ verts,
                             const int* __restrict__ extra, int n_extra, const int* __restrict__ tri, const float* __restrict__ bary,
                             int n_lmk, int V, int B, float* __restrict__ joints) {
    const int nout = NJ + n_extra + n_lmk;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * nout) return;
    const int b = i / nout, q = i - b * nout;
    float o[3];
    if (q < NJ) {
        for (int k = 0; k < 3; ++k) o[k] = Jposed[((size_t)b * NJ + q) * 3 + k] + (transl ? transl[b * 3 + k] : 0.f);
    } else if (q < NJ + n_extra) {
        const float* s = verts + ((size_t)b * V + extra[q - NJ]) * 3;
        for (int k = 0; k < 3; ++k) o[k] = s[k];
    } else {
        const int l = q - NJ - n_extra;
        o[0] = o[1] = o[2] = 0.f;
        for (int c = 0; c < 3; ++c) {
            const float w = bary[l * 3 + c];
            const float* s = verts + ((size_t)b * V + tri[l * 3 + c]) * 3;
            for (int k = 0; k < 3; ++k) o[k] = fmaf(w, s[k], o[k]);
        }
    }
    for (int k = 0; k < 3; ++k) joints[((size_t)b * nout + q) * 3 + k] = o[k];
}

// adjoint of k_joints_fwd, gather form (no atomics): (1) the 55 posed joints: dJp = g (single writer) and dtransl += sum_q g in joint
// order; (2) every distinct vertex referenced by a vertex joint or a landmark adds its entries of the model's inverse map in order.
__global__ void k_joints_bwd(const float* __restrict__ dj, int nout, int n_jv, const int* __restrict__ jv_vid, const int* __restrict__ jv_off,
                             const int* __restrict__ jv_q, const float* __restrict__ jv_w, int V, int B, float* __restrict__ Gv,
                             float* __restrict__ dJp, float* __restrict__ dtr) {
    const int per = NJ * 3 + 3 + n_jv;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * per) return;
    const int b = i / per, t = i - b * per;
    const float* g = dj + (size_t)b * nout * 3;
    if (t < NJ * 3) dJp[(size_t)b * NJ * 3 + t] += g[t];
    else if (t < NJ * 3 + 3) {
        const int k = t - NJ * 3;
        float a = 0.f;
        for (int q = 0; q < NJ; ++q) a += g[q * 3 + k];         // posed joints carry + transl directly (vertex joints get it through Gv)
        dtr[b * 3 + k] += a;
    } else {
        const int u = t - NJ * 3 - 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int e = jv_off[u]; e < jv_off[u + 1]; ++e) {
            const float w = jv_w[e];
            const float* gq = g + jv_q[e] * 3;
            a0 += w * gq[0]; a1 += w * gq[1]; a2 += w * gq[2];
        }
        float* d = Gv + ((size_t)b * V + jv_vid[u]) * 3;
        d[0] += a0; d[1] += a1; d[2] += a2;
    }
}

__global__ void k_gather_rows(const float* __restrict__ src, const int* __restrict__ idx, int B, int V, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n * 3) return;
    const int k = i % 3, r = (i / 3) % n, b = i / (3 * n);
    out[i] = src[((size_t)b * V + idx[r]) * 3 + k];
}
__global__ void k_scatter_rows_add(const float* __restrict__ g, const int* __restrict__ idx, int B, int V, int n, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * n * 3) return;
    const int k = i % 3, r = (i / 3) % n, b = i / (3 * n);
    atomicAdd(&dst[((size_t)b * V + idx[r]) * 3 + k], g[i]);
}

// =============================================================================================
// host-side composition
// =============================================================================================
static PoseK make_posek(const Model* m, const PoseIn& in) {
    PoseK p;
    p.in = in; p.hand_l = m->hand_l; p.hand_r = m->hand_r; p.pose_mean = m->pose_mean; p.npc = m->npc;
    return p;
}

int body_pose_forward(BodyCtx* c, const PoseIn& in, int B, cudaStream_t st) {
    LEMO_CHECK(c && B > 0 && B <= c->maxB, "batch exceeds the size this body handle was created for");
    const Model* m = c->m;
    k_pose_chain_fwd<<<B, 64, 0, st>>>(make_posek(m, in), m->J_template, m->J_dirs, m->tree, m->max_depth, c->full_pose,
                                        c->R, c->X, c->X2, c->G, c->A, c->Jrest, c->Jposed, c->A2);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

static bool both_tc(const Model* m, const BodyCtx* ps) {
    return m->has_tc && blend_tc_enabled() && m->has_skin_tc && ps->A2 && skin_tc_enabled();
}

// VP[B,3V] = X[B,512] . Wt[512,3V]: tcgen05 TF32 GEMM (blend_tc.cu); LEMO_BLEND=simt selects the CUDA-core GEMM (debug A/B).
// When skinning runs on the tensor cores too, the epilogue adds v_template so that VP = v_posed (skin_tc.cu reads it as is).
int body_blend_forward(BodyCtx* c, const BodyCtx* ps, int B, cudaStream_t st) {
    LEMO_CHECK(c && ps && B > 0 && B <= c->maxB, "bad arguments");
    const Model* m = c->m;
    const int V = m->V;
    if (both_tc(m, ps)) return blend_tc_launch_bias(ps->map_x, m->map_w, c->VP, B, 3 * V, m->v_template, st);
    if (m->has_tc && blend_tc_enabled()) return blend_tc_launch(ps->map_x, m->map_w, c->VP, B, 3 * V, st);
    GemmP g = gemm_rowmajor(ps->X, m->Wt, c->VP, B, 3 * V, XK, false);
    return gemm_launch(g, st);
}

// skinning (+ output joints): needs VP from body_blend_forward and A / Jposed from body_chain_forward
static int body_apply_forward(BodyCtx* c, const BodyCtx* ps, const PoseIn& in, int B, float* verts, float* joints, cudaStream_t st) {
    const Model* m = c->m;
    const int V = m->V;
    if (both_tc(m, ps)) LEMO_TRY(skin_tc_launch(m->map_w2, ps->map_a2, c->VP, in.transl, V, B, verts, st));
    else if (V <= SKS_V && g_skin_small)
    {
        static bool cfg = false;
        if (!cfg) { LEMO_CUDA(cudaFuncSetAttribute(k_skin_fwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKS_FWD_SMEM)); cfg = true; }
        k_skin_fwd_small<<<std::min(B, 148 * 3), 256, SKS_FWD_SMEM, st>>>(ps->A, m->w_jm, m->v_template, in.transl, V, B, c->VP, verts);
    }
    else k_skin_fwd<<<dim3(cdiv(V, 256), B), 256, 0, st>>>(ps->A, m->w_jm, m->v_template, in.transl, V, c->VP, verts);
    if (joints) {
        LEMO_CHECK(!m->is_sub, "output joints need the full model");
        const int nout = NJ + m->n_extra + m->n_lmk;
        k_joints_fwd<<<cdiv(B * nout, 128), 128, 0, st>>>(ps->Jposed, in.transl, verts, m->extra_vids, m->n_extra, m->lmk_tri,
                                                            m->lmk_bary, m->n_lmk, V, B, joints);
    }
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int body_skin_forward(BodyCtx* c, const BodyCtx* ps, const PoseIn& in, int B, float* verts, float* joints, cudaStream_t st) {
    LEMO_CHECK(c && ps && verts && B > 0 && B <= c->maxB, "bad arguments");
    LEMO_TRY(body_blend_forward(c, ps, B, st));
    return body_apply_forward(c, ps, in, B, verts, joints, st);
}

int body_grad_begin(BodyCtx* ps, int B, cudaStream_t st) {
    LEMO_CHECK(ps && ps->dA, "body handle was created without backward buffers");
    LEMO_CUDA(cudaMemsetAsync(ps->dA, 0, (size_t)NJ * B * 12 * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(ps->dX, 0, (size_t)B * XK * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(ps->dJp, 0, (size_t)B * NJ * 3 * sizeof(float), st));
    LEMO_CUDA(cudaMemsetAsync(ps->dtr, 0, (size_t)B * 3 * sizeof(float), st));
    return 0;
}

int body_skin_backward(BodyCtx* c, BodyCtx* ps, int B, const float* d_verts, const float* d_joints, cudaStream_t st) {
    LEMO_CHECK(c && ps && c->Gv && ps->dA, "body handle was created without backward buffers");
    const Model* m = c->m;
    const int V = m->V;
    if (d_verts == c->Gv) {}      // the caller assembled the vertex gradient in place (fit_prox.cu): no 12 MB copy
    else if (d_verts) LEMO_CUDA(cudaMemcpyAsync(c->Gv, d_verts, (size_t)B * V * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else LEMO_CUDA(cudaMemsetAsync(c->Gv, 0, (size_t)B * V * 3 * sizeof(float), st));
    if (d_joints) {
        LEMO_CHECK(!m->is_sub, "output joints need the full model");
        const int nout = NJ + m->n_extra + m->n_lmk;
        k_joints_bwd<<<cdiv(B * (NJ * 3 + 3 + m->n_jv), 128), 128, 0, st>>>(d_joints, nout, m->n_jv, m->jv_vid, m->jv_off, m->jv_q, m->jv_w, V, B,
                                                                              c->Gv, ps->dJp, ps->dtr);
    }
    {
        // vertex tiles per CTA: everything in one CTA per frame while that still fills the GPU (sub-models), else ~6 CTAs per frame
        const int ntile = cdiv(V, SKB_TV);
        const int tiles = skin_bwd_tiles(V, B), ctas = cdiv(ntile, tiles);
        float* part = (ctas > 1 && c->part && (size_t)ctas * B * SKB_PART <= c->part_floats) ? c->part : nullptr;
        const bool sparse = g_skin_sparse && m->sk_nslot > 0 && c->part &&
                            (size_t)B * m->sk_nslot * 12 + (size_t)m->sk_ntile * B * 3 <= c->part_floats;
        if (sparse) {
            float* partc = c->part;
            float* part_tr = c->part + (size_t)B * m->sk_nslot * 12;
            k_skin_bwd_act<<<dim3(m->sk_ntile, B), 256, 0, st>>>(ps->A, m->sk_aoff, m->sk_aj, m->sk_w, c->VP, c->Gv, V, B, m->sk_nslot, c->DVP,
                                                                 partc, part_tr);
            k_skin_bwd_act_reduce<<<cdiv(B * SKB_PART, 256), 256, 0, st>>>(partc, m->sk_joff, m->sk_jslot, m->sk_nslot, part_tr, m->sk_ntile, B,
                                                                           ps->dA, ps->dtr);
        } else if (V <= SKS_V && g_skin_small) {
            static bool cfg = false;
            if (!cfg) { LEMO_CUDA(cudaFuncSetAttribute(k_skin_bwd_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SKS_BWD_SMEM)); cfg = true; }
            k_skin_bwd_small<<<std::min(B, 148 * 2), 256, SKS_BWD_SMEM, st>>>(ps->A, m->w_jm, c->VP, c->Gv, V, B, c->DVP, ps->dA, ps->dtr);
        } else {
            k_skin_bwd<<<dim3(ctas, B), 256, 0, st>>>(ps->A, m->w_jm, c->VP, c->Gv, V, B, tiles, c->DVP, ps->dA, ps->dtr, part);
            if (part) k_skin_bwd_reduce<<<cdiv(B * SKB_PART, 256), 256, 0, st>>>(part, ctas, B, ps->dA, ps->dtr);
        }
    }
    LEMO_CUDA(cudaGetLastError());
    // dX[B,512] += DVP[B,3V] . Wt^T        (contraction over 3V: split-K with atomics)
    {
        GemmP g{};
        g.A = c->DVP; g.B = m->Wt; g.C = ps->dX; g.bias = nullptr;
        g.M = B; g.N = XK; g.K = 3 * V;
        g.sAm = 3 * V; g.sAk = 1; g.sBk = 1; g.sBn = 3 * V; g.sCm = XK; g.sCn = 1;
        g.splitk = 1; g.nz = dx_slices(V);                                 // loss-row sub-models: one slice => deterministic
        if (g.nz > 1 && c->part && (size_t)g.nz * B * XK <= c->part_floats) {
            // full mesh: every K slice parks its partial product, the slices are added in order (no float atomics)
            g.splitk = 2; g.C = c->part;
            if (B <= 128 && g_dx_tallk) {
                const int kslice = cdiv(cdiv(3 * V, g.nz), DXK) * DXK;
                static bool dx_cfg = false;
                if (!dx_cfg) { LEMO_CUDA(cudaFuncSetAttribute(k_dx_tallk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DX_SMEM)); dx_cfg = true; }
                k_dx_tallk<<<dim3(XK / 128, g.nz), 256, DX_SMEM, st>>>(c->DVP, m->Wt, B, 3 * V, kslice, c->part);
            } else LEMO_TRY(gemm_launch(g, st));
            k_slices_reduce<<<cdiv((long long)B * XK, 256), 256, 0, st>>>(c->part, g.nz, (long long)B * XK, ps->dX);
            LEMO_CUDA(cudaGetLastError());
        } else LEMO_TRY(gemm_launch(g, st));
    }
    return 0;
}

__global__ void k_copy(const float* __restrict__ s, float* __restrict__ d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = s[i];
}

int body_pose_backward(BodyCtx* c, const PoseIn& in, int B, const PoseGrad& g, cudaStream_t st) {
    LEMO_CHECK(c && c->dA, "body handle was created without backward buffers");
    const Model* m = c->m;
    if (g.betas && in.betas_stride == 0) LEMO_CUDA(cudaMemsetAsync(g.betas, 0, 10 * sizeof(float), st));
    k_chain_bwd<<<B, 64, 0, st>>>(c->R, c->G, c->Jrest, c->dA, c->dJp, c->dX, m->J_dirs, m->tree, m->max_depth, B,
                                   in.betas_stride, c->dR, g.betas, g.expression);
    k_pose_to_rot_bwd<<<B, 64, 0, st>>>(make_posek(m, in), g, B, c->full_pose, c->dR);
    if (g.transl) k_copy<<<cdiv(B * 3, 128), 128, 0, st>>>(c->dtr, g.transl, B * 3);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

int gather_rows(const float* src, const int* idx, int B, int V, int n, float* out, cudaStream_t st) {
    k_gather_rows<<<cdiv((long long)B * n * 3, 256), 256, 0, st>>>(src, idx, B, V, n, out);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}
int scatter_rows_add(const float* g, const int* idx, int B, int V, int n, float* dst, cudaStream_t st) {
    k_scatter_rows_add<<<cdiv((long long)B * n * 3, 256), 256, 0, st>>>(g, idx, B, V, n, dst);
    LEMO_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace lemo
