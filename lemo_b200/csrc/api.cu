// extern "C" entry points for the body model, gathers and the VPoser decoder (see include/lemo_b200.h).
#include "handles.cuh"
#include "../../include/lemo_b200.h"

using namespace lemo;
#define ST(s) ((cudaStream_t)(s))

static PoseIn to_posein(const LemoPoseC* p) {
    PoseIn in;
    in.transl = p->transl; in.global_orient = p->global_orient; in.body_pose = p->body_pose;
    in.jaw = p->jaw_pose; in.leye = p->leye_pose; in.reye = p->reye_pose;
    in.lhand = p->left_hand_pose; in.rhand = p->right_hand_pose;
    in.betas = p->betas; in.expression = p->expression;
    in.R_global = p->R_global; in.R_body = p->R_body;
    in.betas_stride = p->betas_shared ? 0 : 10;
    in.hand_is_pca = p->use_pca ? 1 : 0;
    return in;
}

extern "C" {

int lemo_model_create(const LemoModelDescC* desc, int device, LemoModel** out) {
    LEMO_CHECK(desc && out, "null argument");
    LEMO_CHECK(desc->h_v_template && desc->h_shapedirs && desc->h_posedirs && desc->h_J_regressor && desc->h_lbs_weights &&
               desc->h_parents && desc->h_hand_comp_l && desc->h_hand_comp_r && desc->h_pose_mean, "model tensor missing");
    Model* m = nullptr;
    LEMO_TRY(model_create_from_host(desc, device, &m));
    *out = new LemoModel{m};
    return 0;
}
int lemo_model_select_rows(const LemoModel* model, const int32_t* h_rows, int32_t n_rows, LemoModel** out) {
    LEMO_CHECK(model && out, "null argument");
    Model* s = nullptr;
    LEMO_TRY(model_select_rows(model->m, h_rows, n_rows, &s));
    *out = new LemoModel{s};
    return 0;
}
int lemo_model_destroy(LemoModel* model) {
    if (!model) return 0;
    model_free(model->m);
    delete model;
    return 0;
}
int lemo_model_num_verts(const LemoModel* model) { return model ? model->m->V : 0; }

int lemo_body_create(const LemoModel* model, int32_t max_batch, int32_t with_backward, LemoBody** out) {
    LEMO_CHECK(model && out, "null argument");
    BodyCtx* c = nullptr;
    LEMO_TRY(bodyctx_create(model->m, max_batch, with_backward != 0, &c));
    *out = new LemoBody{c};
    return 0;
}
int lemo_body_destroy(LemoBody* body) {
    if (!body) return 0;
    bodyctx_free(body->c);
    delete body;
    return 0;
}

int lemo_smplx_forward(LemoBody* body, const LemoPoseC* pose, int32_t B, float* verts, float* joints, float* full_pose, void* stream) {
    LEMO_NVTX("lemo_smplx_forward");
    LEMO_CHECK(body && pose && verts, "null argument");
    const PoseIn in = to_posein(pose);
    LEMO_CHECK(in.betas_stride != 0 || in.betas, "betas_shared needs a betas pointer");
    LEMO_TRY(body_pose_forward(body->c, in, B, ST(stream)));
    LEMO_TRY(body_skin_forward(body->c, body->c, in, B, verts, joints, ST(stream)));
    if (full_pose)
        LEMO_CUDA(cudaMemcpyAsync(full_pose, body->c->full_pose, (size_t)B * 165 * sizeof(float), cudaMemcpyDeviceToDevice, ST(stream)));
    return 0;
}

int lemo_smplx_backward(LemoBody* body, const LemoPoseC* pose, int32_t B, const float* d_verts, const float* d_joints,
                        const LemoPoseGradC* grads, void* stream) {
    LEMO_NVTX("lemo_smplx_backward");
    LEMO_CHECK(body && pose && grads, "null argument");
    LEMO_CHECK(B > 0 && B <= body->c->maxB, "batch exceeds the size this body handle was created for");
    const PoseIn in = to_posein(pose);
    PoseGrad g;
    g.transl = grads->transl; g.global_orient = grads->global_orient; g.body_pose = grads->body_pose; g.jaw = grads->jaw_pose;
    g.leye = grads->leye_pose; g.reye = grads->reye_pose; g.lhand = grads->left_hand_pose; g.rhand = grads->right_hand_pose;
    g.betas = grads->betas; g.expression = grads->expression; g.R_global = grads->R_global; g.R_body = grads->R_body;
    LEMO_TRY(body_grad_begin(body->c, B, ST(stream)));
    LEMO_TRY(body_skin_backward(body->c, body->c, B, d_verts, d_joints, ST(stream)));
    LEMO_TRY(body_pose_backward(body->c, in, B, g, ST(stream)));
    return 0;
}

int lemo_debug_set_blend_tc(int32_t on) { blend_tc_set(on); return 0; }
int lemo_debug_set_skin_tc(int32_t on) { skin_tc_set(on); return 0; }
int lemo_debug_set_skin_sparse(int32_t on) { skin_sparse_set(on); return 0; }
int lemo_host_tree_tables(const int32_t* parents, int32_t* tables, int32_t n_tables, int32_t* max_depth) {
    if (!parents || !tables || !max_depth || n_tables < TREE_N) return 1;
    int md = 0;
    const int rc = build_tree_tables(parents, tables, &md);
    *max_depth = md;
    return rc ? 10 + rc : 0;
}

int lemo_gather_rows(const float* src, const int32_t* idx, int32_t B, int32_t V, int32_t n, float* out, void* stream) {
    LEMO_CHECK(src && idx && out, "null argument");
    return gather_rows(src, idx, B, V, n, out, ST(stream));
}
int lemo_scatter_rows_add(const float* g_rows, const int32_t* idx, int32_t B, int32_t V, int32_t n, float* g_dense, void* stream) {
    LEMO_CHECK(g_rows && idx && g_dense, "null argument");
    return scatter_rows_add(g_rows, idx, B, V, n, g_dense, ST(stream));
}

int lemo_vposer_create(const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                       int32_t max_batch, int device, LemoVPoser** out) {
    LEMO_CHECK(out, "null out");
    VPoser* v = nullptr;
    LEMO_TRY(vposer_create(w1, b1, w2, b2, w3, b3, max_batch, device, &v));
    *out = new LemoVPoser{v};
    return 0;
}
int lemo_vposer_destroy(LemoVPoser* vp) {
    if (!vp) return 0;
    vposer_free(vp->v);
    delete vp;
    return 0;
}
int lemo_vposer_decode(LemoVPoser* vp, const float* z, int32_t B, float* R_body, float* aa, void* stream) {
    LEMO_NVTX("lemo_vposer_decode");
    LEMO_CHECK(vp, "null handle");
    return vposer_decode(vp->v, z, B, R_body, aa, ST(stream));
}
int lemo_vposer_decode_backward(LemoVPoser* vp, const float* z, int32_t B, const float* dR_body, float* dz, void* stream) {
    LEMO_NVTX("lemo_vposer_decode_backward");
    LEMO_CHECK(vp, "null handle");
    return vposer_decode_backward(vp->v, z, B, dR_body, dz, ST(stream));
}
}
