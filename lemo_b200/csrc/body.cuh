// SMPL-X body-model path: internal interfaces shared by api.cu and fit.cu.
#pragma once
#include "common.cuh"
#include "../../include/lemo_b200.h"

namespace lemo {

// Kinematic-tree tables (ints, built once per model on the host): the chain walks are done by ONE warp, level by level, lane i taking the
// i-th joint of the level, with __syncwarp() between levels instead of CTA barriers.
constexpr int TREE_PAR = 0;        // [55] parent of joint j (-1 for the root)
constexpr int TREE_ORDER = 56;     // [55] joints sorted by (depth, index)
constexpr int TREE_OFF = 112;      // [max_depth + 2] first position of each level in ORDER
constexpr int TREE_KOFF = 128;     // [56] children of joint j = KLIST[KOFF[j] .. KOFF[j+1]) in ascending index
constexpr int TREE_KLIST = 184;    // [54]
constexpr int TREE_LANE = 240;     // [TREE_MAX_DEPTH + 1][32] one word per (level, lane): joint | (parent + 1) << 8 | KOFF << 16 | n_children << 24, or -1
constexpr int TREE_DEPTH = 240 + 15 * 32;   // [55] depth of joint j (the stand-alone forward kernel walks thread-per-joint)
constexpr int TREE_N = TREE_DEPTH + 56;
constexpr int TREE_MAX_DEPTH = 14;

constexpr int SKIN_TC_FR = 8;      // frames per unit of the tensor-core skinning kernel (layout of BodyCtx::A2)

// Immutable device-resident model (created once per gender per device).
struct Model {
    int device = 0;
    int V = 0;                 // vertices of THIS model (10475, or n rows for a loss-row sub-model)
    int npc = 12;              // hand PCA components
    int n_extra = 0, n_lmk = 0;
    int max_depth = 0;
    bool is_sub = false;
    float* v_template = nullptr;   // [V,3]
    float* Wt = nullptr;           // [512, 3V]  rows 0..485 posedirs, 486..505 shapedirs^T, 506..511 zero
    float* WtT = nullptr;          // K-major copy of Wt for the tcgen05 blend GEMM, stored box by box: [3V/224][16][224][32] (blend_tc.cu)
    alignas(64) unsigned char map_w[128];   // CUtensorMap over WtT
    bool has_tc = false;
    float* W2 = nullptr;           // TF32-split skinning weights, box by box: [V/128][hi j0-31 | hi j32-63 | lo .. | lo ..][128][32]; full models only
    alignas(64) unsigned char map_w2[128];  // CUtensorMap over W2
    bool has_skin_tc = false;
    float* w_jm = nullptr;         // [55, V]    skinning weights, joint-major (lbs_weights^T)
    float* J_template = nullptr;   // [55,3]     J_regressor . v_template
    float* J_dirs = nullptr;       // [55,3,20]  J_regressor . shapedirs
    int* parents = nullptr;        // [55] device
    int* depth = nullptr;          // [55] device
    int* tree = nullptr;           // [TREE_N] device: level-ordered tree tables (see TREE_*)
    float* hand_l = nullptr;       // [npc,45]
    float* hand_r = nullptr;
    float* pose_mean = nullptr;    // [165]
    int* extra_vids = nullptr;     // [n_extra]
    int* lmk_tri = nullptr;        // [n_lmk,3] vertex ids
    float* lmk_bary = nullptr;     // [n_lmk,3]
    // adjoint of the vertex-derived output joints as a gather: distinct vertex u receives sum_e jv_w[e] * d_joints[jv_q[e]] over its entries
    // [jv_off[u], jv_off[u+1]) in output-joint order (fixed order, single writer: no atomics)
    int n_jv = 0;
    int *jv_vid = nullptr, *jv_off = nullptr, *jv_q = nullptr;
    float* jv_w = nullptr;
    // compact view of the skinning weights for the full-mesh adjoint (real SMPL-X weights have a handful of influences per vertex, so a
    // tile of 256 consecutive vertices touches ~10 of the 55 joints: the dense 55-wide adjoint spends most of its FMAs on zeros).
    // Built for full models when the tiles' active-joint lists add up to < 40 % of ntile x 55:
    //   tile t owns slots [sk_aoff[t], sk_aoff[t+1]); slot s = (joint sk_aj[s], weights sk_w[s][256] of the tile's vertices, zero-filled);
    //   joint j is finished from the slots sk_jslot[sk_joff[j] .. sk_joff[j+1]) in tile order (fixed order: deterministic)
    int sk_ntile = 0, sk_nslot = 0;
    int *sk_aoff = nullptr, *sk_aj = nullptr, *sk_joff = nullptr, *sk_jslot = nullptr;
    float* sk_w = nullptr;
    int h_parents[NJ];
    int h_depth[NJ];
};


// Pointers describing one batch of pose inputs (device, fp32, contiguous rows).
struct PoseIn {
    const float* transl = nullptr;          // [B,3] or null (zeros)
    const float* global_orient = nullptr;   // [B,3] aa, or null when R_global is given
    const float* body_pose = nullptr;       // [B,63] aa, or null when R_body is given
    const float* jaw = nullptr;             // [B,3] or null
    const float* leye = nullptr;
    const float* reye = nullptr;
    const float* lhand = nullptr;           // [B,npc] PCA coeffs (hand_is_pca) or [B,45] aa
    const float* rhand = nullptr;
    const float* betas = nullptr;           // [B,10] or [1,10] (betas_stride 0)
    const float* expression = nullptr;      // [B,10] or null
    const float* R_global = nullptr;        // [B,9]   rotation-matrix override for joint 0
    const float* R_body = nullptr;          // [B,21,9] override for joints 1..21
    int betas_stride = 10;                  // floats between consecutive frames' betas (0 = shared)
    int hand_is_pca = 1;
};
struct PoseGrad {                            // all nullable; written (not accumulated) unless noted
    float* transl = nullptr;
    float* global_orient = nullptr;
    float* body_pose = nullptr;
    float* jaw = nullptr;
    float* leye = nullptr;
    float* reye = nullptr;
    float* lhand = nullptr;
    float* rhand = nullptr;
    float* betas = nullptr;                 // [B,10] (or [1,10] accumulated with atomics when betas_stride==0)
    float* expression = nullptr;            // [B,10]
    float* R_global = nullptr;              // [B,9]
    float* R_body = nullptr;                // [B,21,9]
};

// Per-batch scratch/saved state for forward+backward of one model (or sub-model).
struct BodyCtx {
    const Model* m = nullptr;
    int maxB = 0;
    int device = 0;              // copy of m->device: the model may already be gone when this context is freed (Python GC order)
    float* full_pose = nullptr;  // [B,165]
    float* R = nullptr;          // [B,55,9]
    float* X = nullptr;          // [B,512]
    float* X2 = nullptr;         // [B,1024]  TF32 split of X: [Xhi | Xlo]  (A operand of the tcgen05 blend GEMM)
    alignas(64) unsigned char map_x[128];   // CUtensorMap over X2
    float* G = nullptr;          // [B,55,12]
    float* A = nullptr;          // [B,55,12]
    float* A2 = nullptr;         // TF32 split of A^T, box by box: [B/8][hi j0-31 | hi j32-63 | lo j0-31 | lo j32-63][(b%8)*12+k][32]  (skin_tc.cu)
    alignas(64) unsigned char map_a2[128];  // CUtensorMap over A2
    float* Jrest = nullptr;      // [B,55,3]
    float* Jposed = nullptr;     // [B,55,3]  (without transl)
    float* VP = nullptr;         // [B,3V]    v_posed (saved for backward)
    // backward scratch
    float* Gv = nullptr;         // [B,V,3]   dL/dverts working copy
    float* DVP = nullptr;        // [B,3V]
    float* dA = nullptr;         // [55, B*12]  joint-major
    float* dX = nullptr;         // [B,512]
    float* dR = nullptr;         // [B,55,9]
    float* dJp = nullptr;        // [B,55,3]
    float* dtr = nullptr;        // [B,3]
    float* part = nullptr;       // full meshes: per-CTA partials of k_skin_bwd [ctas][B*663] and slice partials of the dX GEMM [nz][B*512]
    size_t part_floats = 0;
};

int model_create_from_host(const ::LemoModelDescC* d, int device, Model** out);
int model_select_rows(const Model* m, const int* rows_host, int n, Model** out);
void model_free(Model* m);

int bodyctx_create(const Model* m, int maxB, bool with_backward, BodyCtx** out);
void bodyctx_free(BodyCtx* c);

// pose -> R, chain -> X, A (shared by a full model and its sub-models: pass the same ctx pose buffers)
int body_pose_forward(BodyCtx* c, const PoseIn& in, int B, cudaStream_t st);
// blend GEMM only / everything after it: body_skin_forward = body_blend_forward + body_apply_forward
int body_blend_forward(BodyCtx* c, const BodyCtx* pose_src, int B, cudaStream_t st);
// X, A -> verts [B,V,3] (+transl) ; joints_out [B,127,3] nullable (full model only)
int body_skin_forward(BodyCtx* c, const BodyCtx* pose_src, const PoseIn& in, int B, float* verts, float* joints, cudaStream_t st);
// d_verts [B,V,3] nullable, d_joints [B,127,3] nullable -> accumulates into pose_src->dA / dX / dtr / dJp
// (call body_grad_begin first), then body_pose_backward turns those into parameter gradients.
int body_grad_begin(BodyCtx* pose_src, int B, cudaStream_t st);
int body_skin_backward(BodyCtx* c, BodyCtx* pose_src, int B, const float* d_verts, const float* d_joints, cudaStream_t st);
int body_pose_backward(BodyCtx* c, const PoseIn& in, int B, const PoseGrad& g, cudaStream_t st);

// generic TF32 tensor-core GEMM (blend_tc.cu): epilogue options
struct TcEpi {
    const float* bias = nullptr;       // [N]
    const float* mask_src = nullptr;   // act == 2: [M, ldc], factor = mask_src > 0 ? 1 : 0.2
    int act = 0;                       // 0 none, 1 LeakyReLU(0.2), 2 LeakyReLU' mask
    long long ldc = 0;                 // row pitch of C / mask_src (0 = N)
    float* split_out = nullptr;        // optional [M, split_ld]: hi at column n, lo at column split_lo + n
    long long split_ld = 0, split_lo = 0;
    int b_tiled = 0;                   // B operand stored box by box (blend_tc_map_w) instead of plain [N][K]
};
int tc_gemm_launch(const void* map_a, const void* map_b, float* C, int M, int N, int K, int lo_col, const TcEpi& ep, cudaStream_t st,
                   const void* map_b_lo = nullptr);   // map_b_lo: B - rn_tf32(B) => fp32-grade 3-term product
int tc_map_a(void* map, const float* base, long long rows, int cols);     // A operand [rows, cols], 128-row boxes
int tc_map_b(void* map, const float* base, long long rows, int cols);     // B operand [rows = N, cols = K], 224-row boxes
int tc_prep_b(const float* src, int rows_src, int cols_src, int transpose, int kpad, float* dst, cudaStream_t st, float* dst_lo = nullptr);
// tcgen05 blend GEMM (blend_tc.cu)
int blend_tc_map_x(const float* X2, int maxB, void* map_x);
int blend_tc_map_w(const float* WtT, int N, void* map_w);
int blend_tc_wtt_floats(int N);       // allocation size of the box-tiled WtT
int blend_tc_transpose(const float* Wt, float* WtT, int N);
int blend_tc_launch(const void* map_x, const void* map_w, float* VP, int B, int N, cudaStream_t st);
int blend_tc_launch_bias(const void* map_x, const void* map_w, float* VP, int B, int N, const float* bias, cudaStream_t st);
bool blend_tc_enabled();
void blend_tc_set(int on);
// tcgen05 skinning (skin_tc.cu)
int skin_tc_vpad(int V);
int skin_tc_prep_w(const float* w_jm, int V, float* W2, void* map_w2);
int skin_tc_map_a(const float* A2, int maxB, void* map_a2);
size_t skin_tc_a2_floats(int maxB);
int skin_tc_launch(const void* map_w2, const void* map_a2, const float* VP, const float* transl, int V, int B, float* verts, cudaStream_t st);
bool skin_tc_enabled();
void skin_tc_set(int on);
void skin_sparse_set(int on);
int build_tree_tables(const int* parents, int* tables /*TREE_N*/, int* max_depth_out);

int gather_rows(const float* src, const int* idx_dev, int B, int V, int n, float* out, cudaStream_t st);
int scatter_rows_add(const float* g_rows, const int* idx_dev, int B, int V, int n, float* g_dense, cudaStream_t st);

}  // namespace lemo
