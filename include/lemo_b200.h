/* lemo_b200 -- C ABI of the B200-native temporal body-fitting engine (liblemo_b200.so).
 *
 * This is the drop-in boundary for the fitting hot path of sanweiliti/LEMO.  The reference has no FFI of its
 * own: its "operator API" is nn.Module.__call__ + autograd.  Each entry point below
 * names the reference call it replaces (path:line under the reference tree).  The Python mirror of the
 * reference modules (lemo_b200/smplx.py, models/, temp_prox/, utils/, fit.py) binds exactly these symbols
 * through ctypes; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - plain C types only; no torch types.  `stream` is a cudaStream_t passed as void*.
 *   - every tensor pointer is a DEVICE pointer to contiguous row-major fp32 (int32 for indices) unless the
 *     parameter name starts with h_ (host).  The caller owns all tensors; the library owns only opaque
 *     handles and the scratch allocated when a handle is created.  Nothing allocates, synchronises or touches
 *     the default stream on the hot path.
 *   - all functions return 0 on success; otherwise lemo_last_error() holds a thread-local message.  Nothing
 *     throws across the ABI.  A handle is used by one host thread at a time; distinct handles are independent.
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels on the handle's device.
 */
#ifndef LEMO_B200_H
#define LEMO_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LEMO_NUM_JOINTS 55
#define LEMO_NUM_OUT_JOINTS 127 /* 55 + 21 vertex joints + 51 landmarks (smplx SMPLX.forward) */

const char* lemo_last_error(void);
int lemo_version(void);
/* debugging aid: cudaDeviceSynchronize + cudaGetLastError (clears the sticky state); 0 = clean, else lemo_last_error() has the text */
int lemo_debug_check(void);

/* ---------------------------------------------------------------- body model (smplx.create) ------------ */
typedef struct LemoModel LemoModel;   /* immutable model tensors on one device        */
typedef struct LemoBody LemoBody;     /* forward/backward state for a fixed max batch  */

/* Host-side description of an SMPL-X model file (smplx==0.1.26 .npz keys). */
typedef struct LemoModelDescC {
    int32_t n_verts;                 /* 10475 */
    int32_t n_faces;                 /* 20908 */
    int32_t num_pca_comps;           /* 12    */
    int32_t n_extra_joints;          /* 21    */
    int32_t n_landmarks;             /* 51    */
    const float* h_v_template;       /* [V,3]            */
    const float* h_shapedirs;        /* [V,3,20] betas(10) then expression(10) */
    const float* h_posedirs;         /* [486, 3V]  element (p, 3v+k)   (body_model.py:126-128) */
    const float* h_J_regressor;      /* [55,V]           */
    const float* h_lbs_weights;      /* [V,55]           */
    const int32_t* h_parents;        /* [55], parents[0] = -1 */
    const float* h_hand_comp_l;      /* [num_pca_comps,45] */
    const float* h_hand_comp_r;
    const float* h_pose_mean;        /* [165]  (hands_mean in the two hand slots when flat_hand_mean=False) */
    const int32_t* h_extra_joint_vids; /* [n_extra_joints] */
    const int32_t* h_faces;          /* [F,3]            */
    const int32_t* h_lmk_faces_idx;  /* [n_landmarks]    */
    const float* h_lmk_bary;         /* [n_landmarks,3]  */
} LemoModelDescC;

/* replaces smplx.create(...)            (opt_amass_temp.py:73-87, temp_prox/main_slide.py:160-179) */
int lemo_model_create(const LemoModelDescC* desc, int device, LemoModel** out);
/* compact sub-model holding only the given vertex rows (the loss rows of a fit: markers + foot verts) */
int lemo_model_select_rows(const LemoModel* model, const int32_t* h_rows, int32_t n_rows, LemoModel** out);
int lemo_model_destroy(LemoModel* model);
int lemo_model_num_verts(const LemoModel* model);

/* Pose inputs of body_model(**params) (utils/utils.py:141-152).  Null pointers mean "module default" (zeros). */
typedef struct LemoPoseC {
    const float* transl;          /* [B,3]  */
    const float* global_orient;   /* [B,3]  axis-angle */
    const float* body_pose;       /* [B,63] axis-angle */
    const float* jaw_pose;        /* [B,3]  */
    const float* leye_pose;       /* [B,3]  */
    const float* reye_pose;       /* [B,3]  */
    const float* left_hand_pose;  /* [B,num_pca_comps] (use_pca) or [B,45] */
    const float* right_hand_pose;
    const float* betas;           /* [B,10] or [1,10] when betas_shared */
    const float* expression;      /* [B,10] */
    const float* R_global;        /* [B,9]    optional rotation-matrix override of global_orient */
    const float* R_body;          /* [B,21,9] optional rotation-matrix override of body_pose     */
    int32_t betas_shared;         /* 1: one betas row for the whole batch */
    int32_t use_pca;              /* 1: hand poses are PCA coefficients   */
} LemoPoseC;

typedef struct LemoPoseGradC {    /* outputs of backward; any may be NULL */
    float* transl; float* global_orient; float* body_pose; float* jaw_pose; float* leye_pose; float* reye_pose;
    float* left_hand_pose; float* right_hand_pose; float* betas; float* expression; float* R_global; float* R_body;
} LemoPoseGradC;

int lemo_body_create(const LemoModel* model, int32_t max_batch, int32_t with_backward, LemoBody** out);
int lemo_body_destroy(LemoBody* body);

/* replaces SMPLX.forward / lbs()       (human_body_prior/body_model/lbs.py:34-119; smplx call sites
 * opt_amass_perframe.py:335, opt_amass_temp.py:357,364, fitting_temp_slide.py:248-258).
 * verts [B,V,3]; joints [B,127,3] (NULL for sub-models / when not needed); full_pose [B,165] (nullable). */
int lemo_smplx_forward(LemoBody* body, const LemoPoseC* pose, int32_t B,
                       float* verts, float* joints, float* full_pose, void* stream);
/* adjoint of lemo_smplx_forward for the same (body, pose, B): d_verts [B,V,3] and/or d_joints [B,127,3]. */
int lemo_smplx_backward(LemoBody* body, const LemoPoseC* pose, int32_t B,
                        const float* d_verts, const float* d_joints, const LemoPoseGradC* grads, void* stream);

/* A/B switch for the blend-shape GEMM: 1 = tcgen05 TF32 kernel (default), 0 = CUDA-core fp32 GEMM.  Debug/measurement only. */
int lemo_debug_set_blend_tc(int32_t on);
/* A/B switch for full-mesh skinning (lbs.py:106-117): 1 = tcgen05 TF32 3-term GEMM with the 3x4 apply as its epilogue (default),
   0 = CUDA-core kernel.  Debug/measurement only. */
int lemo_debug_set_skin_tc(int32_t on);
/* 1 (default): full models whose skinning weights are sparse (< 25 % non-zero, like the real SMPL-X) run the adjoint over the non-zeros;
   0: always the dense 55-wide adjoint.  Both are deterministic; they differ by summation order only. */
int lemo_debug_set_skin_sparse(int32_t on);

/* replaces verts[:, ids, :] gathers     (opt_amass_temp.py:359,366,416-425; bit-exact integer indexing) */
int lemo_gather_rows(const float* src, const int32_t* idx, int32_t B, int32_t V, int32_t n, float* out, void* stream);
int lemo_scatter_rows_add(const float* g_rows, const int32_t* idx, int32_t B, int32_t V, int32_t n, float* g_dense, void* stream);

/* ---------------------------------------------------------------- rotation conversions (utils/utils.py) - */
int lemo_rot6d_to_rotmat(const float* x6, int32_t n, float* R, void* stream);                /* utils.py:64-70  */
int lemo_rot6d_to_rotmat_backward(const float* x6, const float* dR, int32_t n, float* dx6, void* stream);
int lemo_rotmat_to_aa(const float* R, int32_t n, float* aa, void* stream);                   /* utils.py:74-81 (tgm) */
/* adjoint of lemo_rotmat_to_aa through the selected quaternion branch (what autograd does through tgm): daa [n,3] -> dR [n,9];
 * makes vposer.decode(Z,'aa') / convert_to_3D_rot differentiable like the reference's graph (utils.py:148, opt_amass_temp.py:356). */
int lemo_rotmat_to_aa_backward(const float* R, const float* daa, int32_t n, float* dR, void* stream);
int lemo_aa_to_rot6d(const float* aa, int32_t n, float* x6, void* stream);                   /* utils.py:127-130 (tgm) */
int lemo_rodrigues(const float* aa, int32_t n, float* R, void* stream);                      /* lbs.py:166-193  */
int lemo_rodrigues_backward(const float* aa, const float* dR, int32_t n, float* daa, void* stream);

/* ---------------------------------------------------------------- VPoser decoder ---------------------- */
typedef struct LemoVPoser LemoVPoser;
/* weights: nn.Linear layout [out,in]; replaces load_vposer(...).decode (vposer_smpl.py:107-121) */
int lemo_vposer_create(const float* h_fc1_w, const float* h_fc1_b, const float* h_fc2_w, const float* h_fc2_b,
                       const float* h_out_w, const float* h_out_b, int32_t max_batch, int device, LemoVPoser** out);
int lemo_vposer_destroy(LemoVPoser* vp);
/* z [B,32] -> R_body [B,21,9] (output_type 'matrot') and, if aa != NULL, [B,63] axis-angle (tgm, 'aa') */
int lemo_vposer_decode(LemoVPoser* vp, const float* z, int32_t B, float* R_body, float* aa, void* stream);
int lemo_vposer_decode_backward(LemoVPoser* vp, const float* z, int32_t B, const float* dR_body, float* dz, void* stream);

/* ---------------------------------------------------------------- motion priors (models/AE_sep.py, AE.py) */
typedef struct LemoConvNet LemoConvNet;
/* kind 0: Enc(downsample=False, z_channel=64)  (models/AE_sep.py:77-99) -- 10 conv3x3 + LeakyReLU(0.2)
 * kind 1: AE(downsample=True, in_channel=C, kernel=3) (models/AE.py:79-108)
 * h_weights: concatenation of the state_dict tensors in the reference's key order
 * (enc_blc{1-5}.main.{0,2}.{weight,bias} [, dec_blc{1-5}.deconv{1,2}.{weight,bias}]). */
int lemo_convnet_create(int32_t kind, int32_t in_channels, const float* h_weights, int64_t n_weights,
                        int32_t max_n, int32_t H, int32_t W, int32_t with_backward, int device, LemoConvNet** out);
int lemo_convnet_destroy(LemoConvNet* net);
int lemo_convnet_set_weights(LemoConvNet* net, const float* weights_dev, void* stream); /* device, same order  */
int lemo_convnet_get_weights(LemoConvNet* net, float* weights_dev, void* stream);
int64_t lemo_convnet_num_weights(const LemoConvNet* net);
/* Enc.forward: x [N,1,H,W] -> z [N,64,H,W]                                  (opt_amass_temp.py:389) */
int lemo_enc_forward(LemoConvNet* net, const float* x, int32_t N, float* z, void* stream);
/* input gradient of the last lemo_enc_forward: dz [N,64,H,W] -> dx [N,1,H,W] */
int lemo_enc_backward_input(LemoConvNet* net, const float* dz, int32_t N, float* dx, void* stream);
/* measurement hook: relaunch ONE conv layer of an Enc handle `reps` times on its resident activations
 * (forward: layer l of 0..9; backward: the input-gradient conv of layer l >= 1) so bench.py can time the
 * dominant kernel alone with CUDA events. */
/* A/B switch for the Enc conv stack.  0 = fp32 CUDA-core kernels, 1 = tcgen05 pair kernel (3-term bf16 split), 8192 = tcgen05
 * weights-in-TMEM kernel (4-term bf16 split), -1 = back to the default (environment LEMO_CONV=simt|pair|wt, else the built-in default).
 * Other values select measured experiment variants (csrc/conv_tc.cu).  Debug/measurement only. */
int lemo_debug_set_conv_tc(int32_t on);
int lemo_enc_debug_backward(LemoConvNet* net, const float* dz, int32_t N, int32_t stop_layer, float* out, void* stream);
int lemo_convnet_profile_layer(LemoConvNet* net, int32_t layer, int32_t N, int32_t backward, int32_t reps, void* stream);
/* AE.forward: x [N,C,H,W] -> rec [N,1,H,W], z [N,256,h5,w5] (nullable)      (opt_amass_perframe.py:160) */
int lemo_ae_forward(LemoConvNet* net, const float* x, int32_t N, float* rec, float* z, void* stream);
/* weight gradients of the last lemo_ae_forward: d_rec [N,1,H,W] -> d_weights (same order as weights) */
int lemo_ae_backward_weights(LemoConvNet* net, const float* d_rec, int32_t N, float* d_weights, void* stream);

/* one step of the self-supervised fine-tune (opt_amass_perframe.py:152-173): AE forward on x [N,C,H,W] (already masked and
 * reflect-padded), loss = mean |rec[:,0] - x[:,0]| over the rows with row_mask[y] != 0 (row_mask [H] floats, n_rows_selected
 * of them set), weight-gradient backward, Adam(lr, betas .9/.999, eps 1e-8) on the handle's weights.  t = 1-based step (t == 1
 * resets the moments).  loss_out: device float, nullable. */
int lemo_ae_finetune_step(LemoConvNet* net, const float* x, const float* row_mask, int32_t n_rows_selected, int32_t N, double lr,
                          int32_t t, float* loss_out, void* stream);

/* the whole fine-tune loop (opt_amass_perframe.py:152-173): `steps` x lemo_ae_finetune_step with a fresh Adam, the step captured ONCE as a
 * CUDA graph (step counter / bias corrections on the device) and replayed; losses_out: device float[steps], nullable.  Every reduction
 * (split-K convolutions of the small planes, weight gradients) has a fixed order: the fine-tuned weights are bitwise reproducible. */
int lemo_ae_finetune_run(LemoConvNet* net, const float* x, const float* row_mask, int32_t n_rows_selected, int32_t N, double lr,
                         int32_t steps, float* losses_out, void* stream);

/* ---------------------------------------------------------------- infill pre-stage --- */
/* Body representation of one clip (utils/utils.py:209-265 get_local_markers_4chan + the loader's normalisation and layout,
 * loader/optimize_loader_amass_new.py:359-361,376-377).  body [T,68,3] = pelvis joint + the 67 SSM2 markers, world z up; contact [T,4].
 * repr: float32 [4, 208, T-1] (channel, row, frame) = what the dataset hands to the AE; d_stats (device, nullable) = 420 doubles
 * {Xmean_local[208], Xstd_local[208], Xmean_global_xy, Xstd_global_xy, Xmean_global_r, Xstd_global_r}: NULL leaves the values
 * un-normalised.  rot_0_pivot: device double[1].  workspace: device, >= 8*T doubles.  Arithmetic in double like the numpy reference. */
int lemo_repr_local_markers_4chan(const float* body, const float* contact, int32_t T, const double* d_stats, float* repr,
                                  double* rot_0_pivot, double* workspace, void* stream);
/* utils/utils.py:180-203 reconstruct_global_body on its own: packed [T,70,3] = zero reference joint, local pelvis + 67 markers, one
 * trajectory row (vx, vy, r); rot_0_pivot device double[1]; out [T,68,3] world positions.  workspace: >= 8*T doubles. */
int lemo_reconstruct_global_body(const float* packed, const double* rot_0_pivot, int32_t T, float* out, double* workspace, void* stream);
/* clip [4,d,T] (d must be 208) -> x_pad [4,d+2,T+16]: upper-body marker rows and the contact rows of channel 0 zeroed, then reflect
 * padding (8,8,1,1) (opt_amass_temp.py:164-184).  row_mask (device [d+2], nullable) receives 1 on the rows of the fine-tune loss
 * (opt_amass_temp.py:196-200) for lemo_ae_finetune_step; *n_rows_selected (host, nullable) their count. */
int lemo_infill_prepare_input(const float* clip, int32_t d, int32_t T, float* x_pad, float* row_mask, int32_t* n_rows_selected,
                              void* stream);
/* AE output on the padded clip rec_pad [d+2,T+16] + the clip itself -> infilled markers in world coordinates markers_rec [T,67,3],
 * contact labels [T,4] (sigmoid > .5), and optionally the same reconstruction of the un-infilled input markers_input [T,67,3]:
 * crop, de-normalise, reconstruct_global_body (opt_amass_temp.py:205-325, utils/utils.py:180-203).  workspace: >= 8*T doubles. */
int lemo_infill_finalize(const float* rec_pad, const float* clip, const double* d_stats, const double* rot_0_pivot, int32_t d, int32_t T,
                         float* markers_rec, float* contact_lbl, float* markers_input, double* workspace, void* stream);

/* ---------------------------------------------------------------- Chamfer (temp_prox/dist_chamfer.py) --- */
/* xyz1 [B,n,3]; xyz2 [B,m,3] with xyz2_batch_stride floats between batches (0 = one shared scene).
 * dist = squared L2 to the nearest neighbour (first minimum wins), idx int32.  (dist_chamfer.py:10-28)
 * Arithmetic is pinned: d = fma(dz,dz, fma(dy,dy, dx*dx)) with dx = x1 - x2, so indices are bit-exact against
 * oracle/csrc/chamfer_ref.c, ties included.  dist2 and idx2 may both be NULL: the xyz2 -> xyz1 direction is then skipped
 * (the PROX contact term consumes dist1 only, fitting_temp_slide.py:749-753). */
int lemo_chamfer_forward(const float* xyz1, int32_t B, int32_t n, const float* xyz2, int32_t m,
                         int64_t xyz2_batch_stride, float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                         void* stream);
/* grad 2*g*(x1-x2) scattered to both clouds (dist_chamfer.py:30-45).  d_xyz2 has the same batch stride
 * as xyz2 (a shared scene accumulates over the batch).  Outputs are overwritten.  g_dist2 / idx2 may both be NULL. */
int lemo_chamfer_backward(const float* xyz1, int32_t B, int32_t n, const float* xyz2, int32_t m,
                          int64_t xyz2_batch_stride, const float* g_dist1, const float* g_dist2,
                          const int32_t* idx1, const int32_t* idx2, float* d_xyz1, float* d_xyz2, void* stream);

/* Static scene accelerator for the PROX contact term (fitting_temp_slide.py:743-753): the scene mesh of a recording is fixed
 * (fit_temp_loadprox_slide.py:366-372), so its points are k-d sorted once (median splits of the longest axis) into boxed tiles and a query scans only the tiles that
 * can still hold a closer point.  dist1 / idx1 are IDENTICAL to lemo_chamfer_forward's against the same scene (same pinned
 * arithmetic, same first-minimum rule; idx1 = index into the ORIGINAL scene array).  lemo_scene_create synchronises (create time). */
typedef struct LemoScene LemoScene;
int lemo_scene_create(const float* scene_points /* device [m,3] */, int32_t m, LemoScene** out);
int lemo_scene_destroy(LemoScene* scene);
int lemo_scene_query(const LemoScene* scene, const float* xyz1 /* [B,n,3] */, int32_t B, int32_t n, float* dist1, int32_t* idx1, void* stream);

/* ---------------------------------------------------------------- PROX scene terms (temp_prox/fitting_temp_slide.py) --- */
/* PerspectiveCamera.forward (temp_prox/camera.py:93-116): img = f * (R p + t).xy / (R p + t).z + c.  points [n,3] -> out [n,2].
 * h_R [9] / h_t [3] are HOST arrays (the camera is fixed in the shipped configs, S2.yaml camera_mode 'fixed'); NULL = identity / zero. */
int lemo_camera_project(const float* points, int64_t n, const float* h_R, const float* h_t, float fx, float fy, float cx, float cy,
                        float* out, void* stream);
int lemo_camera_project_backward(const float* points, int64_t n, const float* h_R, const float* h_t, float fx, float fy, float cx,
                                 float cy, const float* d_out, float* d_points, void* stream);
/* camera -> world: out = R p + t (fitting_temp_slide.py:677-678); adjoint = 1 computes R^T g (the gradient) */
int lemo_rigid_transform(const float* points, int64_t n, const float* h_R, const float* h_t, int32_t adjoint, float* out, void* stream);
/* F.grid_sample(sdf, norm_vertices[:, :, [2,1,0]], padding_mode='border') (fitting_temp_slide.py:684-687): trilinear lookup of a
 * [dim,dim,dim] signed-distance volume (indexed [x][y][z], shared by the batch) at world points [n,3]; align_corners=False. */
int lemo_sdf_sample(const float* points, int64_t n, const float* sdf, int32_t dim, const float* h_grid_min, const float* h_grid_max,
                    float* values, void* stream);
int lemo_sdf_sample_backward(const float* points, int64_t n, const float* sdf, int32_t dim, const float* h_grid_min,
                             const float* h_grid_max, const float* d_values, float* d_points, void* stream);

/* ---------------------------------------------------------------- optimiser (torch.optim.Adam.step) ----- */
/* p,g,m,v [n]; t = 1-based step; bias-corrected, no weight decay / amsgrad (opt_amass_temp.py:345,455) */
int lemo_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                   double eps, int32_t t, void* stream);

/* ---------------------------------------------------------------- fused fitting drivers ----------------- */
typedef struct LemoFit LemoFit;
typedef struct LemoFitConfigC {
    int32_t mode;               /* 0 temporal (opt_amass_temp.py:329-455), 1 per-frame (opt_amass_perframe.py:293-361) */
    int32_t n_seq;              /* S sequences fitted side by side on this device          */
    int32_t n_frames;           /* T frames per sequence                                   */
    float w_rec, w_vposer, w_shape, w_hand, w_contact, w_smooth;  /* loss weights (argparse defaults of the scripts) */
    float vel_thres;            /* 0.1  (opt_amass_temp.py:428) */
    float fps;                  /* 30                            */
    const int32_t* h_markers67; /* [67]  loader/SSM2.json                                   */
    const int32_t* h_markers81; /* [81]  loader/SSM2_withhand.json                          */
    const int32_t* h_foot_ids[4];   /* left_heel, right_heel, left_toe, right_toe (opt_amass_temp.py:97-113) */
    int32_t n_foot[4];
    const float* h_smooth_mean; /* [243] */
    const float* h_smooth_std;  /* [243] */
    int32_t use_cuda_graph;     /* capture one iteration and replay it */
} LemoFitConfigC;

int lemo_fit_create(const LemoModel* model, LemoVPoser* vposer_or_null, const float* h_vposer_weights_unused,
                    LemoConvNet* enc_or_null, const LemoFitConfigC* cfg, int device, LemoFit** out);
int lemo_fit_destroy(LemoFit* fit);
/* per-sequence inputs (device pointers): init params [T,72] (transl3, aa3, betas10, z32, lh12, rh12),
 * target markers [T,67,3], contact labels [T,4].  Temporal mode.  (opt_amass_temp.py:253,329-341) */
int lemo_fit_set_sequence(LemoFit* fit, int32_t s, const float* init72, const float* markers_rec, const float* contact,
                          void* stream);
/* all S sequences of the handle in one call: init72 [S,T,72], markers_rec [S,T,67,3], contact [S,T,4] (device pointers); two copies and
 * one kernel instead of 3 S launches -- the batched form of the per-sequence loads of opt_amass_temp.py:253,329-341.  Temporal mode. */
int lemo_fit_set_sequences(LemoFit* fit, const float* init72, const float* markers_rec, const float* contact, void* stream);
/* n_iters Adam iterations with the script's LR schedule (lr0 until step>lr_switch, then lr1).
 * Entirely on device: no host synchronisation inside. */
int lemo_fit_run(LemoFit* fit, int32_t n_iters, float lr0, float lr1, int32_t lr_switch, void* stream);
/* per-frame mode: runs all T frames x n_iters of sequence-parallel B=1 problems (lr .1/.01 -> .01@>60 -> .003@>80).  Default: one
 * persistent kernel launch, one 8-CTA thread-block cluster per sequence, no host involvement until the last frame is done. */
int lemo_fit_run_perframe(LemoFit* fit, int32_t n_iters, void* stream);
/* results: params72 [S,T,72] as of the LAST forward (what the scripts save), losses [S,8] of the last iteration
 * (total, rec, vposer, shape, hand, contact, smooth, reserved) */
int lemo_fit_get(LemoFit* fit, float* params72, float* losses, void* stream);
/* raw optimisation state after the last step: transl [S,T,3], rot6d [S,T,6], other [S,T,56]; grads of last iteration */
int lemo_fit_get_state(LemoFit* fit, float* transl, float* rot6d, float* other, float* g_transl, float* g_rot6d,
                       float* g_other, void* stream);
int64_t lemo_fit_kernel_launches(const LemoFit* fit);   /* kernels enqueued by this handle so far */
/* A/B switch for lemo_fit_run_perframe: 1 = persistent cluster kernel (default: ONE launch for all frames x all steps, csrc/perframe_mega.cuh),
 * 0 = one CUDA graph of ~23 kernels per step (the round-1 path), -1 = back to the default / LEMO_PERFRAME=graph.  Debug/measurement only. */
int lemo_debug_set_perframe(int32_t mode);

/* ---------------------------------------------------------------- fused PROX stage-2 driver ------------ */
/* One B-frame sliding window of temp_prox: FittingMonitor.run_fitting + create_fitting_closure.fitting_func
 * (fitting_temp_slide.py:169-313) over SMPLifyLoss.forward (:564-1062) with the terms PROXD_temp_S2.yaml selects + `contact`.
 * Loss weights carry the reference's attribute names (SMPLifyLoss.__init__ :339-387, reset_loss_weights :548-562). */
typedef struct LemoProxFit LemoProxFit;
typedef struct LemoProxWeightsC {
    float data_weight, body_pose_weight, shape_weight, bending_prior_weight, hand_prior_weight, expr_prior_weight, jaw_prior_weight;
    float sdf_penetration_weight, contact_loss_weight, motion_prior_smooth_weight, friction_normal_weight, friction_tangent_weight;
    int32_t use_joints_conf;          /* weights = joint_weights * joints_conf (:577-579) */
} LemoProxWeightsC;
typedef struct LemoProxConfigC {
    int32_t n_frames;                 /* B: frames of the window (100 in the shipped configs)                          */
    int32_t n_joints_mapped;          /* Jm: keypoints after the OpenPose JointMapper (118)                            */
    const int32_t* h_joint_map;       /* [Jm] indices into the 127 model joints (misc_utils.smpl_to_openpose); NULL = identity */
    float cam_R[9], cam_t[3], fx, fy, cx, cy;     /* fixed PerspectiveCamera (camera.py:93-116)                        */
    float R[9], t[3];                 /* cam2world (fit_temp_loadprox_slide.py:307-310)                                 */
    const float* sdf;                 /* DEVICE [dim,dim,dim] signed distances, indexed [x][y][z], caller-owned; NULL = no scene SDF */
    int32_t sdf_dim;
    float grid_min[3], grid_max[3];
    int32_t sdf_penetration, use_friction, contact, use_motion_smooth_prior;      /* term switches (S2.yaml)            */
    const int32_t* h_fric_ids; int32_t n_fric;            /* contact_fric_verts_ids (fit_temp_loadprox_slide.py:349-354) */
    const int32_t* h_contact_ids; int32_t n_contact;      /* contact_verts_ids (:356-362)                                */
    const int32_t* h_markers81;       /* smooth_marker_ids (loader/SSM2_withhand.json)                                  */
    const float* scene_v; int32_t n_scene;                /* DEVICE [m,3] scene vertices, caller-owned                   */
    const float* h_smooth_mean; const float* h_smooth_std;    /* [243] preprocess_stats_smooth_withHand_global_markers   */
    LemoProxWeightsC weights;
    int32_t use_cuda_graph;
} LemoProxConfigC;
typedef struct LemoProxWindowC {      /* device pointers, [B, .]; NULL = zeros (joints_conf NULL = ones) */
    const float* transl; const float* global_orient; const float* pose_embedding; const float* left_hand_pose; const float* right_hand_pose;
    const float* jaw_pose; const float* leye_pose; const float* reye_pose; const float* expression; const float* betas;
    const float* gt_joints;           /* [B,Jm,2] */
    const float* joints_conf;         /* [B,Jm]   */
    const float* joint_weights;       /* [B,Jm]   */
} LemoProxWindowC;
typedef struct LemoProxParamsOutC {   /* device pointers, any may be NULL */
    float* transl; float* global_orient; float* pose_embedding; float* left_hand_pose; float* right_hand_pose;
    float* jaw_pose; float* leye_pose; float* reye_pose; float* expression;
} LemoProxParamsOutC;
int lemo_fit_prox_create(const LemoModel* model, LemoVPoser* vposer, LemoConvNet* enc_or_null, const LemoProxConfigC* cfg, int device,
                         LemoProxFit** out);
int lemo_fit_prox_destroy(LemoProxFit* fit);
/* loss.reset_loss_weights(curr_weights) per optimisation stage + the closure's gradient erase: frames [0, erase_n) are frozen
 * (erase_n = int(bs*0.15) when first_batch_flag is False, else 0; fitting_temp_slide.py:281-288). */
int lemo_fit_prox_set_weights(LemoProxFit* fit, const LemoProxWeightsC* w, int32_t erase_n, void* stream);
int lemo_fit_prox_set_window(LemoProxFit* fit, const LemoProxWindowC* window, void* stream);
/* n_iters closure steps with a fresh Adam(lr, betas .9/.999, eps 1e-8) (optim_factory.py:77-80), no host synchronisation.
 * resume = 1 continues the previous call's moments / step count / lr (a run split into chunks to read the loss in between). */
int lemo_fit_prox_run(LemoProxFit* fit, int32_t n_iters, float lr, int32_t resume, void* stream);
/* one closure evaluation (loss terms + gradients after the erase) without an optimiser step: the closure-level parity hook */
int lemo_fit_prox_eval(LemoProxFit* fit, void* stream);
/* current parameters / gradients of the last closure / losses16 = {joint, pprior, shape, angle, hand(l+r), expr, jaw, sdf_penetration,
 * fric_tangent, fric_normal, contact, motion_prior_smooth, 0, 0, 0, total} of the last closure */
int lemo_fit_prox_get(LemoProxFit* fit, const LemoProxParamsOutC* params, const LemoProxParamsOutC* grads, float* losses16, void* stream);
int64_t lemo_fit_prox_kernel_launches(const LemoProxFit* fit);

/* ---------------------------------------------------------------- test hooks (host math, no GPU) -------- */
/* Host-side builder of the kinematic-tree tables the chain kernels walk (csrc/body.cuh TREE_*; replaces the Python loop over `parents` of
   lbs.py:246-251 `batch_rigid_transform`).  parents: 55 ints, parents[j] < j, parents[0] ignored (root).  tables: >= 776 ints:
     [0,55) parent (-1 root) | [56,111) joints ordered by (depth, index) | [112,..) first position of each level in that order |
     [128,184) children of joint j = list[koff[j] .. koff[j+1]) | [184,238) that list, ascending per parent |
     [240,720) one word per (level, lane < 32): joint | (parent+1) << 8 | koff << 16 | n_children << 24, or -1 | [720,775) depth.
   Returns 0, 1 (bad arguments), 11 (parents[j] >= j), 12 (deeper than 14 levels), 13 (more than 32 joints on a level). */
int lemo_host_tree_tables(const int32_t* parents, int32_t* tables, int32_t n_tables, int32_t* max_depth);
void lemo_host_rodrigues(const float* aa, float* R);
void lemo_host_rodrigues_bwd(const float* aa, const float* dR, float* daa);
void lemo_host_gs6d(const float* x6, float* R);
void lemo_host_gs6d_bwd(const float* x6, const float* dR, float* dx6);
void lemo_host_rotmat_to_aa(const float* R, float* aa);
void lemo_host_rotmat_to_aa_bwd(const float* R, const float* daa, float* dR);
void lemo_host_aa_to_rotmat_tgm(const float* aa, float* R);

#ifdef __cplusplus
}
#endif
#endif
