"""Seeded synthetic inputs: data generation only (numpy / seeded torch RNG), no compute path and no oracle import.
Used by bench.py, the scripts' --synthetic mode and (re-exported as oracle.synth) by the tests.

The licensed SMPL-X .npz, the VPoser checkpoint and AMASS are not available, so
every config of BASELINE.json runs on SMPL-X-*shaped* random tensors.  Shapes / key names follow the
smplx==0.1.26 model file; value distributions follow SURVEY.md section 8d.
"""
import os
import numpy as np

V, J, NB, P = 10475, 55, 20, 486
N_FACES = 20908

# standard SMPL-X kinematic tree (joint names: /root/reference/utils/utils.py:269-294)
PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                    20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                    21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53], np.int32)

# smplx vertex_ids['smplx'] extra joints
EXTRA_JOINT_VIDS = np.array([9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
                             5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022], np.int32)

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(_HERE, '..', 'tests', 'golden')
ASSETS = os.path.join(_HERE, 'assets')


def make_smplx_model(seed=0, n_verts=V, weights_nnz=0):
    """Random SMPL-X-shaped model.  seed 0 = 'neutral/male', seed 1 = 'female'.
    weights_nnz>0 keeps at most that many skinning weights per vertex, on tree-adjacent joints (realistic sparsity variant)."""
    g = np.random.default_rng(seed)
    f32 = np.float32
    m = {}
    m['v_template'] = (0.3 * g.standard_normal((n_verts, 3))).astype(f32)
    m['shapedirs'] = (0.01 * g.standard_normal((n_verts, 3, NB))).astype(f32)
    # posedirs in the smplx layout after its reshape: [P, 3V], element (p, 3v+k)
    m['posedirs'] = (0.001 * g.standard_normal((P, 3 * n_verts))).astype(f32)
    jr = g.random((J, n_verts)) ** 16           # a few dominant vertices per joint
    m['J_regressor'] = (jr / jr.sum(1, keepdims=True)).astype(f32)
    w = g.random((n_verts, J))
    if weights_nnz:
        # like the real model: every vertex is bound to a few joints that are neighbours in the kinematic tree, and the vertex numbering
        # is spatially coherent (consecutive vertices belong to the same body part): home joint by vertex index, then its ancestors
        home = np.minimum((np.arange(n_verts) * J) // max(n_verts, 1), J - 1)
        keep = np.zeros((n_verts, J), bool)
        cur = home.copy()
        for _ in range(weights_nnz):
            keep[np.arange(n_verts), cur] = True
            cur = np.where(PARENTS[cur] >= 0, PARENTS[cur], cur)
        w = np.where(keep, w, 0.0)
    m['lbs_weights'] = (w / w.sum(1, keepdims=True)).astype(f32)
    m['parents'] = PARENTS.copy()
    m['hands_componentsl'] = (0.1 * g.standard_normal((45, 45))).astype(f32)
    m['hands_componentsr'] = (0.1 * g.standard_normal((45, 45))).astype(f32)
    m['hands_meanl'] = (0.1 * g.standard_normal(45)).astype(f32)
    m['hands_meanr'] = (0.1 * g.standard_normal(45)).astype(f32)
    m['faces'] = g.integers(0, n_verts, (N_FACES, 3)).astype(np.int32)
    m['lmk_faces_idx'] = g.integers(0, N_FACES, 51).astype(np.int32)
    b = g.random((51, 3))
    m['lmk_bary_coords'] = (b / b.sum(1, keepdims=True)).astype(f32)
    m['extra_joint_vids'] = np.minimum(EXTRA_JOINT_VIDS, n_verts - 1).astype(np.int32)
    return m


def make_vposer_weights(seed=1):
    """VPoser decoder 32->512->512->126 (vposer_smpl.py:83-89) with nn.Linear default init."""
    import torch
    torch.manual_seed(seed)
    fc1, fc2, out = torch.nn.Linear(32, 512), torch.nn.Linear(512, 512), torch.nn.Linear(512, 126)
    return {'dec_fc1_w': fc1.weight.detach().numpy().copy(), 'dec_fc1_b': fc1.bias.detach().numpy().copy(),
            'dec_fc2_w': fc2.weight.detach().numpy().copy(), 'dec_fc2_b': fc2.bias.detach().numpy().copy(),
            'dec_out_w': out.weight.detach().numpy().copy(), 'dec_out_b': out.bias.detach().numpy().copy()}


def load_tables():
    return dict(np.load(os.path.join(ASSETS, 'lemo_tables.npz')))


def load_enc_weights():
    return dict(np.load(os.path.join(ASSETS, 'enc_smooth_15217.npz')))


def seed_clip(s):
    """One of the ten shipped [119,72] result clips + its [119,4] contact labels (data fixture)."""
    z = np.load(os.path.join(GOLDEN, 'seed_clips.npz'))
    stage = ('perframe', 'temp')[(s // 5) % 2]
    c = (0, 20, 40, 60, 80)[s % 5]
    return z['%s_params_%d' % (stage, c)].astype(np.float32), z['%s_contact_%d' % (stage, c)].astype(np.float32)


def make_sequence(s, T=119, noise=0.02):
    """Synthetic AMASS-shaped sequence s (SURVEY.md section 8d 'Sequences').
    Returns (params_clean [T,72], params_init [T,72], contact [T,4]).  72 = transl3, aa3, betas10,
    vposer-z32, lhand-pca12, rhand-pca12 (utils/utils.py:141-152).  T>119 pads by repeating the last frame."""
    import torch
    base, contact = seed_clip(s % 10)
    gen = torch.Generator().manual_seed(100 + s)
    clean = torch.from_numpy(base) + noise * torch.randn(base.shape, generator=gen)
    clean[:, 6:16] = clean[0:1, 6:16]                     # shape is per-sequence constant
    init = clean + noise * torch.randn(base.shape, generator=gen)
    init[:, 6:16] = clean[:, 6:16]
    clean, init = clean.numpy(), init.numpy()
    if T != clean.shape[0]:
        idx = np.minimum(np.arange(T), clean.shape[0] - 1)
        clean, init, contact = clean[idx], init[idx], contact[idx]
    return clean.astype(np.float32), init.astype(np.float32), contact.astype(np.float32)


def _rodrigues_np(aa):
    """Rodrigues formula for one axis-angle vector (float64 arithmetic, float32 result): fixed camera / scene rotations only."""
    a = np.asarray(aa, np.float64)
    th = np.linalg.norm(a)
    k = a / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K).astype(np.float32)


def synth_marker_clip(seed, T=120):
    """A moving, turning 68-point body (pelvis + 67 markers, z up) with float32 coordinates and 0/1 contact labels
    (input of the infill pre-stage, opt_amass_temp.py:141-150)."""
    g = np.random.default_rng(seed)
    base = g.standard_normal((68, 3)) * np.array([0.25, 0.12, 0.45]) + np.array([0.0, 0.0, 0.9])
    base[[27, 57], 0] += np.array([-0.2, 0.2])          # shoulders / hips apart along x so that `across` is well defined
    base[[28, 58], 0] += np.array([-0.15, 0.15])
    t = np.arange(T) / 30.0
    yaw = 0.6 * np.sin(0.7 * t) + 0.3 * t
    c, s_ = np.cos(yaw), np.sin(yaw)
    R = np.stack([np.stack([c, -s_, 0 * c], -1), np.stack([s_, c, 0 * c], -1), np.stack([0 * c, 0 * c, 1 + 0 * c], -1)], -2)
    pos = np.stack([0.8 * t + 0.1 * np.sin(2 * t), 0.3 * np.sin(0.9 * t), 0.02 * np.sin(5 * t)], -1)
    body = np.einsum('tij,kj->tki', R, base) + pos[:, None] + 0.01 * g.standard_normal((T, 68, 3))
    contact = (g.random((T, 4)) > 0.5).astype(np.float32)
    return body.astype(np.float32), contact


def make_prox_problem(B, D=32, m_scene=3000, seed=0):
    """Synthetic PROX stage-2 window: parameters on a wavy floor SDF, 2-D keypoint targets, camera,
    friction / contact vertex sets and scene points.  Returns (P numpy dict, cfg dict of torch tensors)."""
    import torch
    g = np.random.default_rng(seed)
    f32 = np.float32
    clean, _, _ = make_sequence(seed, T=B)
    P = dict(transl=clean[:, 0:3] + np.array([0, 0, 3.0], f32), global_orient=clean[:, 3:6], pose_embedding=clean[:, 16:48],
             left_hand_pose=clean[:, 48:60], right_hand_pose=clean[:, 60:72], jaw_pose=(0.05 * g.standard_normal((B, 3))).astype(f32),
             leye_pose=np.zeros((B, 3), f32), reye_pose=np.zeros((B, 3), f32), expression=(0.3 * g.standard_normal((B, 10))).astype(f32),
             betas=np.repeat(clean[:1, 6:16], B, 0))
    # the reference's own tables (exported data, tools/export_assets.py prox): OpenPose coco25 map with hands + face (118 joints,
    # temp_prox/main_slide.py:160-179), friction (307) and contact (1121) vertex ids (fit_temp_loadprox_slide.py:349-362)
    pt = np.load(os.path.join(ASSETS, 'prox_tables.npz'))
    jm = pt['smplx_coco25_h1_f1_c0'].astype(np.int64)
    Rc = torch.from_numpy(_rodrigues_np([0.02, -0.01, 0.03]))
    tc = torch.tensor([0.01, 0.02, 0.0])
    Rw = torch.from_numpy(_rodrigues_np([1.4, 0.1, -0.1]))
    tw = torch.tensor([0.1, -0.1, 0.45]) - Rw @ torch.from_numpy(P['transl'].mean(0))      # body centre just above the wavy floor
    xs = np.linspace(-3, 3, D, dtype=f32)
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing='ij')
    sdf = (Z - 0.3 + 0.2 * np.sin(2 * X) * np.cos(1.5 * Y)).astype(f32)             # wavy floor
    cfg = dict(gt_joints=torch.from_numpy((900 * g.random((B, 118, 2)) + 50).astype(f32)), joints_conf=torch.from_numpy((0.3 + 0.7 * g.random((B, 118))).astype(f32)),
               joint_weights=torch.ones(B, 118), joint_map=torch.from_numpy(jm), camera=(Rc, tc, 1060.53, 1060.38, torch.tensor([951.30, 536.77])),
               cam2world=(Rw, tw), sdf=torch.from_numpy(sdf), grid_min=torch.tensor([-3., -3., -3.]), grid_max=torch.tensor([3., 3., 3.]),
               fric_ids=torch.from_numpy(pt['friction_ids'].astype(np.int64)), contact_ids=torch.from_numpy(pt['contact_ids'].astype(np.int64)),
               scene_v=torch.from_numpy((g.random((m_scene, 3)) * np.array([6, 6, 0.1]) - np.array([3, 3, -0.25])).astype(f32)),
               w=dict(data=1.0, body_pose=4.78e-5 * 1e3, hand_prior=4.78e-5 * 1e3, expr=0.03, jaw=0.03, sdf=0.003, fric_t=20.0, fric_n=10.0,
                      contact=1.0, smooth=1e8))
    return P, cfg
