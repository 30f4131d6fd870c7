"""PROX stage-2 loss (config 4 shape, reduced sizes): value and gradients of every term against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import synth, ref_body as rb, ref_prox
from gpu_common import DEV, smplx_module, vposer_module, enc_module, oracle_ctx, model_np, rel

pytestmark = pytest.mark.gpu
PKEYS = ['transl', 'global_orient', 'pose_embedding', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression']


def _setup(B, D=32, m_scene=3000, seed=0):
    g = np.random.default_rng(seed)
    f32 = np.float32
    clean, _, _ = synth.make_sequence(seed, T=B)
    P = dict(transl=clean[:, 0:3] + np.array([0, 0, 3.0], f32), global_orient=clean[:, 3:6], pose_embedding=clean[:, 16:48],
             left_hand_pose=clean[:, 48:60], right_hand_pose=clean[:, 60:72], jaw_pose=(0.05 * g.standard_normal((B, 3))).astype(f32),
             leye_pose=np.zeros((B, 3), f32), reye_pose=np.zeros((B, 3), f32), expression=(0.3 * g.standard_normal((B, 10))).astype(f32),
             betas=np.repeat(clean[:1, 6:16], B, 0))
    jm = g.permutation(127)[:118].astype(np.int64)
    Rc = rb.rodrigues(torch.tensor([[0.02, -0.01, 0.03]]))[0]
    tc = torch.tensor([0.01, 0.02, 0.0])
    Rw = rb.rodrigues(torch.tensor([[1.4, 0.1, -0.1]]))[0]
    tw = torch.tensor([0.1, -0.1, 0.45]) - Rw @ torch.from_numpy(P['transl'].mean(0))      # body centre just above the wavy floor
    xs = np.linspace(-3, 3, D, dtype=f32)
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing='ij')
    sdf = (Z - 0.3 + 0.2 * np.sin(2 * X) * np.cos(1.5 * Y)).astype(f32)             # wavy floor
    cfg = dict(gt_joints=torch.from_numpy((900 * g.random((B, 118, 2)) + 50).astype(f32)), joints_conf=torch.from_numpy((0.3 + 0.7 * g.random((B, 118))).astype(f32)),
               joint_weights=torch.ones(B, 118), joint_map=torch.from_numpy(jm), camera=(Rc, tc, 1060.53, 1060.38, torch.tensor([951.30, 536.77])),
               cam2world=(Rw, tw), sdf=torch.from_numpy(sdf), grid_min=torch.tensor([-3., -3., -3.]), grid_max=torch.tensor([3., 3., 3.]),
               fric_ids=torch.from_numpy(g.choice(synth.V, 307, replace=False)), contact_ids=torch.from_numpy(g.choice(synth.V, 1121, replace=False)),
               scene_v=torch.from_numpy((g.random((m_scene, 3)) * np.array([6, 6, 0.1]) - np.array([3, 3, -0.25])).astype(f32)),
               w=dict(data=1.0, body_pose=4.78e-5 * 1e3, hand_prior=4.78e-5 * 1e3, expr=0.03, jaw=0.03, sdf=0.003, fric_t=20.0, fric_n=10.0,
                      contact=1.0, smooth=1e8))
    return P, cfg


def test_s2_loss_value_and_gradients():
    from lemo_b200.temp_prox.camera import PerspectiveCamera
    from lemo_b200.temp_prox.fitting_temp_slide import SMPLifyLoss
    B = 24
    P, cfg = _setup(B)
    c32 = oracle_ctx(torch.float32)
    # ---- oracle (CPU, fp32)
    Pt = {k: torch.from_numpy(v).requires_grad_(k in PKEYS) for k, v in P.items()}
    tot_ref, T_ref = ref_prox.s2_loss(Pt, c32, cfg)
    tot_ref.backward()
    # ---- lemo_b200 operators
    tab = synth.load_tables()
    body = smplx_module(synth.V)
    body.joint_mapper = lambda j: j[:, cfg['joint_map'].to(DEV)]
    vp, enc = vposer_module(), enc_module()
    Rc, tc, fx, fy, cc = cfg['camera']
    cam = PerspectiveCamera(rotation=Rc[None].repeat(B, 1, 1), translation=tc[None].repeat(B, 1), focal_length_x=fx, focal_length_y=fy,
                            batch_size=B, center=cc[None].repeat(B, 1)).to(DEV)
    lossf = SMPLifyLoss(cfg['w'], cam, cfg['cam2world'], cfg['sdf'].to(DEV), cfg['grid_min'], cfg['grid_max'], cfg['fric_ids'].to(DEV),
                        cfg['contact_ids'].to(DEV), cfg['scene_v'].to(DEV), torch.from_numpy(tab['markers81']).long().to(DEV), enc,
                        torch.from_numpy(tab['smooth_Xmean']).view(1, 1, 243).to(DEV), torch.from_numpy(tab['smooth_Xstd']).to(DEV),
                        cfg['joint_weights'].to(DEV))
    Pg = {k: torch.from_numpy(v).to(DEV).requires_grad_(k in PKEYS) for k, v in P.items()}
    R_body = vp.decode(Pg['pose_embedding'], 'matrot').reshape(B, 21, 9)
    kw = {k: Pg[k] for k in ('transl', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression', 'betas')}
    out = body(return_verts=True, return_full_pose=True, R_body=R_body, **kw)
    mapper = body.joint_mapper
    body.joint_mapper = None
    raw = body(return_verts=True, R_body=R_body, **kw).joints
    body.joint_mapper = mapper
    tot, T = lossf(out, raw, cfg['gt_joints'].to(DEV), cfg['joints_conf'].to(DEV), Pg['pose_embedding'])
    tot.backward()
    body.joint_mapper = None
    for k in T_ref:
        a, b = float(T[k]), float(T_ref[k])
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-7, (k, a, b)
    assert float(T_ref['sdf']) > 0 and float(T_ref['contact']) > 0 and float(T_ref['fric_t']) > 0          # the terms are exercised
    for k in PKEYS:
        if Pt[k].grad is None:
            continue
        e = rel(Pg[k].grad, Pt[k].grad)
        assert e < 5e-3, (k, e)
