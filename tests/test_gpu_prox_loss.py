"""PROX stage-2 loss (config 4 shape, reduced sizes): value and gradients of every term against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import synth, ref_body as rb, ref_prox
from gpu_common import DEV, smplx_module, vposer_module, enc_module, oracle_ctx, model_np, rel

pytestmark = pytest.mark.gpu
PKEYS = ['transl', 'global_orient', 'pose_embedding', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression']


_setup = synth.make_prox_problem


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
