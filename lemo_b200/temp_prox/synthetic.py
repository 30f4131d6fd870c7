"""Synthetic PROX stage-2 window (BASELINE config 4) on the fused driver: shared by bench.py and tools/."""
import torch

from .. import synth
from .fused import ProxFitter


def s2_weights(w):
    """synth.make_prox_problem's short weight names -> the reference's SMPLifyLoss attribute names."""
    return dict(data_weight=w['data'], body_pose_weight=w['body_pose'], shape_weight=w.get('shape', 0.0),
                bending_prior_weight=3.17 * w['body_pose'], hand_prior_weight=w['hand_prior'], expr_prior_weight=w['expr'],
                jaw_prior_weight=w['jaw'], sdf_penetration_weight=w['sdf'], contact_loss_weight=w.get('contact', 0.0),
                motion_prior_smooth_weight=w['smooth'], friction_normal_weight=w['fric_n'], friction_tangent_weight=w['fric_t'])


def make_window(body, vposer, enc, B=100, D=256, m_scene=100000, seed=3, device='cuda', contact=True, first_batch_flag=False,
                use_cuda_graph=True):
    """A ProxFitter with a synthetic window loaded: B frames, full mesh, keypoints + priors + SDF D^3 + friction + contact against
    m_scene shared scene points + Enc smoothness.  Returns (fitter, P numpy dict, cfg)."""
    P, cfg = synth.make_prox_problem(B, D=D, m_scene=m_scene, seed=seed)
    dev = torch.device(device)
    fit = ProxFitter(body, vposer, enc, B, dev, joint_map=cfg['joint_map'], camera=cfg['camera'], cam2world=cfg['cam2world'],
                     sdf=cfg['sdf'].to(dev), grid_min=cfg['grid_min'], grid_max=cfg['grid_max'], fric_ids=cfg['fric_ids'],
                     contact_ids=cfg['contact_ids'] if contact else None, scene_v=cfg['scene_v'].to(dev) if contact else None,
                     weights=s2_weights(cfg['w']), contact=contact, use_cuda_graph=use_cuda_graph)
    fit.set_weights(s2_weights(cfg['w']), 0 if first_batch_flag else int(B * 0.15), True)
    fit.set_window(P, cfg['gt_joints'], cfg['joints_conf'], cfg['joint_weights'])
    return fit, P, cfg
