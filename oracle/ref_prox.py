"""CPU restatement of the PROX stage-2 loss (reference temp_prox/fitting_temp_slide.py:564-1062, terms active in
cfg_files/PROXD_temp_S2.yaml plus the `contact` term BASELINE.json's config 4 switches on).  TEST INFRASTRUCTURE ONLY."""
import numpy as np
import torch
import torch.nn.functional as F

from . import ref_body as rb
from . import ref_priors as rp
from .ref_loops import smooth_input


def project(points, R, t, fx, fy, c):
    """temp_prox/camera.py:93-116."""
    p = torch.einsum('ki,bji->bjk', R, points) + t
    return torch.stack([fx * p[..., 0] / p[..., 2] + c[0], fy * p[..., 1] / p[..., 2] + c[1]], -1)


def s2_loss(P, ctx, cfg):
    """P: dict of parameter tensors (transl, global_orient, pose_embedding, left/right_hand_pose, jaw_pose, leye_pose, reye_pose,
    expression, betas).  cfg: dict with gt_joints [B,Jm,2], joints_conf [B,Jm], joint_weights [B,Jm], joint_map [Jm], camera (R,t,fx,fy,c),
    cam2world (R,t), sdf [D,D,D], grid_min, grid_max, fric_ids, contact_ids, scene_v [m,3], weights dict.  Returns (total, terms)."""
    w = cfg['w']
    B = P['transl'].shape[0]
    body_pose = ctx.vposer.decode_aa(P['pose_embedding']).view(B, -1)                                   # :243
    verts, joints, full_pose = ctx.smplx(transl=P['transl'], global_orient=P['global_orient'], betas=P['betas'], body_pose=body_pose,
                                         left_hand_pose=P['left_hand_pose'], right_hand_pose=P['right_hand_pose'],
                                         expression=P['expression'], jaw_pose=P['jaw_pose'], leye_pose=P['leye_pose'], reye_pose=P['reye_pose'])
    mapped = joints[:, cfg['joint_map']]
    cam = cfg['camera']
    proj = project(mapped, *cam)                                                                         # :574
    wts = (cfg['joint_weights'] * cfg['joints_conf']).unsqueeze(-1)
    T = {}
    T['joint'] = torch.mean(wts ** 2 * torch.abs(cfg['gt_joints'] - proj)) * w['data']                   # :577-581
    T['pprior'] = P['pose_embedding'].pow(2).sum() * w['body_pose'] ** 2                                  # :587
    T['shape'] = P['betas'].pow(2).sum() * w.get('shape', 0.0) ** 2                                       # :592 (L2Prior on betas)
    idx = torch.tensor([55, 58, 12, 15]) - 3                                                             # prior.py:63-89
    sgn = torch.tensor([1., -1., -1., -1.], dtype=verts.dtype)
    T['angle'] = torch.sum(torch.exp(full_pose[:, 3:66][:, idx] * sgn)) * (3.17 * w['body_pose']) ** 2    # :596, fit_temp_loadprox_slide.py:524
    T['hand'] = (P['left_hand_pose'].pow(2).sum() + P['right_hand_pose'].pow(2).sum()) * w['hand_prior'] ** 2   # :601-607
    T['expr'] = P['expression'].pow(2).sum() * w['expr'] ** 2                                             # :612
    T['jaw'] = (P['jaw_pose'] * w['jaw']).pow(2).sum()                                                    # :616
    Rw, tw = cfg['cam2world']
    vw = torch.matmul(Rw, verts.permute(0, 2, 1)).permute(0, 2, 1) + tw                                  # :677
    jw = torch.matmul(Rw, joints.permute(0, 2, 1)).permute(0, 2, 1) + tw
    nv = vw.shape[1]
    D = cfg['sdf'].shape[-1]
    norm = (vw - cfg['grid_min']) / (cfg['grid_max'] - cfg['grid_min']) * 2 - 1
    # :684 -- the reference samples a [B,1,D,D,D] replica of ONE scene volume; sampling that volume frame by frame is the same arithmetic
    # without the B-fold copy (6.7 GB at B=100, D=256)
    vol = cfg['sdf'].to(verts.dtype)[None, None]
    body_sdf = torch.cat([F.grid_sample(vol, norm[b:b + 1][:, :, [2, 1, 0]].view(1, nv, 1, 1, 3), padding_mode='border',
                                        align_corners=False).view(1, nv) for b in range(B)], 0)
    neg = body_sdf < 0
    T['sdf'] = w['sdf'] * body_sdf[neg].abs().sum() if bool(neg.any()) else torch.zeros((), dtype=verts.dtype)   # :688-694
    # friction (:699-739), scene normal = +z
    fr = vw[:, cfg['fric_ids']]
    vel = fr[1:] - fr[:-1]
    sel = body_sdf[:-1][:, cfg['fric_ids']] < 0.01
    T['fric_t'] = torch.zeros((), dtype=verts.dtype)
    T['fric_n'] = torch.zeros((), dtype=verts.dtype)
    if bool(sel.any()):
        v = vel[sel]
        vn = v[:, 2]
        vt = torch.norm(torch.stack([v[:, 0], v[:, 1], torch.zeros_like(vn)], -1), dim=-1)
        if bool((vt > 1e-4).any()):
            T['fric_t'] = vt[vt > 1e-4].abs().mean() * w['fric_t']
        if bool((vn < 0).any()):
            T['fric_n'] = vn[vn < 0].abs().mean() * w['fric_n']
    # contact (:743-753), Chamfer to the (shared) scene
    T['contact'] = torch.zeros((), dtype=verts.dtype)
    if w.get('contact', 0) > 0:
        cv = vw[:, cfg['contact_ids']]
        scene = cfg['scene_v'].to(verts.dtype)
        if B * cv.shape[1] * scene.shape[0] <= 1 << 27:
            d1, _, _, _ = rp.chamfer(cv, scene[None].expand(B, -1, -1))
        else:
            # config-4 scale (100 x 1121 x 100 000): the [B,n,m] distance tensor of the torch restatement would need 45 GB; take the
            # nearest-neighbour INDICES from the C restatement (oracle/csrc/chamfer_ref.c) and form the distance differentiably,
            # which is what chamferFunction's forward + backward compute (dist_chamfer.py:10-45)
            from . import ref_chamfer
            _, idx = ref_chamfer.chamfer_nn(cv.detach().float().numpy(), scene.float().numpy())
            d1 = ((cv - scene[torch.from_numpy(idx).long()]) ** 2).sum(-1)
        r = torch.sqrt(d1 + 1e-4)
        T['contact'] = w['contact'] * (r / (r + 1.0)).mean()
    # smoothness prior on world markers (:997-1031): same canonicalisation as the AMASS script but in world coordinates
    xin = smooth_input(vw[:, ctx.m81], jw[0], ctx)
    z = rp.enc_forward(xin, ctx.enc_sd)
    T['smooth'] = torch.mean((z[..., 1:] - z[..., :-1]) ** 2) * w['smooth']
    return sum(T.values()), T


PKEYS = ['transl', 'global_orient', 'pose_embedding', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression']


def fit_window(P_np, ctx, cfg, n_iters, lr=0.005, first_batch_flag=False, trace=None):
    """FittingMonitor.run_fitting + create_fitting_closure.fitting_func (fitting_temp_slide.py:169-313) for one window with
    torch.optim.Adam (optim_factory.py:77-80): closure = loss + backward + `grad[0:int(bs*0.15)] = 0` unless it is the first window.
    Returns (dict of fitted numpy params, last loss)."""
    # .copy(): torch.from_numpy shares memory with the caller's arrays and Adam updates in place
    P = {k: torch.from_numpy(np.array(v, copy=True)).to(ctx.dtype).requires_grad_(k in PKEYS) for k, v in P_np.items()}
    params = [P[k] for k in PKEYS]
    opt = torch.optim.Adam(params, lr=lr)
    bs = P['transl'].shape[0]
    erase_n = int(bs * 0.15)
    last = None
    for _ in range(n_iters):
        def closure():
            opt.zero_grad()
            tot, T = s2_loss(P, ctx, cfg)
            tot.backward()
            if not first_batch_flag:
                for p_ in params:
                    if p_.grad is not None:
                        p_.grad[0:erase_n, :] = 0
            if trace is not None:
                trace.append({k: float(v) for k, v in T.items()})
            return tot
        last = float(opt.step(closure))
    return {k: v.detach().numpy() for k, v in P.items()}, last


def scan_terms(vertices, scan, scan_num, vis, body_mask, rho_s2m, rho_m2s, w_s2m, w_m2s, s2m=True, m2s=True):
    """TEST ORACLE for the scan-to-mesh / mesh-to-scan terms (reference temp_prox/fitting_temp_slide.py:638-670), brute-force distances.
    vertices [bs,V,3] (may require grad), scan [bs,N,3], scan_num [bs], vis [bs,V] bool, body_mask [V] bool.  Follows the reference's
    batch behaviour: its Chamfer wrapper sizes the batch from the scan slice (dist_chamfer.py:13, batch 1), so the visible vertices are
    always read from frame 0 of `vertices[:, visible_i, :]`."""
    gm = lambda r, rho: rho ** 2 * r ** 2 / (r ** 2 + rho ** 2)
    l1, l2 = [], []
    for i in range(vertices.shape[0]):
        cur = scan[i, :int(scan_num[i])]
        if not bool(vis[i].any()):
            continue
        if s2m and w_s2m > 0:
            d = ((cur[:, None, :] - vertices[0][vis[i]][None]) ** 2).sum(-1).min(dim=1)[0]
            l1.append(gm(torch.sqrt(d), rho_s2m).mean())
        if m2s and w_m2s > 0:
            d = ((vertices[0][vis[i] & body_mask][:, None, :] - cur[None]) ** 2).sum(-1).min(dim=1)[0]
            l2.append(gm(torch.sqrt(d), rho_m2s).mean())
    z = torch.zeros((), dtype=vertices.dtype)
    return (sum(l1) / len(l1) * w_s2m if l1 else z), (sum(l2) / len(l2) * w_m2s if l2 else z)
