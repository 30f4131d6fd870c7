"""Drop-in for the reference's temp_prox/fitting_temp_slide.py: `create_loss`, `SMPLifyLoss`, `FittingMonitor` with the
reference's signatures (SMPLifyLoss.__init__ :327-538, reset_loss_weights :548-562, forward :564-1062; FittingMonitor.__init__
:137-154, run_fitting :169-217, create_fitting_closure :220-313), so `fit_temp_loadprox_slide.fit_single_frame` runs on it unchanged.

Two execution paths behind that surface:

* fused (default): when the optimiser is Adam (optim_factory.create_optimizer 'adam', S2.yaml) and the loss's active terms are the
  ones the device driver covers -- 2-D keypoints, VPoser/shape/angle/hand/expression/jaw priors, SDF penetration, friction, contact,
  Enc smoothness prior, first-15 % gradient erase -- `run_fitting` hands the whole `maxiters` loop to ONE C-ABI call
  (lemo_fit_prox_run, csrc/fit_prox.cu: a CUDA graph per closure step, no host synchronisation) and writes the fitted parameters
  back into `body_model` / `pose_embedding`.
* eager: `optimizer.step(closure)` exactly like the reference for anything else (LBFGS, s2m/m2s, smooth_acc/vel, GMM prior ...);
  the closure evaluates `SMPLifyLoss.forward`, whose heavy operators (SMPL-X, VPoser, camera, SDF lookup, Chamfer, Enc) are the
  lemo_b200 kernels wrapped as autograd Functions and whose elementwise glue is PyTorch on the device.

There is no CPU path in either.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import misc_utils as utils
from .camera import cam_to_world, sdf_sample, one_volume
from .dist_chamfer import chamferDist
from .fused import ProxFitter, WEIGHT_KEYS

distChamfer = chamferDist()


def _zero(ref):
    return torch.zeros((), dtype=ref.dtype, device=ref.device)


def _masked_abs_mean(x, mask):
    """`x[mask].abs().mean()` if any(mask) else 0 -- on device, without the reference's `.item()` host sync."""
    m = mask.to(x.dtype)
    n = m.sum()
    return torch.where(n > 0, (x.abs() * m).sum() / n.clamp(min=1.0), _zero(x))


def _canonical_rotmat(joints_frame0):
    """fitting_temp_slide.py:1003-1010 (also :763-770): x = hip axis projected on the floor, z up, y = z x x."""
    x_axis = joints_frame0[2, :] - joints_frame0[1, :]
    x_axis = torch.stack([x_axis[0], x_axis[1], torch.zeros((), dtype=x_axis.dtype, device=x_axis.device)])
    x_axis = x_axis / torch.norm(x_axis)
    z_axis = torch.tensor([0., 0., 1.], dtype=x_axis.dtype, device=x_axis.device)
    y_axis = torch.cross(z_axis, x_axis, dim=0)
    y_axis = y_axis / torch.norm(y_axis)
    return torch.stack([x_axis, y_axis, z_axis], dim=1)


class SMPLifyLoss(nn.Module):

    def __init__(self, search_tree=None, pen_distance=None, tri_filtering_module=None,
                 body_pose_prior=None, shape_prior=None, expr_prior=None, angle_prior=None, jaw_prior=None,
                 use_joints_conf=True, use_face=True, use_hands=True, left_hand_prior=None, right_hand_prior=None,
                 interpenetration=True, dtype=torch.float32,
                 data_weight=1.0, body_pose_weight=0.0, shape_weight=0.0, bending_prior_weight=0.0, hand_prior_weight=0.0,
                 expr_prior_weight=0.0, jaw_prior_weight=0.0, coll_loss_weight=0.0,
                 s2m=False, m2s=False, rho_s2m=1, rho_m2s=1, s2m_weight=0.0, m2s_weight=0.0, head_mask=None, body_mask=None,
                 sdf_penetration=False, voxel_size=None, grid_min=None, grid_max=None, sdf=None, sdf_normals=None,
                 sdf_penetration_weight=0.0, R=None, t=None,
                 contact=False, contact_loss_weight=0.0, contact_verts_ids=None,
                 smooth_acc=False, smooth_acc_weight=0.0, smooth_vel=False, smooth_vel_weight=0.0,
                 use_motion_smooth_prior=False, motion_prior_smooth_weight=0.0, motion_smooth_model=None,
                 use_friction=False, friction_normal_weight=0.0, friction_tangent_weight=0.0, contact_fric_verts_ids=None,
                 use_motion_infill_prior=False, motion_infill_rec_weight=0.0, motion_infill_contact_weight=0.0,
                 motion_infill_model=None, infill_pretrain_weights=None,
                 device=None, **kwargs):
        super(SMPLifyLoss, self).__init__()
        self.device = device
        self.use_joints_conf = use_joints_conf
        self.angle_prior = angle_prior
        self.s2m, self.m2s = s2m, m2s
        self.s2m_robustifier = utils.GMoF(rho=rho_s2m)
        self.m2s_robustifier = utils.GMoF(rho=rho_m2s)
        self.body_pose_prior = body_pose_prior
        self.shape_prior = shape_prior
        self.body_mask, self.head_mask = body_mask, head_mask
        self.R, self.t = R, t
        self.interpenetration = interpenetration
        if self.interpenetration:
            self.search_tree, self.tri_filtering_module, self.pen_distance = search_tree, tri_filtering_module, pen_distance
        self.use_hands = use_hands
        if self.use_hands:
            self.left_hand_prior, self.right_hand_prior = left_hand_prior, right_hand_prior
        self.use_face = use_face
        if self.use_face:
            self.expr_prior, self.jaw_prior = expr_prior, jaw_prior

        reg = lambda name, v: self.register_buffer(name, torch.tensor(v, dtype=dtype))
        reg('data_weight', data_weight); reg('body_pose_weight', body_pose_weight); reg('shape_weight', shape_weight)
        reg('bending_prior_weight', bending_prior_weight)
        if self.use_hands:
            reg('hand_prior_weight', hand_prior_weight)
        if self.use_face:
            reg('expr_prior_weight', expr_prior_weight); reg('jaw_prior_weight', jaw_prior_weight)
        if self.interpenetration:
            reg('coll_loss_weight', coll_loss_weight)
        reg('s2m_weight', s2m_weight); reg('m2s_weight', m2s_weight)

        self.sdf_penetration = sdf_penetration
        self.use_friction = use_friction
        if self.sdf_penetration or self.use_friction:
            # ONE [D,D,D] volume is kept (the reference's caller repeats it B times, fit_temp_loadprox_slide.py:297-299: both accepted)
            self.sdf = None if sdf is None else one_volume(sdf)
            self.sdf_normals, self.voxel_size = sdf_normals, voxel_size
            self.grid_min = None if grid_min is None else grid_min.reshape(-1, 3)[0]
            self.grid_max = None if grid_max is None else grid_max.reshape(-1, 3)[0]
        if self.sdf_penetration:
            reg('sdf_penetration_weight', sdf_penetration_weight)
        self.contact = contact
        if self.contact:
            self.contact_verts_ids = contact_verts_ids
            reg('contact_loss_weight', contact_loss_weight)
        self.smooth_acc = smooth_acc
        if self.smooth_acc:
            reg('smooth_acc_weight', smooth_acc_weight)
        self.smooth_vel = smooth_vel
        if self.smooth_vel:
            reg('smooth_vel_weight', smooth_vel_weight)
        if self.use_friction:
            self.contact_fric_verts_ids = contact_fric_verts_ids
            reg('friction_normal_weight', friction_normal_weight); reg('friction_tangent_weight', friction_tangent_weight)

        self.use_motion_infill_prior = use_motion_infill_prior
        if self.use_motion_infill_prior:
            raise NotImplementedError('the in-loss motion-infilling prior of PROX stage 3 (fitting_temp_slide.py:757-955) is not on the '
                                      'stage-2 path; the infill pre-stage itself is lemo_b200.infill.InfillStage')
        self.use_motion_smooth_prior = use_motion_smooth_prior
        if self.use_motion_smooth_prior:
            self.motion_smooth_model = motion_smooth_model
            reg('motion_prior_smooth_weight', motion_prior_smooth_weight)

        # index tables / statistics: the reference opens ../loader/*.json and ../preprocess_stats/*.npz relative to temp_prox/
        # (:513-527); the same data ships in lemo_b200/assets (tools/export_assets.py)
        from ..fit import load_tables
        tb = load_tables()
        self.smooth_marker_ids = [int(i) for i in tb['markers81']]
        if self.use_motion_smooth_prior:
            self.infill_marker_ids = [int(i) for i in tb['markers67']]
            self.Xmean_global_markers = torch.from_numpy(tb['smooth_Xmean']).float().reshape(1, 1, 243).to(device)
            self.Xstd_global_markers = torch.from_numpy(tb['smooth_Xstd']).float().to(device)
        self._fused = {}

    def reset_loss_weights(self, loss_weight_dict):
        for key in loss_weight_dict:
            if hasattr(self, key):
                weight_tensor = getattr(self, key)
                if 'torch.Tensor' in str(type(loss_weight_dict[key])):
                    weight_tensor = loss_weight_dict[key].clone().detach()
                else:
                    weight_tensor = torch.tensor(loss_weight_dict[key], dtype=weight_tensor.dtype, device=weight_tensor.device)
                setattr(self, key, weight_tensor)

    # ------------------------------------------------------------------------------------------------ fused-path helpers
    def weight_dict(self):
        """The driver's weights (reference attribute names) as floats; terms switched off contribute weight 0."""
        g = lambda k, on=True: float(getattr(self, k)) if (on and hasattr(self, k)) else 0.0
        return dict(data_weight=g('data_weight'), body_pose_weight=g('body_pose_weight'), shape_weight=g('shape_weight'),
                    bending_prior_weight=g('bending_prior_weight'),
                    hand_prior_weight=g('hand_prior_weight', self.use_hands and self.left_hand_prior is not None),
                    expr_prior_weight=g('expr_prior_weight', self.use_face), jaw_prior_weight=g('jaw_prior_weight', self.use_face),
                    sdf_penetration_weight=g('sdf_penetration_weight', self.sdf_penetration),
                    contact_loss_weight=g('contact_loss_weight', self.contact),
                    motion_prior_smooth_weight=g('motion_prior_smooth_weight', self.use_motion_smooth_prior),
                    friction_normal_weight=g('friction_normal_weight', self.use_friction),
                    friction_tangent_weight=g('friction_tangent_weight', self.use_friction))

    def fusable(self, use_vposer=True, scan_tensor=None):
        """(ok, reason): can lemo_fit_prox_run evaluate this loss?"""
        from .prior import L2Prior, SMPLifyAnglePrior
        if not use_vposer:
            return False, 'use_vposer=False (body-pose prior on axis-angles)'
        if self.interpenetration and float(self.coll_loss_weight) > 0:
            return False, 'self-interpenetration term active'
        if (self.s2m or self.m2s) and (float(self.s2m_weight) > 0 or float(self.m2s_weight) > 0) and scan_tensor is not None:
            return False, 's2m / m2s terms active'
        if (self.smooth_acc and float(self.smooth_acc_weight) > 0) or (self.smooth_vel and float(self.smooth_vel_weight) > 0):
            return False, 'smooth_acc / smooth_vel terms active'
        for name in ('shape_prior', 'left_hand_prior', 'right_hand_prior', 'expr_prior', 'jaw_prior'):
            p = getattr(self, name, None)
            if p is not None and not isinstance(p, L2Prior):
                return False, '%s is not an L2Prior' % name
        if not isinstance(self.angle_prior, SMPLifyAnglePrior):
            return False, 'angle_prior is not SMPLifyAnglePrior'
        if self.R is None or self.t is None:
            return False, 'no cam2world transform'
        if self.use_hands and (self.left_hand_prior is None) != (self.right_hand_prior is None):
            return False, 'only one hand prior'
        return True, ''

    # ------------------------------------------------------------------------------------------------ eager evaluation
    def forward(self, body_model, body_model_output, smplx_joints, camera, gt_joints, joints_conf, marker_mask,
                body_model_faces, joint_weights, use_vposer=False, pose_embedding=None, scan_tensor=None, scan_point_num=None,
                scene_v=None, opt_step=None, **kwargs):
        out = body_model_output
        dev = out.joints.device
        zero = torch.zeros((), device=dev)
        # ---- 2d keypoint loss (:573-581)
        projected_joints = camera(out.joints)
        weights = (joint_weights * joints_conf if self.use_joints_conf else joint_weights).unsqueeze(dim=-1)
        joint_loss = torch.mean(weights ** 2 * torch.abs(gt_joints - projected_joints)) * self.data_weight
        # ---- pose / shape priors (:584-616)
        if use_vposer:
            pprior_loss = pose_embedding.pow(2).sum() * self.body_pose_weight ** 2
        else:
            pprior_loss = torch.sum(self.body_pose_prior(out.body_pose, out.betas)) * self.body_pose_weight ** 2
        shape_loss = torch.sum(self.shape_prior(out.betas)) * self.shape_weight ** 2
        body_pose = out.full_pose[:, 3:66]
        angle_prior_loss = torch.sum(self.angle_prior(body_pose)) * self.bending_prior_weight ** 2
        left_hand_prior_loss, right_hand_prior_loss = zero, zero
        if self.use_hands and self.left_hand_prior is not None:
            left_hand_prior_loss = torch.sum(self.left_hand_prior(out.left_hand_pose)) * self.hand_prior_weight ** 2
        if self.use_hands and self.right_hand_prior is not None:
            right_hand_prior_loss = torch.sum(self.right_hand_prior(out.right_hand_pose)) * self.hand_prior_weight ** 2
        expression_loss, jaw_prior_loss = zero, zero
        if self.use_face:
            expression_loss = torch.sum(self.expr_prior(out.expression)) * self.expr_prior_weight ** 2
            if hasattr(self, 'jaw_prior'):
                jaw_prior_loss = torch.sum(self.jaw_prior(out.jaw_pose.mul(self.jaw_prior_weight)))
        # ---- self-penetration (:619-635): needs the external mesh_intersection BVH (out of scope, gated off in S2/S3)
        pen_loss = zero
        if self.interpenetration and float(self.coll_loss_weight) > 0:
            if self.search_tree is None or self.pen_distance is None:
                raise RuntimeError('interpenetration needs the external mesh_intersection package (search_tree / pen_distance)')
            bs = projected_joints.shape[0]
            triangles = torch.index_select(out.vertices, 1, body_model_faces).view(bs, -1, 3, 3)
            with torch.no_grad():
                collision_idxs = self.search_tree(triangles).detach()
            if self.tri_filtering_module is not None:
                for i in range(bs):
                    collision_idxs[i:i + 1] = self.tri_filtering_module(collision_idxs[i:i + 1])
            if collision_idxs.ge(0).sum().item() > 0:
                pen_loss = torch.sum(self.coll_loss_weight * self.pen_distance(triangles, collision_idxs))
        # ---- scan <-> mesh (:638-670) on the lemo Chamfer kernels; visibility from the caller (kwargs['vis'], [B,V] 0/1) or psbody
        s2m_dist, m2s_dist = zero, zero
        if (self.s2m or self.m2s) and (self.s2m_weight > 0 or self.m2s_weight > 0) and scan_tensor is not None:
            s2m_dist, m2s_dist = self._scan_terms(out, body_model_faces, scan_tensor, scan_point_num, kwargs.get('vis'))
        # ---- to world coordinates (:673-679)
        vertices_world = smplx_joints_world = None
        if self.R is not None and self.t is not None:
            vertices_world = cam_to_world(out.vertices, self.R, self.t)
            smplx_joints_world = cam_to_world(smplx_joints, self.R, self.t)
        # ---- SDF penetration (:682-694) and friction (:699-739); one lookup serves both
        sdf_penetration_loss, loss_fric_tangent, loss_fric_normal = zero, zero, zero
        body_sdf = None
        if (self.sdf_penetration and self.sdf_penetration_weight > 0) or self.use_friction:
            body_sdf = sdf_sample(self.sdf, vertices_world, self.grid_min, self.grid_max)          # [B,V]
        if self.sdf_penetration and self.sdf_penetration_weight > 0:
            sdf_penetration_loss = self.sdf_penetration_weight * torch.clamp(-body_sdf, min=0).sum()
        if self.use_friction:
            ids = torch.as_tensor(np.asarray(self.contact_fric_verts_ids), dtype=torch.long, device=dev)
            fr = vertices_world[:, ids, :]
            vel = fr[1:] - fr[:-1]
            sel = body_sdf[:-1][:, ids] < 0.01
            vn = vel[..., 2]
            vt = torch.sqrt(vel[..., 0] ** 2 + vel[..., 1] ** 2 + 1e-30)
            loss_fric_tangent = _masked_abs_mean(vt, sel & (vt > 1e-4)) * self.friction_tangent_weight
            loss_fric_normal = _masked_abs_mean(vn, sel & (vn < 0)) * self.friction_normal_weight
        # ---- contact (:743-753): nearest scene vertex, ONE shared scene (the reference repeats it B times)
        contact_loss = zero
        if self.contact and self.contact_loss_weight > 0:
            ids = torch.as_tensor(np.asarray(self.contact_verts_ids), dtype=torch.long, device=dev)
            contact_body_vertices = vertices_world[:, ids, :]
            contact_dist, _, _, _ = distChamfer(contact_body_vertices.contiguous(), scene_v.reshape(1, -1, 3).contiguous())
            contact_dist = torch.sqrt(contact_dist + 1e-4) / (torch.sqrt(contact_dist + 1e-4) + 1.0)
            contact_loss = self.contact_loss_weight * contact_dist.mean()
        # ---- smooth acceleration / velocity (:756-774)
        smooth_acc_loss, smooth_vel_loss = zero, zero
        if (self.smooth_acc and self.smooth_acc_weight > 0) or (self.smooth_vel and self.smooth_vel_weight > 0):
            markers_smooth = out.vertices[:, self.smooth_marker_ids, :]
            markers_vel = markers_smooth[1:] - markers_smooth[0:-1]
            if self.smooth_acc and self.smooth_acc_weight > 0:
                markers_acc = markers_vel[1:] - markers_vel[0:-1]
                smooth_acc_loss = torch.mean(markers_acc ** 2) * self.smooth_acc_weight
            if self.smooth_vel and self.smooth_vel_weight > 0:
                smooth_vel_loss = torch.mean(markers_vel ** 2) * self.smooth_vel_weight
        motion_infill_loss, motion_infill_contact_loss = zero, zero
        # ---- motion smoothness prior (:997-1031)
        motion_prior_smooth_loss = zero
        if self.use_motion_smooth_prior:
            markers_smooth = vertices_world[:, self.smooth_marker_ids, :]
            joints_3d = smplx_joints_world[:, 0:75]
            transf_rotmat = _canonical_rotmat(joints_3d[0].detach())
            markers_frame0 = markers_smooth[0].detach()
            markers_smooth = torch.matmul(markers_smooth - markers_frame0[0], transf_rotmat)
            clip_img = markers_smooth.reshape(markers_smooth.shape[0], -1).unsqueeze(0)
            clip_img = (clip_img - self.Xmean_global_markers.to(dev)) / self.Xstd_global_markers.to(dev)
            clip_img = clip_img.permute(0, 2, 1).unsqueeze(1)
            clip_img_v = F.pad(clip_img[:, :, :, 1:] - clip_img[:, :, :, 0:-1], (8, 8, 1, 1), 'reflect')
            motion_z = self.motion_smooth_model(clip_img_v)[0]
            motion_z_v = motion_z[:, :, :, 1:] - motion_z[:, :, :, 0:-1]
            motion_prior_smooth_loss = torch.mean(motion_z_v ** 2) * self.motion_prior_smooth_weight

        total_loss = (joint_loss + pprior_loss + shape_loss + angle_prior_loss + pen_loss + jaw_prior_loss + expression_loss +
                      left_hand_prior_loss + right_hand_prior_loss + m2s_dist + s2m_dist + sdf_penetration_loss + contact_loss +
                      smooth_acc_loss + smooth_vel_loss + motion_prior_smooth_loss + loss_fric_tangent + loss_fric_normal +
                      motion_infill_loss + motion_infill_contact_loss)
        return {'total_loss': total_loss, 'joint_loss': joint_loss, 's2m_dist': s2m_dist, 'm2s_dist': m2s_dist,
                'self_penetration_loss': pen_loss, 'sdf_penetration_loss': sdf_penetration_loss, 'contact_loss': contact_loss,
                'smooth_acc_loss': smooth_acc_loss, 'smooth_vel_loss': smooth_vel_loss,
                'motion_prior_smooth_loss': motion_prior_smooth_loss, 'loss_fric_tangent': loss_fric_tangent,
                'loss_fric_normal': loss_fric_normal, 'motion_infill_loss': motion_infill_loss,
                'motion_infill_contact_loss': motion_infill_contact_loss,
                # extra keys (not in the reference's dict): the prior terms, for term-by-term parity tests
                'pprior_loss': pprior_loss, 'shape_loss': shape_loss, 'angle_prior_loss': angle_prior_loss,
                'hand_prior_loss': left_hand_prior_loss + right_hand_prior_loss, 'expression_loss': expression_loss,
                'jaw_prior_loss': jaw_prior_loss}

    def _scan_terms(self, out, body_model_faces, scan_tensor, scan_point_num, vis):
        """s2m / m2s (:638-670): per frame, Chamfer between the valid scan points and the camera-visible vertices (m2s: visible AND
        body_mask), GMoF-robustified, averaged over frames.  `vis` [B,V] (0/1) replaces psbody's visibility_compute when given.
        Reference behaviour kept: the reference passes `vertices[:, visible_i, :]` ([bs, n_vis, 3]) beside a [1, N, 3] scan slice, and
        its Chamfer wrapper takes the batch size from the FIRST argument (dist_chamfer.py:13) -- so for every frame i the scan of frame i
        is matched against the vertices of frame 0 selected by frame i's visibility.  At bs = 1 (the single-frame PROX fit the term was
        written for) this is the expected pairing; both shipped temporal configurations switch the terms off."""
        dev = out.vertices.device
        bs = out.vertices.shape[0]
        if vis is None:
            try:
                from psbody.mesh.visibility import visibility_compute
                from psbody.mesh import Mesh
            except ImportError as e:
                raise RuntimeError('s2m / m2s need per-vertex visibility: pass vis=[B,V] to the loss or install psbody.mesh') from e
            v_np = out.vertices.detach().cpu().numpy()
            f_np = body_model_faces.detach().cpu().numpy().reshape(-1, 3)
            vis = np.stack([visibility_compute(v=Mesh(v=v_np[i], f=f_np).v, f=f_np.astype(np.uint32),
                                               cams=np.array([[0.0, 0.0, 0.0]]))[0].squeeze() for i in range(bs)])
        vis = torch.as_tensor(np.asarray(vis) if not torch.is_tensor(vis) else vis).to(dev) > 0
        body_mask = torch.as_tensor(np.asarray(self.body_mask), device=dev).bool() if self.body_mask is not None else torch.ones_like(vis[0])
        s2m_list, m2s_list = [], []
        for i in range(bs):
            cur = scan_tensor[i:i + 1][:, 0:int(scan_point_num[i])].contiguous()
            if self.s2m and self.s2m_weight > 0 and bool(vis[i].any()):
                d, _, _, _ = distChamfer(cur, out.vertices[0:1][:, vis[i], :].contiguous())
                s2m_list.append(self.s2m_robustifier(torch.sqrt(d + 1e-30)).mean())
            if self.m2s and self.m2s_weight > 0 and bool(vis[i].any()):
                _, d, _, _ = distChamfer(cur, out.vertices[0:1][:, vis[i] & body_mask, :].contiguous())
                m2s_list.append(self.m2s_robustifier(torch.sqrt(d + 1e-30)).mean())
        zero = torch.zeros((), device=dev)
        s2m = sum(s2m_list) / len(s2m_list) * self.s2m_weight if s2m_list else zero
        m2s = sum(m2s_list) / len(m2s_list) * self.m2s_weight if m2s_list else zero
        return s2m, m2s


def create_loss(loss_type='smplify', **kwargs):
    """fitting_temp_slide.py:316-322.  'camera_init' (SMPLifyCameraInitLoss) belongs to PROX stage 1, not to this path."""
    if loss_type == 'smplify':
        return SMPLifyLoss(**kwargs)
    raise ValueError('Unknown loss type: {}'.format(loss_type))


class FittingMonitor(object):
    def __init__(self, summary_steps=1, maxiters=100, ftol=2e-09, gtol=1e-05, body_color=(1.0, 1.0, 0.9, 1.0), model_type='smpl',
                 **kwargs):
        super(FittingMonitor, self).__init__()
        self.maxiters = maxiters
        self.ftol, self.gtol = ftol, gtol
        self.summary_steps = summary_steps
        self.body_color = body_color
        self.model_type = model_type
        self.steps = 0
        self.fused = os.environ.get('LEMO_PROX_FUSED', '1') != '0'      # A/B switch: '0' forces the eager closure loop
        self.log_every = int(kwargs.get('fused_log_every', 50))           # fused path: loss read-back cadence when a writer is attached
        self.last_path = None

    def __enter__(self):
        self.steps = 0
        return self

    def __exit__(self, exception_type, exception_value, traceback):
        print('total steps:', self.steps)

    # ------------------------------------------------------------------------------------------------ run_fitting (:169-217)
    def run_fitting(self, optimizer, closure, params, body_model, use_vposer=True, pose_embedding=None, vposer=None, **kwargs):
        spec = getattr(closure, 'lemo_spec', None)
        ok, why = self._can_fuse(optimizer, spec, params, body_model, use_vposer, pose_embedding, vposer)
        if ok:
            self.last_path = 'fused'
            return self._run_fused(optimizer, spec, body_model, pose_embedding, vposer)
        self.last_path = 'eager (%s)' % why
        prev_loss = None
        for n in range(self.maxiters):
            loss = optimizer.step(closure)
            if torch.isnan(loss).sum() > 0:
                print('NaN loss value, stopping!')
                break
            if torch.isinf(loss).sum() > 0:
                print('Infinite loss value, stopping!')
                break
            prev_loss = loss.item()
        return prev_loss

    def _can_fuse(self, optimizer, spec, params, body_model, use_vposer, pose_embedding, vposer):
        if not self.fused:
            return False, 'LEMO_PROX_FUSED=0'
        if spec is None:
            return False, 'closure was not made by create_fitting_closure'
        if not isinstance(optimizer, torch.optim.Adam) or len(optimizer.param_groups) != 1:
            return False, 'optimizer is not a single-group Adam'
        g = optimizer.param_groups[0]
        if tuple(g['betas']) != (0.9, 0.999) or g['eps'] != 1e-8 or g['weight_decay'] != 0 or g.get('amsgrad', False):
            return False, 'Adam hyper-parameters differ from the defaults the driver implements'
        if not use_vposer or vposer is None or pose_embedding is None:
            return False, 'no VPoser embedding'
        if spec['create_graph']:
            return False, 'create_graph=True'
        loss = spec['loss']
        if not isinstance(loss, SMPLifyLoss):
            return False, 'loss is not SMPLifyLoss'
        ok, why = loss.fusable(use_vposer=True, scan_tensor=spec['scan_tensor'])
        if not ok:
            return False, why
        cam = spec['camera']
        for name in ('rotation', 'translation', 'focal_length_x', 'focal_length_y', 'center'):
            v = getattr(cam, name)
            if v.requires_grad or not bool((v == v[0:1]).all()):
                return False, 'camera.%s is optimised or varies over the batch' % name
        want = {id(p) for p in body_model.parameters() if p.requires_grad and p is not getattr(body_model, 'body_pose', None)}
        want.add(id(pose_embedding))
        have = {id(p) for p in params}
        if not want.issubset(have | {id(getattr(body_model, 'body_pose', None))}):
            return False, 'a body parameter is not being optimised'
        if getattr(body_model, 'betas').requires_grad:
            return False, 'betas are optimised'
        return True, ''

    def _fitter(self, spec, body_model, pose_embedding, vposer):
        loss, cam = spec['loss'], spec['camera']
        B = pose_embedding.shape[0]
        dev = pose_embedding.device
        jm = getattr(body_model.joint_mapper, 'joint_maps', None) if body_model.joint_mapper is not None else None
        key = (id(body_model), id(vposer), B, dev.index, None if jm is None else tuple(int(i) for i in jm.tolist()))
        fit = loss._fused.get(key)
        if fit is None:
            fit = ProxFitter(body_model, vposer, loss.motion_smooth_model if loss.use_motion_smooth_prior else None, B, dev, joint_map=jm,
                             camera=(cam.rotation[0], cam.translation[0], float(cam.focal_length_x[0]), float(cam.focal_length_y[0]),
                                     (float(cam.center[0, 0]), float(cam.center[0, 1]))),
                             cam2world=(loss.R, loss.t), sdf=getattr(loss, 'sdf', None), grid_min=getattr(loss, 'grid_min', None),
                             grid_max=getattr(loss, 'grid_max', None),
                             fric_ids=loss.contact_fric_verts_ids if loss.use_friction else None,
                             contact_ids=loss.contact_verts_ids if loss.contact else None, markers81=loss.smooth_marker_ids,
                             scene_v=spec['scene_v'] if loss.contact else None,
                             smooth_stats=(loss.Xmean_global_markers, loss.Xstd_global_markers) if loss.use_motion_smooth_prior else None,
                             weights=loss.weight_dict(), use_joints_conf=loss.use_joints_conf, sdf_penetration=loss.sdf_penetration,
                             use_friction=loss.use_friction, contact=loss.contact, use_motion_smooth_prior=loss.use_motion_smooth_prior)
            loss._fused[key] = fit
        return fit

    def _run_fused(self, optimizer, spec, body_model, pose_embedding, vposer):
        loss = spec['loss']
        fit = self._fitter(spec, body_model, pose_embedding, vposer)
        B = pose_embedding.shape[0]
        erase_n = 0 if spec['first_batch_flag'] else int(B * 0.15)
        fit.set_weights(loss.weight_dict(), erase_n, loss.use_joints_conf)
        names = ['transl', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression', 'betas']
        P = {k: getattr(body_model, k) for k in names}
        P['pose_embedding'] = pose_embedding
        fit.set_window(P, spec['gt_joints'], spec['joints_conf'], spec['joint_weights'])
        lr = float(optimizer.param_groups[0]['lr'])
        writer = spec['writer']
        chunk = self.maxiters if writer is None else max(1, self.log_every)
        done, final = 0, None
        while done < self.maxiters:
            n = min(chunk, self.maxiters - done)
            fit.run(n, lr, resume=done > 0)        # one call for the whole run unless a writer wants intermediate scalars
            done += n
            self.steps += n
            if writer is not None or done >= self.maxiters:
                ld = fit.losses()
                final = ld['total_loss']
                if writer is not None:
                    for k in ('total_loss', 'joint_loss', 'sdf_penetration_loss', 'contact_loss', 'motion_prior_smooth_loss',
                              'loss_fric_tangent', 'loss_fric_normal'):
                        writer.add_scalar('optimize/' + k, float(ld[k]), self.steps - 1)
                if bool(torch.isnan(final).sum() > 0):
                    print('NaN loss value, stopping!')
                    break
                if bool(torch.isinf(final).sum() > 0):
                    print('Infinite loss value, stopping!')
                    break
        out = fit.params()
        with torch.no_grad():
            for k, v in out.items():
                tgt = pose_embedding if k == 'pose_embedding' else getattr(body_model, k)
                tgt.copy_(v.reshape(tgt.shape))
        return None if final is None else float(final)

    # ------------------------------------------------------------------------------------------------ closure (:220-313)
    def create_fitting_closure(self, optimizer, body_model, camera=None, gt_joints=None, loss=None, joints_conf=None,
                               marker_mask=None, joint_weights=None, return_verts=True, return_full_pose=False, use_vposer=False,
                               vposer=None, pose_embedding=None, scan_tensor=None, scan_point_num=None, scene_v=None,
                               create_graph=False, writer=None, first_batch_flag=None, **kwargs):
        faces_tensor = body_model.faces_tensor.view(-1)
        append_wrists = self.model_type == 'smpl' and use_vposer

        def fitting_func(backward=True):
            if backward:
                optimizer.zero_grad()
            body_pose = vposer.decode(pose_embedding, output_type='aa').view(pose_embedding.shape[0], -1) if use_vposer else None
            if append_wrists:
                wrist_pose = torch.zeros([body_pose.shape[0], 6], dtype=body_pose.dtype, device=body_pose.device)
                body_pose = torch.cat([body_pose, wrist_pose], dim=1)
            # ONE SMPL-X evaluation serves both the mapped (OpenPose) joints and the raw SMPL-X joints: the reference calls the model
            # twice with identical arguments, toggling joint_mapper in between (:248-258)
            joint_mapper = body_model.joint_mapper
            body_model.joint_mapper = None
            raw = body_model(return_verts=True, body_pose=body_pose, return_full_pose=True)
            body_model.joint_mapper = joint_mapper
            smplx_joints = raw.joints
            body_model_output = raw._replace(joints=joint_mapper(raw.joints) if joint_mapper is not None else raw.joints)
            loss_dict = loss(body_model=body_model, body_model_output=body_model_output, smplx_joints=smplx_joints, camera=camera,
                             gt_joints=gt_joints, body_model_faces=faces_tensor, joints_conf=joints_conf, marker_mask=marker_mask,
                             joint_weights=joint_weights, pose_embedding=pose_embedding, use_vposer=use_vposer,
                             scan_tensor=scan_tensor, scan_point_num=scan_point_num, scene_v=scene_v, opt_step=self.steps, **kwargs)
            if backward:
                loss_dict['total_loss'].backward(create_graph=create_graph)
            # bs=100: erase gradient for the first 15 frames (:281-288)
            bs = smplx_joints.shape[0]
            erase_n = int(bs * 0.15)
            if not first_batch_flag:
                for body_param in body_model.parameters():
                    if body_param.grad is not None:
                        body_param.grad[0:erase_n, :] = 0
                if pose_embedding is not None and pose_embedding.grad is not None:
                    pose_embedding.grad[0:erase_n, :] = 0
            if writer is not None:
                for k in ('total_loss', 'joint_loss', 's2m_dist', 'm2s_dist', 'self_penetration_loss', 'sdf_penetration_loss',
                          'contact_loss', 'smooth_acc_loss', 'smooth_vel_loss', 'motion_prior_smooth_loss', 'loss_fric_tangent',
                          'loss_fric_normal', 'motion_infill_loss', 'motion_infill_contact_loss'):
                    writer.add_scalar('optimize/' + k, loss_dict[k].item(), self.steps)
            self.steps += 1
            fitting_func.last_loss_dict = loss_dict
            return loss_dict['total_loss']

        fitting_func.lemo_spec = dict(loss=loss, camera=camera, gt_joints=gt_joints, joints_conf=joints_conf, joint_weights=joint_weights,
                                      scene_v=scene_v, scan_tensor=scan_tensor, first_batch_flag=first_batch_flag, writer=writer,
                                      create_graph=create_graph, use_vposer=use_vposer)
        return fitting_func
