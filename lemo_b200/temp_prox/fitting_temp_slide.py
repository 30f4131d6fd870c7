"""PROX stage-2 loss and fitting loop on the lemo_b200 operators -- mirror of the reference's
temp_prox/fitting_temp_slide.py (SMPLifyLoss.forward :564-1062, FittingMonitor.run_fitting :169-313) for the terms active in
cfg_files/PROXD_temp_S2.yaml (+ the Chamfer `contact` term).

Every heavy operator is a lemo_b200 CUDA kernel behind the C ABI: full-mesh SMPL-X forward/backward, VPoser decode, camera
projection, camera->world transform, SDF trilinear lookup, Enc forward/input-gradient, Chamfer.  What remains in PyTorch here is
the reference's own elementwise glue (weighted means, L2 priors, masked means) -- kept on device without the reference's
`.item()` host syncs; fusing it into a `lemo_fit_prox_run` driver like the AMASS stages is listed in DESIGN.md section 7.
"""
import torch
import torch.nn.functional as F

from .camera import cam_to_world, sdf_sample
from .dist_chamfer import chamferDist


def _masked_mean(x, mask):
    """`x[mask].abs().mean()` if any(mask) else 0 -- on device, no host sync."""
    m = mask.to(x.dtype)
    n = m.sum()
    return torch.where(n > 0, (x.abs() * m).sum() / n.clamp(min=1.0), torch.zeros((), dtype=x.dtype, device=x.device))


class SMPLifyLoss(torch.nn.Module):
    def __init__(self, weights, camera, cam2world, sdf, grid_min, grid_max, fric_ids, contact_ids, scene_v, smooth_marker_ids,
                 smooth_enc, Xmean, Xstd, joint_weights):
        super().__init__()
        self.w = dict(weights)
        self.camera = camera
        self.R, self.t = cam2world
        self.sdf, self.grid_min, self.grid_max = sdf, grid_min, grid_max
        self.fric_ids, self.contact_ids, self.scene_v = fric_ids, contact_ids, scene_v
        self.smooth_marker_ids, self.enc = smooth_marker_ids, smooth_enc
        self.Xmean, self.Xstd = Xmean, Xstd
        self.joint_weights = joint_weights
        self.chamfer = chamferDist()

    def forward(self, body_model_output, smplx_joints, gt_joints, joints_conf, pose_embedding):
        w, out = self.w, body_model_output
        dev = out.vertices.device
        T = {}
        proj = self.camera(out.joints)
        wts = (self.joint_weights * joints_conf).unsqueeze(-1)
        T['joint'] = torch.mean(wts ** 2 * torch.abs(gt_joints - proj)) * w['data']
        T['pprior'] = pose_embedding.pow(2).sum() * w['body_pose'] ** 2
        idx = torch.tensor([55, 58, 12, 15], device=dev) - 3
        sgn = torch.tensor([1., -1., -1., -1.], device=dev)
        T['angle'] = torch.sum(torch.exp(out.full_pose[:, 3:66][:, idx] * sgn)) * (3.17 * w['body_pose']) ** 2
        T['hand'] = (out.left_hand_pose.pow(2).sum() + out.right_hand_pose.pow(2).sum()) * w['hand_prior'] ** 2
        T['expr'] = out.expression.pow(2).sum() * w['expr'] ** 2
        T['jaw'] = (out.jaw_pose * w['jaw']).pow(2).sum()
        vw = cam_to_world(out.vertices, self.R, self.t)
        jw = cam_to_world(smplx_joints, self.R, self.t)
        body_sdf = sdf_sample(self.sdf, vw, self.grid_min, self.grid_max)                    # [B,V]
        T['sdf'] = w['sdf'] * torch.clamp(-body_sdf, min=0).sum()
        fr = vw[:, self.fric_ids]
        vel = fr[1:] - fr[:-1]
        sel = body_sdf[:-1][:, self.fric_ids] < 0.01
        vn = vel[..., 2]
        vt = torch.sqrt(vel[..., 0] ** 2 + vel[..., 1] ** 2 + 1e-30)
        T['fric_t'] = _masked_mean(vt, sel & (vt > 1e-4)) * w['fric_t']
        T['fric_n'] = _masked_mean(vn, sel & (vn < 0)) * w['fric_n']
        T['contact'] = torch.zeros((), device=dev)
        if w.get('contact', 0) > 0:
            d1, _, _, _ = self.chamfer(vw[:, self.contact_ids].contiguous(), self.scene_v[None])     # shared scene, not replicated
            r = torch.sqrt(d1 + 1e-4)
            T['contact'] = w['contact'] * (r / (r + 1.0)).mean()
        m = vw[:, self.smooth_marker_ids]
        j0 = jw[0].detach()
        x_axis = j0[2] - j0[1]
        x_axis = torch.stack([x_axis[0], x_axis[1], torch.zeros((), device=dev)])
        x_axis = x_axis / torch.norm(x_axis)
        z_axis = torch.tensor([0., 0., 1.], device=dev)
        y_axis = torch.cross(z_axis, x_axis, dim=0)
        y_axis = y_axis / torch.norm(y_axis)
        Rt = torch.stack([x_axis, y_axis, z_axis], 1)
        g = torch.matmul(m - m[0].detach()[0], Rt)
        img = ((g.reshape(g.shape[0], -1).unsqueeze(0) - self.Xmean) / self.Xstd).permute(0, 2, 1).unsqueeze(1)
        v = F.pad(img[:, :, :, 1:] - img[:, :, :, :-1], (8, 8, 1, 1), 'reflect')
        z = self.enc(v)[0]
        T['smooth'] = torch.mean((z[..., 1:] - z[..., :-1]) ** 2) * w['smooth']
        total = sum(T.values())
        return total, T


def create_loss(loss_type='smplify', **kwargs):
    """fitting_temp_slide.py:316-322: 'smplify' -> SMPLifyLoss (the 'camera_init' loss belongs to PROX stage 1 and is not on this path)."""
    if loss_type == 'smplify':
        return SMPLifyLoss(**kwargs)
    raise ValueError('Unknown loss type: {}'.format(loss_type))


class FittingMonitor:
    """run_fitting (:169-313) for the Adam branch: `maxiters` closure steps (:196-197), NaN/Inf stop (:197-203), first-15 % gradient
    erase (:281-288).

    use_cuda_graph=True (EXPERIMENTAL, written after the round's GPU budget was spent -- not yet run on a GPU; default off): the whole
    step (closure, backward, gradient erase, optimizer.step) is captured once with torch.cuda.graphs after three eager warm-up steps and
    replayed; every lemo_b200 operator enqueues on torch's current stream, so it is captured like a torch op.  Needs an optimizer built
    with `capturable=True` and a closure free of host syncs (SMPLifyLoss here is)."""

    def __init__(self, maxiters=900, erase_first=False, use_cuda_graph=False, check_every=50):
        self.maxiters, self.erase_first, self.use_cuda_graph = maxiters, erase_first, use_cuda_graph
        self.check_every = max(1, int(check_every))

    def _diverged(self, loss, n):
        """The reference tests the loss on the host after EVERY step (:197-203, one sync per iteration) and stops; here the test runs
        every `check_every` steps and on the last one -- once the loss is NaN/Inf the parameters already are, so the result is the same
        and the steady state has no host sync.  Same messages."""
        if (n + 1) % self.check_every and n + 1 != self.maxiters:
            return False
        if bool(torch.isnan(loss).sum() > 0):
            print('NaN loss value, stopping!')
            return True
        if bool(torch.isinf(loss).sum() > 0):
            print('Infinite loss value, stopping!')
            return True
        return False

    def _step(self, optimizer, closure, params):
        loss = closure()
        loss.backward()
        if self.erase_first:
            for p in params:
                if p.grad is not None:
                    p.grad[0:int(p.shape[0] * 0.15)] = 0
        optimizer.step()
        return loss

    def run_fitting(self, optimizer, closure, params):
        loss = None
        if self.use_cuda_graph and self.maxiters > 3:
            if not optimizer.defaults.get('capturable', False):
                raise RuntimeError('use_cuda_graph=True needs an optimizer created with capturable=True')
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):                                  # eager warm-up steps (they count towards maxiters)
                    optimizer.zero_grad(set_to_none=True)
                    loss = self._step(optimizer, closure, params)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                loss = self._step(optimizer, closure, params)       # gradients live in the graph's private pool: static addresses
            for n in range(4, self.maxiters):
                graph.replay()
                if self._diverged(loss, n):
                    break
        else:
            for n in range(self.maxiters):
                optimizer.zero_grad()
                loss = self._step(optimizer, closure, params)
                if self._diverged(loss, n):
                    break
        return loss
