"""CPU restatement of the reference's Adam fitting loops.  TEST INFRASTRUCTURE ONLY.

  temporal stage   /root/reference/opt_amass_temp.py:329-455   (B=T, marker L1 + smoothness prior +
                                                                contact-velocity + 3 L2 priors)
  per-frame stage  /root/reference/opt_amass_perframe.py:293-361 (B=1, warm start frame to frame)

The op sequence deliberately keeps the reference's redundancies (SMPL-X + VPoser evaluated twice per
iteration, 6D -> R -> aa -> Rodrigues round trip, eager autograd, torch.optim.Adam) because this
module is also the timed CPU baseline (bench.py cpu_baseline / --impl reference).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ref_body as rb
from . import ref_priors as rp


class FitContext:
    """Everything the loops need that is constant over a run."""

    def __init__(self, model_np, vposer_w, enc_w, tables, dtype=torch.float32):
        self.dtype = dtype
        self.smplx = rb.SMPLXRef(model_np, dtype=dtype)
        self.vposer = rb.VPoserRef(vposer_w, dtype=dtype)
        self.enc_sd = {k: torch.from_numpy(v).to(dtype) for k, v in enc_w.items()}
        self.m67 = torch.from_numpy(tables['markers67']).long()
        self.m81 = torch.from_numpy(tables['markers81']).long()
        self.foot = [torch.from_numpy(tables[k]).long() for k in ('left_heel', 'right_heel', 'left_toe', 'right_toe')]
        self.Xmean = torch.from_numpy(tables['smooth_Xmean']).to(dtype).view(1, 1, 243)
        self.Xstd = torch.from_numpy(tables['smooth_Xstd']).to(dtype)


W_TEMP = dict(rec=1.0, contact=0.03, smooth=1e6, vposer=0.02, shape=0.01, hand=0.01)   # opt_amass_temp.py:46-51


def smooth_input(markers81, joints0, ctx):
    """opt_amass_temp.py:366-387: canonical frame (detached), normalise, temporal diff, reflect pad."""
    j0 = joints0.detach()
    x_axis = j0[2] - j0[1]
    x_axis = torch.stack([x_axis[0], x_axis[1], torch.zeros((), dtype=x_axis.dtype)])
    x_axis = x_axis / torch.norm(x_axis)
    z_axis = torch.tensor([0., 0., 1.], dtype=x_axis.dtype)
    y_axis = torch.cross(z_axis, x_axis, dim=0)
    y_axis = y_axis / torch.norm(y_axis)
    Rt = torch.stack([x_axis, y_axis, z_axis], 1)
    g = torch.matmul(markers81 - markers81[0].detach()[0], Rt)
    img = g.reshape(g.shape[0], -1).unsqueeze(0)
    img = (img - ctx.Xmean) / ctx.Xstd
    img = img.permute(0, 2, 1).unsqueeze(1)
    v = img[:, :, :, 1:] - img[:, :, :, :-1]
    return F.pad(v, (8, 8, 1, 1), 'reflect')


def contact_vel_loss(verts, contact, ctx, thres=0.1):
    """opt_amass_temp.py:407-447."""
    vel = (verts[1:] - verts[:-1]) * 30
    total = torch.zeros((), dtype=verts.dtype)
    for part in range(4):
        sel = vel[:, ctx.foot[part], :][contact[:-1, part] == 1]
        nrm = torch.norm(sel, dim=-1)
        if (nrm - thres).gt(0).sum().item() >= 1:
            total = total + nrm[nrm > thres].abs().mean()
    return total


def temp_losses(transl, rot6d, other, shape, markers_rec, contact, ctx, w=W_TEMP, faithful=True):
    """One forward of opt_amass_temp.py:355-449.  Returns (loss, dict of terms, params72)."""
    x75 = torch.cat([transl, rot6d, shape, other], -1)
    p72 = rb.convert_to_3D_rot(x75)
    verts, joints = rb.gen_body_mesh(p72, ctx.smplx, ctx.vposer)
    if faithful:                                            # second SMPL-X + VPoser evaluation (:364)
        _, joints = rb.gen_body_mesh(p72, ctx.smplx, ctx.vposer)
    m67 = verts[:, ctx.m67]
    terms = {}
    terms['rec'] = F.l1_loss(m67, markers_rec)
    terms['vposer'] = torch.mean(p72[:, 16:48] ** 2)
    terms['shape'] = torch.mean(p72[:, 6:16] ** 2)
    terms['hand'] = torch.mean(p72[:, 48:] ** 2)
    if w.get('smooth', 0) > 0:
        xin = smooth_input(verts[:, ctx.m81], joints[0], ctx)
        z = rp.enc_forward(xin, ctx.enc_sd)
        terms['smooth'] = torch.mean((z[..., 1:] - z[..., :-1]) ** 2)
    else:
        terms['smooth'] = torch.zeros((), dtype=verts.dtype)
    if w.get('contact', 0) > 0:
        terms['contact'] = contact_vel_loss(verts, contact, ctx)
    else:
        terms['contact'] = torch.zeros((), dtype=verts.dtype)
    loss = sum(w[k] * terms[k] for k in ('rec', 'vposer', 'shape', 'hand', 'contact', 'smooth'))
    return loss, terms, p72


def split_init(init72, dtype=torch.float32):
    """opt_amass_temp.py:332-341: [T,72] -> transl, rot6d (via tgm aa->R), shape, other(56)."""
    p = torch.from_numpy(np.asarray(init72)).to(dtype)
    return p[:, 0:3].clone(), rb.convert_to_6D_all(p[:, 3:6]), p[:, 6:16].clone(), p[:, 16:].clone()


def fit_temp(init72, markers_rec, contact, ctx, n_iters=100, lr0=0.01, lr1=0.005, lr_switch=60,
             w=W_TEMP, faithful=True, trace=None):
    """opt_amass_temp.py:343-455.  Returns (params72 of the LAST forward, dict(final raw params))."""
    transl, rot6d, shape, other = split_init(init72, ctx.dtype)
    for t in (transl, rot6d, other):
        t.requires_grad_(True)
    mrec = torch.from_numpy(np.asarray(markers_rec)).to(ctx.dtype)
    con = torch.from_numpy(np.asarray(contact)).to(ctx.dtype)
    opt = torch.optim.Adam([transl, rot6d, other], lr=lr0)
    p72 = None
    for step in range(n_iters):
        if step > lr_switch:
            for g in opt.param_groups:
                g['lr'] = lr1
        opt.zero_grad()
        loss, terms, p72 = temp_losses(transl, rot6d, other, shape, mrec, con, ctx, w, faithful)
        loss.backward()
        if trace is not None:
            trace.append({'loss': float(loss), **{k: float(v) for k, v in terms.items()},
                          'g_transl': transl.grad.clone(), 'g_rot6d': rot6d.grad.clone(), 'g_other': other.grad.clone()})
        opt.step()
    return p72.detach().numpy(), dict(transl=transl.detach().numpy(), rot6d=rot6d.detach().numpy(),
                                      other=other.detach().numpy())


W_PF = dict(rec=1.0, vposer=0.02, shape=0.01, hand=0.01)     # opt_amass_perframe.py:40-43


def perframe_losses(transl, rot6d, other, shape, markers_rec_t, ctx, w=W_PF):
    """opt_amass_perframe.py:332-353."""
    p72 = rb.convert_to_3D_rot(torch.cat([transl, rot6d, shape, other], -1))
    verts, _ = rb.gen_body_mesh(p72, ctx.smplx, ctx.vposer)
    loss = (w['rec'] * F.l1_loss(verts[:, ctx.m67], markers_rec_t) + w['vposer'] * torch.mean(p72[:, 16:48] ** 2)
            + w['shape'] * torch.mean(p72[:, 6:16] ** 2) + w['hand'] * torch.mean(p72[:, 48:] ** 2))
    return loss, p72


def fit_perframe(markers_rec, betas, ctx, n_frames=None, n_iters=100, trace=None):
    """opt_amass_perframe.py:293-361: T sequential B=1 problems, warm-started, fresh Adam per frame,
    lr .1 (frame 0) / .01, ->.01 @step>60, ->.003 @step>80.  Returns [T,72]."""
    dt = ctx.dtype
    mrec = torch.from_numpy(np.asarray(markers_rec)).to(dt)
    T = mrec.shape[0] if n_frames is None else n_frames
    shape = torch.from_numpy(np.asarray(betas)).to(dt).view(1, 10)
    transl = torch.tensor([[0., 0.4, 1.0]], dtype=dt)
    rot6d = rb.convert_to_6D_all(torch.tensor([[0., 1.6, 3.14]], dtype=dt))
    other = torch.zeros(1, 56, dtype=dt)
    for t_ in (transl, rot6d, other):
        t_.requires_grad_(True)
    out = []
    for t in range(T):
        opt = torch.optim.Adam([transl, rot6d, other], lr=0.1 if t == 0 else 0.01)
        for step in range(n_iters):
            if step > 60:
                for g in opt.param_groups:
                    g['lr'] = 0.01
            if step > 80:
                for g in opt.param_groups:
                    g['lr'] = 0.003
            opt.zero_grad()
            loss, p72 = perframe_losses(transl, rot6d, other, shape, mrec[t:t + 1], ctx)
            loss.backward()
            if trace is not None:
                trace.append(float(loss))
            opt.step()
        out.append(p72[0].detach().numpy())
    return np.asarray(out)
