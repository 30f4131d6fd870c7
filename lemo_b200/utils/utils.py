"""Drop-in for the hot-path functions of the reference's utils/utils.py: 6D <-> axis-angle conversions and the
params72 -> body mesh helpers (lines 50-169), the clip representation builder and its inverse (lines 180-265)."""
import torch

from .. import _lib


class _Rot6dToMat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x6):
        x6 = x6.contiguous().float()
        n = x6.numel() // 6
        R = torch.empty(n, 9, device=x6.device)
        _lib.call('lemo_rot6d_to_rotmat', _lib.ptr(x6), n, _lib.ptr(R), _lib.cur_stream(x6.device))
        ctx.save_for_backward(x6)
        return R.view(n, 3, 3)

    @staticmethod
    def backward(ctx, gR):
        (x6,) = ctx.saved_tensors
        n = x6.numel() // 6
        dx = torch.empty_like(x6)
        _lib.call('lemo_rot6d_to_rotmat_backward', _lib.ptr(x6), _lib.ptr(gR.contiguous().float()), n, _lib.ptr(dx),
                  _lib.cur_stream(x6.device))
        return dx


class _MatToAA(torch.autograd.Function):
    """tgm.rotation_matrix_to_angle_axis on [n,9] (utils/utils.py:74-81) with its adjoint through the selected quaternion branch."""

    @staticmethod
    def forward(ctx, R):
        R = R.contiguous().float()
        n = R.numel() // 9
        aa = torch.empty(n, 3, device=R.device)
        _lib.call('lemo_rotmat_to_aa', _lib.ptr(R), n, _lib.ptr(aa), _lib.cur_stream(R.device))
        ctx.save_for_backward(R)
        return aa

    @staticmethod
    def backward(ctx, gaa):
        (R,) = ctx.saved_tensors
        n = R.numel() // 9
        dR = torch.empty_like(R)
        _lib.call('lemo_rotmat_to_aa_backward', _lib.ptr(R), _lib.ptr(gaa.contiguous().float()), n, _lib.ptr(dR), _lib.cur_stream(R.device))
        return dR


def _no_grad_input(x, what):
    if torch.is_grad_enabled() and x.requires_grad:
        raise RuntimeError('%s: this conversion is forward-only here (the reference uses it at initialisation, without gradient); '
                           'differentiable paths: ContinousRotReprDecoder.decode / matrot2aa, convert_to_3D_rot, vposer.decode' % what)
    return x.detach()


def _ew(name, x, in_w, out_w):
    x = x.contiguous().float()
    n = x.numel() // in_w
    out = torch.empty(n, out_w, device=x.device)
    _lib.call(name, _lib.ptr(x), n, _lib.ptr(out), _lib.cur_stream(x.device))
    return out


class ContinousRotReprDecoder:
    """utils/utils.py:50-90."""

    @staticmethod
    def decode(module_input):
        return _Rot6dToMat.apply(module_input.reshape(-1, 6))

    @staticmethod
    def matrot2aa(pose_matrot):
        return _MatToAA.apply(pose_matrot.reshape(-1, 9))

    @staticmethod
    def aa2matrot(pose):
        x6 = _ew('lemo_aa_to_rot6d', _no_grad_input(pose, 'aa2matrot').reshape(-1, 3), 3, 6)
        return _Rot6dToMat.apply(x6)      # exact for a rotation's own first two columns


def convert_to_6D_all(x_batch):
    """utils/utils.py:127-130 (init only, no gradient in the reference's use)."""
    return _ew('lemo_aa_to_rot6d', _no_grad_input(x_batch, 'convert_to_6D_all').reshape(-1, 3), 3, 6)


def convert_to_3D_all(x_batch):
    return ContinousRotReprDecoder.matrot2aa(ContinousRotReprDecoder.decode(x_batch))


def convert_to_3D_rot(x_batch):
    """utils/utils.py:111-123 ([bs,75] -> [bs,72]), differentiable like the reference (6D Gram-Schmidt adjoint + tgm R->aa adjoint),
    so `gen_body_mesh_v1(convert_to_3D_rot(x75), ...)` carries the data-term gradient to the 6D rotation (opt_amass_temp.py:356-357).
    gen_body_mesh_v1(params75) is the shorter equivalent path (no R -> aa -> Rodrigues round trip)."""
    xr_aa = convert_to_3D_all(x_batch[:, 3:9])
    return torch.cat([x_batch[:, :3], xr_aa, x_batch[:, 9:]], dim=-1)


def gen_body_mesh_v1(body_params, smplx_model, vposer_model, return_joints=False):
    """utils/utils.py:141-154.  body_params [T,72] (aa global orient) or [T,75] (6D global orient, differentiable:
    the global rotation and the VPoser body rotations enter the body model as matrices, which is gradient-equivalent
    to the reference's R -> aa -> Rodrigues round trip)."""
    bs = body_params.shape[0]
    six = body_params.shape[1] == 75
    o = 3 if six else 0
    kw = dict(transl=body_params[:, 0:3], betas=body_params[:, 6 + o:16 + o],
              left_hand_pose=body_params[:, 48 + o:60 + o], right_hand_pose=body_params[:, 60 + o:72 + o])
    if six:
        kw['R_global'] = ContinousRotReprDecoder.decode(body_params[:, 3:9]).reshape(bs, 9)
    else:
        kw['global_orient'] = body_params[:, 3:6]
    kw['R_body'] = vposer_model.decode(body_params[:, 16 + o:48 + o], output_type='matrot').reshape(bs, 21, 9)
    out = smplx_model(return_verts=True, **kw)
    return out.joints if return_joints else out.vertices


def gen_body_joints_v1(body_params, smplx_model, vposer_model):
    """utils/utils.py:156-169."""
    return gen_body_mesh_v1(body_params, smplx_model, vposer_model, return_joints=True)


def get_local_markers_4chan(cur_body, contact_lbls, device='cuda'):
    """utils/utils.py:209-265 on the device: cur_body [T,1+67,3], contact_lbls [T,4] -> (cur_body [4,T-1,208], rot_0_pivot [1]).
    Un-normalised, in the function's own [channel, frame, row] order (a permuted view of the [4,208,T-1] wire layout)."""
    from ..infill import body_repr
    rep, rot0 = body_repr(cur_body, contact_lbls, stats=None, device=device)
    return rep.permute(0, 2, 1), rot0


def reconstruct_global_body(body_joints_input, rot_0_pivot, device='cuda'):
    """utils/utils.py:180-203 on the device: [T, 1+68+1, 3] (zero reference joint, local pelvis + markers, (vx, vy, r) trajectory)
    -> [T, 68, 3] world positions (float32; the heading / translation recurrence runs in double like the reference)."""
    import numpy as np
    from .. import _lib as L
    x = torch.as_tensor(np.asarray(body_joints_input) if not torch.is_tensor(body_joints_input) else body_joints_input)
    x = x.to(device, torch.float32).contiguous()
    T = x.shape[0]
    assert tuple(x.shape) == (T, 70, 3), 'expected [T, 1+68+1, 3]'
    rot0 = torch.as_tensor(np.asarray(rot_0_pivot, np.float64) if not torch.is_tensor(rot_0_pivot) else rot_0_pivot)
    rot0 = rot0.to(x.device, torch.float64).reshape(1)
    out = torch.empty(T, 68, 3, device=x.device)
    ws = torch.empty(8 * T, dtype=torch.float64, device=x.device)
    L.call('lemo_reconstruct_global_body', L.ptr(x), L.ptr(rot0), T, L.ptr(out), L.ptr(ws), L.cur_stream(x.device))
    for t in (x, rot0, ws):
        t.record_stream(torch.cuda.current_stream(x.device))
    return out
