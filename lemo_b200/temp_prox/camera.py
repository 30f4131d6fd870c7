"""Drop-in for the reference's temp_prox/camera.py `PerspectiveCamera` (lines 42-116) and the scene look-ups of
temp_prox/fitting_temp_slide.py:673-694 (camera->world transform, SDF grid_sample)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib


def _host(a, n):
    return np.ascontiguousarray(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, np.float32).reshape(n))


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, R, t, fx, fy, cx, cy):
        p = points.contiguous().float()
        n = p.numel() // 3
        out = torch.empty(*p.shape[:-1], 2, device=p.device)
        _lib.call('lemo_camera_project', _lib.ptr(p), n, _lib.ptr(R), _lib.ptr(t), fx, fy, cx, cy, _lib.ptr(out), _lib.cur_stream(p.device))
        ctx.save_for_backward(p)
        ctx.cam = (R, t, fx, fy, cx, cy)
        return out

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        R, t, fx, fy, cx, cy = ctx.cam
        dp = torch.empty_like(p)
        _lib.call('lemo_camera_project_backward', _lib.ptr(p), p.numel() // 3, _lib.ptr(R), _lib.ptr(t), fx, fy, cx, cy,
                  _lib.ptr(g.contiguous().float()), _lib.ptr(dp), _lib.cur_stream(p.device))
        return dp, None, None, None, None, None, None


class PerspectiveCamera(nn.Module):
    """Fixed pinhole camera (S2/S3 configs use camera_mode 'fixed': rotation/translation are not optimised)."""
    FOCAL_LENGTH = 5000

    def __init__(self, rotation=None, translation=None, focal_length_x=None, focal_length_y=None, batch_size=1, center=None,
                 dtype=torch.float32, **kwargs):
        super().__init__()
        self.batch_size = batch_size
        fx = self.FOCAL_LENGTH if focal_length_x is None else focal_length_x
        fy = self.FOCAL_LENGTH if focal_length_y is None else focal_length_y
        self.register_buffer('focal_length_x', torch.full([batch_size], float(fx), dtype=dtype))
        self.register_buffer('focal_length_y', torch.full([batch_size], float(fy), dtype=dtype))
        self.register_buffer('center', torch.zeros([batch_size, 2], dtype=dtype) if center is None else center)
        rot = torch.eye(3, dtype=dtype).unsqueeze(0).repeat(batch_size, 1, 1) if rotation is None else rotation
        self.rotation = nn.Parameter(rot, requires_grad=False)
        tr = torch.zeros([batch_size, 3], dtype=dtype) if translation is None else translation
        self.translation = nn.Parameter(tr, requires_grad=False)

    def forward(self, points):
        if points.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        R, t = _host(self.rotation[0], 9), _host(self.translation[0], 3)
        c = self.center[0].detach().cpu()
        return _Project.apply(points, R, t, float(self.focal_length_x[0]), float(self.focal_length_y[0]), float(c[0]), float(c[1]))


class _Rigid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, R, t):
        p = points.contiguous().float()
        out = torch.empty_like(p)
        _lib.call('lemo_rigid_transform', _lib.ptr(p), p.numel() // 3, _lib.ptr(R), _lib.ptr(t), 0, _lib.ptr(out), _lib.cur_stream(p.device))
        ctx.R = R
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float()
        dp = torch.empty_like(g)
        _lib.call('lemo_rigid_transform', _lib.ptr(g), g.numel() // 3, _lib.ptr(ctx.R), None, 1, _lib.ptr(dp), _lib.cur_stream(g.device))
        return dp, None, None


def cam_to_world(points, R, t):
    """vertices_world = (R @ v^T)^T + t  (fitting_temp_slide.py:677)."""
    return _Rigid.apply(points, _host(R, 9), _host(t, 3))


class _SdfSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, sdf, gmin, gmax):
        p = points.contiguous().float()
        dim = sdf.shape[-1]
        val = torch.empty(p.shape[:-1], device=p.device)
        _lib.call('lemo_sdf_sample', _lib.ptr(p), p.numel() // 3, _lib.ptr(sdf), dim, _lib.ptr(gmin), _lib.ptr(gmax), _lib.ptr(val),
                  _lib.cur_stream(p.device))
        ctx.save_for_backward(p, sdf)
        ctx.g = (gmin, gmax, dim)
        return val

    @staticmethod
    def backward(ctx, gval):
        p, sdf = ctx.saved_tensors
        gmin, gmax, dim = ctx.g
        dp = torch.empty_like(p)
        _lib.call('lemo_sdf_sample_backward', _lib.ptr(p), p.numel() // 3, _lib.ptr(sdf), dim, _lib.ptr(gmin), _lib.ptr(gmax),
                  _lib.ptr(gval.contiguous().float()), _lib.ptr(dp), _lib.cur_stream(p.device))
        return dp, None, None, None


def one_volume(sdf):
    """[D,D,D] view of an SDF given as [D,D,D], [bs,D,D,D] or [bs,1,D,D,D] (the reference's caller repeats ONE scene volume bs times,
    fit_temp_loadprox_slide.py:297-299: the first replica is the scene)."""
    while sdf.dim() > 3:
        sdf = sdf[0]
    return sdf


def sdf_sample(sdf, vertices_world, grid_min, grid_max):
    """body_sdf of fitting_temp_slide.py:682-687 as [B, V]: trilinear, border padding, align_corners=False.  `sdf` is ONE
    [dim,dim,dim] volume on the device (the reference repeats it B times), indexed [x][y][z]."""
    sdf = one_volume(sdf).contiguous().float()
    return _SdfSample.apply(vertices_world, sdf, _host(grid_min, 3), _host(grid_max, 3))
