"""GPU parity of the PROX scene operators against the reference's own torch expressions evaluated on CPU
(temp_prox/camera.py:93-116, fitting_temp_slide.py:673-694 with F.grid_sample)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_common import DEV, rel

pytestmark = pytest.mark.gpu


def _ref_project(points, R, t, fx, fy, c):
    # the reference's einsum formulation (camera.py:100-116)
    B = points.shape[0]
    T = torch.zeros(B, 4, 4, dtype=points.dtype)
    T[:, :3, :3] = R; T[:, :3, 3] = t; T[:, 3, 3] = 1
    ph = torch.cat([points, torch.ones(*points.shape[:-1], 1, dtype=points.dtype)], -1)
    proj = torch.einsum('bki,bji->bjk', T, ph)
    img = proj[:, :, :2] / proj[:, :, 2:3]
    cam = torch.zeros(B, 2, 2, dtype=points.dtype); cam[:, 0, 0] = fx; cam[:, 1, 1] = fy
    return torch.einsum('bki,bji->bjk', cam, img) + c.unsqueeze(1)


def test_perspective_camera_matches_reference_expression():
    from lemo_b200.temp_prox.camera import PerspectiveCamera
    g = torch.Generator().manual_seed(0)
    B, J = 7, 118
    pts = torch.randn(B, J, 3, generator=g) * 0.5 + torch.tensor([0., 0., 3.])
    from oracle import ref_body as rb
    R = rb.rodrigues(torch.tensor([[0.1, -0.2, 0.05]]))[0]
    t = torch.tensor([0.02, -0.01, 0.1])
    fx, fy, c = 1060.53, 1060.38, torch.tensor([951.30, 536.77])
    cam = PerspectiveCamera(rotation=R[None].repeat(B, 1, 1), translation=t[None].repeat(B, 1), focal_length_x=fx, focal_length_y=fy,
                            batch_size=B, center=c[None].repeat(B, 1)).to(DEV)
    w = torch.randn(B, J, 2, generator=g)
    p64 = pts.double().requires_grad_(True)
    (_ref_project(p64, R.double().expand(B, 3, 3), t.double().expand(B, 3), fx, fy, c.double().expand(B, 2)) * w.double()).sum().backward()
    pg = pts.to(DEV).requires_grad_(True)
    out = cam(pg)
    (out * w.to(DEV)).sum().backward()
    ref = _ref_project(pts, R.expand(B, 3, 3), t.expand(B, 3), fx, fy, c.expand(B, 2))
    assert rel(out, ref) < 2e-6
    assert rel(pg.grad, p64.grad) < 1e-5


def test_sdf_sample_matches_grid_sample():
    from lemo_b200.temp_prox.camera import sdf_sample, cam_to_world
    g = torch.Generator().manual_seed(1)
    D, B, V = 32, 3, 500
    sdf = torch.randn(D, D, D, generator=g)
    gmin, gmax = torch.tensor([-1.0, -2.0, -0.5]), torch.tensor([2.0, 1.5, 2.5])
    pts = torch.rand(B, V, 3, generator=g) * (gmax - gmin) * 1.2 + gmin - 0.1 * (gmax - gmin)     # some points outside -> border clamp
    w = torch.randn(B, V, generator=g)

    def ref(p, dtype):
        norm = (p - gmin.to(dtype)) / (gmax - gmin).to(dtype) * 2 - 1
        vol = sdf.to(dtype)[None, None].expand(B, 1, D, D, D)
        return F.grid_sample(vol, norm[:, :, [2, 1, 0]].view(-1, V, 1, 1, 3), padding_mode='border', align_corners=False).view(B, V)

    p64 = pts.double().requires_grad_(True)
    (ref(p64, torch.float64) * w.double()).sum().backward()
    pg = pts.to(DEV).requires_grad_(True)
    val = sdf_sample(sdf.to(DEV), pg, gmin, gmax)
    (val * w.to(DEV)).sum().backward()
    assert rel(val, ref(pts, torch.float32)) < 1e-5
    assert rel(pg.grad, p64.grad) < 1e-4
    # penetration term of fitting_temp_slide.py:688-694 on top of the lookup
    pen = val[val < 0].abs().sum()
    assert abs(float(pen) - float(ref(pts, torch.float32)[ref(pts, torch.float32) < 0].abs().sum())) < 1e-3 * float(pen)
    # camera -> world transform and its adjoint
    from oracle import ref_body as rb
    R = rb.rodrigues(torch.tensor([[0.3, 0.1, -0.2]]))[0]
    t = torch.tensor([0.5, -0.2, 1.0])
    q = pts.to(DEV).requires_grad_(True)
    out = cam_to_world(q, R, t)
    (out * pts.to(DEV)).sum().backward()
    assert rel(out, torch.matmul(R, pts.permute(0, 2, 1)).permute(0, 2, 1) + t) < 2e-6
    assert rel(q.grad, torch.matmul(R.t(), pts.permute(0, 2, 1)).permute(0, 2, 1)) < 2e-6
