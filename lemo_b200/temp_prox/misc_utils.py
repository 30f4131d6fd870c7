"""Mirror of the pieces of the reference's temp_prox/misc_utils.py that sit on the PROX fitting path: the keypoint robustifier
(GMoF, misc_utils.py:61-85), the joint mapper (:45-58) and the SMPL-X -> OpenPose joint map (:87-197).

Elementwise glue on whatever device the caller's tensors live on (the heavy operators of the PROX loss are the lemo_b200 kernels,
see fitting_temp_slide.py here).  The joint maps are DATA: they are read from assets/prox_tables.npz, which tools/export_assets.py
wrote by calling the reference's own smpl_to_openpose for every flag combination -- so they are equal to the reference's by
construction, and tests/test_prox_tables.py checks the invariants the reference's loss relies on.
"""
import os

import numpy as np
import torch
import torch.nn as nn

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'assets')
_TABLES = None


def prox_tables():
    """OpenPose joint maps + the friction (307) / contact (1121) vertex-id lists of fit_temp_loadprox_slide.py:349-362."""
    global _TABLES
    if _TABLES is None:
        _TABLES = dict(np.load(os.path.join(_ASSETS, 'prox_tables.npz')))
    return _TABLES


class GMoF(nn.Module):
    """Geman-McClure robustifier, scaled: rho^2 * r^2 / (r^2 + rho^2)."""

    def __init__(self, rho=1):
        super().__init__()
        self.rho = rho

    def extra_repr(self):
        return 'rho = {}'.format(self.rho)

    def forward(self, residual):
        sq = residual ** 2
        return self.rho ** 2 * torch.div(sq, sq + self.rho ** 2)


class GMoF_unscaled(nn.Module):
    """Geman-McClure robustifier without the rho^2 factor: r^2 / (r^2 + rho^2)."""

    def __init__(self, rho=1):
        super().__init__()
        self.rho = rho

    def extra_repr(self):
        return 'rho = {}'.format(self.rho)

    def forward(self, residual):
        sq = residual ** 2
        return torch.div(sq, sq + self.rho ** 2)


class JointMapper(nn.Module):
    """joints[:, joint_maps] (identity when joint_maps is None); assign an instance to `body_model.joint_mapper`."""

    def __init__(self, joint_maps=None):
        super().__init__()
        if joint_maps is None:
            self.joint_maps = None
        else:
            self.register_buffer('joint_maps', torch.as_tensor(np.asarray(joint_maps), dtype=torch.long))

    def forward(self, joints, **kwargs):
        if self.joint_maps is None:
            return joints
        return torch.index_select(joints, 1, self.joint_maps)


def smpl_to_openpose(model_type='smplx', use_hands=True, use_face=True, use_face_contour=False, openpose_format='coco25'):
    """Indices that gather the SMPL-X output joints [B,127(+17 contour),3] into OpenPose order.  Only model_type='smplx' is on the
    LEMO path (temp_prox/main_slide.py:160-179); 'coco25' -> 118 joints with hands and face, 'coco19' -> 112."""
    if model_type != 'smplx':
        raise ValueError('Unknown model type: {}'.format(model_type) if model_type not in ('smpl', 'smplh')
                         else 'only model_type="smplx" is on the LEMO fitting path')
    if openpose_format.lower() not in ('coco25', 'coco19'):
        raise ValueError('Unknown joint format: {}'.format(openpose_format))
    key = 'smplx_%s_h%d_f%d_c%d' % (openpose_format.lower(), int(bool(use_hands)), int(bool(use_face)), int(bool(use_face_contour)))
    return prox_tables()[key].copy()
