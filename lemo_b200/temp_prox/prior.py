"""Mirror of the priors of the reference's temp_prox/prior.py that the shipped PROX configurations select
(`body_prior_type: 'l2'`, `left/right_hand_prior_type: 'l2'`, `jaw/expr_prior_type: 'l2'`, plus the elbow/knee angle prior that
fit_temp_loadprox_slide.py always adds): create_prior (:33-50), SMPLifyAnglePrior (:53-89), L2Prior (:92-98).
The GMM prior (MaxMixturePrior, :100-231) needs a licensed pickle and is gated off in S2/S3 -- out of scope.
Elementwise glue on the caller's device."""
import numpy as np
import torch
import torch.nn as nn

DEFAULT_DTYPE = torch.float32


def create_prior(prior_type=None, **kwargs):
    if prior_type == 'l2':
        return L2Prior(**kwargs)
    if prior_type == 'angle':
        return SMPLifyAnglePrior(**kwargs)
    if prior_type == 'none' or prior_type is None:
        def no_prior(*args, **kwargs):
            return 0.0
        return no_prior
    if prior_type == 'gmm':
        raise ValueError('Prior gmm needs the licensed GMM pickle and is gated off in the shipped PROX configurations: not provided')
    raise ValueError('Prior {}'.format(prior_type) + ' is not implemented')


class SMPLifyAnglePrior(nn.Module):
    """exp(sign * angle) on the bending axis of left/right elbow and left/right knee (full-pose columns 55, 58, 12, 15)."""

    def __init__(self, dtype=torch.float32, **kwargs):
        super().__init__()
        self.register_buffer('angle_prior_idxs', torch.tensor(np.array([55, 58, 12, 15], dtype=np.int64), dtype=torch.long))
        self.register_buffer('angle_prior_signs', torch.tensor([1., -1., -1., -1.], dtype=dtype))

    def forward(self, pose, with_global_pose=False):
        idxs = self.angle_prior_idxs - (not with_global_pose) * 3
        return torch.exp(pose[:, idxs] * self.angle_prior_signs)


class L2Prior(nn.Module):
    def __init__(self, dtype=DEFAULT_DTYPE, reduction='sum', **kwargs):
        super().__init__()

    def forward(self, module_input, *args):
        return torch.sum(module_input.pow(2))
