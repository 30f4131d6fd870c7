"""Mirror of the priors of the reference's temp_prox/prior.py that the shipped PROX configurations select
(`body_prior_type: 'l2'`, `left/right_hand_prior_type: 'l2'`, `jaw/expr_prior_type: 'l2'`, plus the elbow/knee angle prior that
fit_temp_loadprox_slide.py always adds): create_prior (:33-50), SMPLifyAnglePrior (:53-89), L2Prior (:92-98).
MaxMixturePrior (:100-231, the 'gmm' body-pose prior used when VPoser is off) is mirrored too: its pickle (gmm_08.pkl) is licensed and
not shipped, so it is read from `prior_folder` like the reference does, or taken from a dict (tests use a synthetic mixture and golden
outputs of the reference class).  Elementwise / tiny-matrix glue on the caller's device."""
import os
import pickle

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
        return MaxMixturePrior(**kwargs)
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


class MaxMixturePrior(nn.Module):
    """Mixture-of-Gaussians body-pose prior ('gmm', prior.py:100-231): for every pose the negative log-likelihood under its best
    component.  The mixture comes from `{prior_folder}/gmm_{num_gaussians:02d}.pkl` (a dict or an sklearn GMM, as in the reference) or
    directly from `gmm=dict(means [M,D], covars [M,D,D], weights [M])`.  Buffers keep the reference's names (means, covs, precisions,
    nll_weights, weights, cov_dets).  Both likelihood forms are evaluated for all components at once (no Python loop over components)."""

    def __init__(self, prior_folder='prior', num_gaussians=6, dtype=DEFAULT_DTYPE, epsilon=1e-16, use_merged=True, gmm=None, **kwargs):
        super().__init__()
        if dtype not in (torch.float32, torch.float64):
            raise ValueError('Unknown float type {}'.format(dtype))
        npd = np.float32 if dtype == torch.float32 else np.float64
        self.epsilon, self.use_merged = epsilon, use_merged
        if gmm is None:
            fn = os.path.join(prior_folder, 'gmm_{:02d}.pkl'.format(num_gaussians))
            if not os.path.exists(fn):
                raise FileNotFoundError('The path to the mixture prior "{}" does not exist'.format(fn))
            with open(fn, 'rb') as f:
                gmm = pickle.load(f, encoding='latin1')
        if not isinstance(gmm, dict):                      # sklearn.mixture GMM object
            gmm = dict(means=gmm.means_, covars=gmm.covars_, weights=gmm.weights_)
        mu, cov, w = (np.asarray(gmm[k]) for k in ('means', 'covars', 'weights'))
        self.num_gaussians, self.random_var_dim = mu.shape
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        cov_t = cov.astype(npd)
        self.register_buffer('means', t(mu.astype(npd)))
        self.register_buffer('covs', t(cov_t))
        self.register_buffer('precisions', t(np.linalg.inv(cov_t).astype(npd)))
        # mixture weight over the normalisation constant, the determinant taken relative to the smallest one (:150-156)
        root_det = np.sqrt(np.linalg.det(cov))
        self.register_buffer('nll_weights', t(w / ((2 * np.pi) ** (69 / 2.) * (root_det / root_det.min()))).unsqueeze(0))
        self.register_buffer('weights', t(w).unsqueeze(0))
        self.register_buffer('pi_term', torch.log(t(2 * np.pi)))
        self.register_buffer('cov_dets', t(np.log(np.linalg.det(cov_t) + epsilon)))

    def get_mean(self):
        return self.weights @ self.means

    def _mahalanobis(self, pose):
        d = pose[:, None, :] - self.means[None]                                   # [B,M,D]
        return torch.einsum('bmi,mij,bmj->bm', d, self.precisions, d)

    def merged_log_likelihood(self, pose, betas):
        return (0.5 * self._mahalanobis(pose) - torch.log(self.nll_weights)).min(dim=1)[0]

    def log_likelihood(self, pose, betas, *args, **kwargs):
        # NOTE the reference's quirks, kept: the quadratic form enters without the factor 1/2 (:199-203) and the result is indexed
        # `[:, min_idx]`, i.e. it is [B,B] -- column j holds everybody's value for pose j's best component (:213-216)
        const = 0.5 * (torch.log(torch.det(self.covs) + self.epsilon) + self.random_var_dim * self.pi_term)
        ll = self._mahalanobis(pose) + const[None]
        best = torch.argmin(ll, dim=1)
        return -torch.log(self.nll_weights[:, best]) + ll[:, best]

    def forward(self, pose, betas):
        return self.merged_log_likelihood(pose, betas) if self.use_merged else self.log_likelihood(pose, betas)
