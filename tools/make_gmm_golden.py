"""Golden vectors for the MaxMixturePrior mirror: run the REFERENCE class (/root/reference/temp_prox/prior.py:100-231) on a synthetic
8-component, 69-d mixture (the licensed gmm_08.pkl is not available) and store inputs + outputs in tests/golden/reference_golden_gmm.npz.
Build container only (reads /root/reference)."""
import os
import pickle
import sys
import tempfile

import numpy as np
import torch

sys.path.insert(0, '/root/reference')
from temp_prox.prior import MaxMixturePrior as RefGMM        # noqa: E402

g = np.random.default_rng(7)
M, D = 8, 69
means = 0.3 * g.standard_normal((M, D))
A = 0.2 * g.standard_normal((M, D, D))
covars = A @ A.transpose(0, 2, 1) + 0.05 * np.eye(D)
w = g.random(M); w /= w.sum()
gmm = dict(means=means, covars=covars, weights=w)
d = tempfile.mkdtemp()
with open(os.path.join(d, 'gmm_08.pkl'), 'wb') as f:
    pickle.dump(gmm, f, protocol=2)
pose = torch.from_numpy((0.4 * g.standard_normal((12, D))).astype(np.float32))
betas = torch.zeros(12, 10)
out = {'means': means, 'covars': covars, 'weights': w, 'pose': pose.numpy()}
for merged in (True, False):
    ref = RefGMM(prior_folder=d, num_gaussians=8, use_merged=merged)
    out['nll_merged' if merged else 'nll_full'] = ref(pose, betas).detach().numpy()
    out['mean_pose'] = ref.get_mean().numpy()
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_golden_gmm.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, {k: v.shape for k, v in out.items()})
