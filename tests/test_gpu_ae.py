"""GPU parity of the AE (motion-infilling prior): forward vs the REAL reference's outputs (golden), weight gradients and the
fused fine-tune step vs the oracle (autograd + torch.optim.Adam)."""
import numpy as np
import pytest
import torch

from oracle import ref_priors as rp
from oracle.make_golden import AE_SHAPES, rng_state_dict
from gpu_common import DEV, rel, rel_q

pytestmark = pytest.mark.gpu


def _ae():
    from lemo_b200.models.AE import AE
    ae = AE(downsample=True, in_channel=4, kernel=3)
    ae.load_state_dict(rng_state_dict(AE_SHAPES, 41))
    return ae.to(DEV)


def test_ae_forward_golden(golden):
    ae = _ae()
    assert list(ae.state_dict().keys()) == list(AE_SHAPES.keys())          # reference state_dict key order
    for tag in ('small', 'full'):
        rec, z = ae(torch.from_numpy(golden['ae_%s_x' % tag]).to(DEV))
        assert rec.shape == golden['ae_%s_rec' % tag].shape and z.shape == golden['ae_%s_z' % tag].shape
        assert rel(rec, golden['ae_%s_rec' % tag]) < 1e-4, rel(rec, golden['ae_%s_rec' % tag])
        assert rel(z, golden['ae_%s_z' % tag]) < 1e-4


def test_ae_weight_gradients(golden):
    ae = _ae()
    x = torch.from_numpy(golden['ae_small_x']).to(DEV)
    rec, _ = ae(x)
    (rec[:, 0] - x[:, 0]).abs().mean().backward()
    g = ae.flat.grad
    sd = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in rng_state_dict(AE_SHAPES, 41).items()}
    xr = torch.from_numpy(golden['ae_small_x']).double()
    r64, _ = rp.ae_forward(xr, sd)
    (r64[:, 0] - xr[:, 0]).abs().mean().backward()
    off = 0
    for k, shp in AE_SHAPES.items():
        n = int(np.prod(shp))
        mine = g[off:off + n].view(shp)
        off += n
        # quantile metric: LeakyReLU kinks / max-pool ties / sign(0) make isolated elements implementation-dependent
        assert rel_q(mine, sd[k].grad, 0.98) < 2e-4, (k, rel_q(mine, sd[k].grad, 0.98))
        assert rel(mine, sd[k].grad) < 5e-2, (k, rel(mine, sd[k].grad))
    # REAL reference gradients (golden) for three tensors
    views = {}
    off = 0
    for k, shp in AE_SHAPES.items():
        n = int(np.prod(shp)); views[k] = g[off:off + n].view(shp); off += n
    assert rel_q(views['enc_blc1.main.0.weight'], golden['ae_small_gw_first'], 0.98) < 5e-4
    assert rel_q(views['dec_blc5.deconv2.weight'], golden['ae_small_gw_last'], 0.98) < 5e-4
    assert rel_q(views['dec_blc2.deconv1.bias'], golden['ae_small_gb_mid'], 0.98) < 5e-4


def test_ae_finetune_matches_torch_adam(golden):
    """3 fused fine-tune steps (forward, masked L1, weight backward, Adam lr 3e-6) vs the reference op sequence on CPU."""
    ae = _ae()
    x = torch.from_numpy(golden['ae_small_x'])
    H = x.shape[2]
    rows = [r for r in range(H) if r % 3 != 1][:-5]                      # an arbitrary row subset, like upper_body_row[0:-5]
    losses = ae.finetune(x.to(DEV), rows, steps=3, lr=3e-6)
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in rng_state_dict(AE_SHAPES, 41).items()}
    opt = torch.optim.Adam(list(sd.values()), lr=3e-6)
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        rec, _ = rp.ae_forward(x, sd)
        loss = (rec[:, 0] - x[:, 0])[:, rows].abs().mean()
        loss.backward()
        opt.step()
        ref_losses.append(float(loss))
    assert np.allclose(losses.cpu().numpy(), ref_losses, rtol=2e-4), (losses.cpu().numpy(), ref_losses)
    new = ae.state_dict()
    w0 = rng_state_dict(AE_SHAPES, 41)
    for k in ('enc_blc1.main.0.weight', 'enc_blc5.main.2.weight', 'dec_blc3.deconv1.weight', 'dec_blc5.deconv2.bias'):
        step_ref = sd[k].detach().numpy() - w0[k]
        step_mine = new[k].cpu().numpy() - w0[k]
        # Adam steps are ~lr*sign(g): compare the update directions on the elements whose gradient is not ~0
        agree = np.mean(np.sign(step_ref) == np.sign(step_mine))
        assert agree > 0.97, (k, agree)
        assert np.abs(step_mine).max() < 4 * 3e-6 * 3
