"""The oracle (oracle/) re-checked against outputs of the REAL reference (tests/golden/reference_golden.npz,
written by oracle/make_golden.py in the build container).  CPU only."""
import numpy as np
import torch

from oracle import synth, ref_body as rb, ref_priors as rp
from oracle.make_golden import AE_SHAPES, rng_state_dict
from conftest import relerr


def test_lbs_small_forward_and_grad(golden):
    m = rb.model_to_torch(synth.make_smplx_model(0, n_verts=640))
    bt = torch.from_numpy(golden['lbs_small_betas']).requires_grad_(True)
    pt = torch.from_numpy(golden['lbs_small_pose']).requires_grad_(True)
    v, j, _ = rb.lbs(bt, pt, m)
    assert relerr(v, golden['lbs_small_verts']) < 2e-6
    assert relerr(j, golden['lbs_small_joints']) < 2e-6
    gw = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(v.shape)).astype(np.float32))
    gj = torch.from_numpy(np.random.default_rng(6).standard_normal(tuple(j.shape)).astype(np.float32))
    ((v * gw).sum() + (j * gj).sum()).backward()
    assert relerr(pt.grad, golden['lbs_small_gpose']) < 1e-4
    assert relerr(bt.grad, golden['lbs_small_gbetas']) < 1e-4


def test_lbs_full_rows(golden):
    m = rb.model_to_torch(synth.make_smplx_model(0))
    v, j, _ = rb.lbs(torch.from_numpy(golden['lbs_full_betas']), torch.from_numpy(golden['lbs_full_pose']), m)
    rows = golden['lbs_full_rows']
    assert np.array_equal(rows, synth.load_tables()['markers81'])          # index table is bit-exact data
    assert relerr(v[:, rows], golden['lbs_full_verts_rows']) < 2e-6
    assert relerr(j, golden['lbs_full_joints']) < 2e-6
    assert np.allclose(v.double().sum((1, 2)).numpy(), golden['lbs_full_verts_sum'], rtol=1e-5, atol=1e-2)


def test_enc_real_weights(golden):
    sd = {k: torch.from_numpy(v) for k, v in synth.load_enc_weights().items()}
    x = torch.from_numpy(golden['enc_small_x']).requires_grad_(True)
    z = rp.enc_forward(x, sd)
    assert relerr(z, golden['enc_small_z']) < 1e-5
    loss = (z[..., 1:] - z[..., :-1]).pow(2).mean()
    loss.backward()
    assert abs(float(loss) - float(golden['enc_small_loss'])) < 1e-5 * abs(float(golden['enc_small_loss']))
    assert relerr(x.grad, golden['enc_small_gx']) < 1e-4
    zf = rp.enc_forward(torch.from_numpy(golden['enc_full_x']), sd)
    assert relerr(zf[:, ::8, ::7, ::9], golden['enc_full_z_sub']) < 1e-5


def test_ae_rng_weights(golden):
    sd = {k: torch.from_numpy(v) for k, v in rng_state_dict(AE_SHAPES, 41).items()}
    for tag in ('small', 'full'):
        rec, z = rp.ae_forward(torch.from_numpy(golden['ae_%s_x' % tag]), sd)
        assert relerr(rec, golden['ae_%s_rec' % tag]) < 1e-5
        assert relerr(z, golden['ae_%s_z' % tag]) < 1e-5


def test_tables_and_seed_clips():
    t = synth.load_tables()
    assert t['markers67'].shape == (67,) and t['markers81'].shape == (81,)
    assert np.array_equal(t['markers81'][:67], t['markers67'])
    assert [t[k].shape[0] for k in ('left_heel', 'left_toe', 'right_heel', 'right_toe')] == [32, 57, 34, 49]
    assert int(t['markers67'][[16, 30, 47, 60]].tolist() == [8846, 5787, 8634, 8481])      # LHEE/LTOE/RHEE/RTOE
    clean, init, contact = synth.make_sequence(3)
    assert clean.shape == (119, 72) and contact.shape == (119, 4) and set(np.unique(contact)) <= {0.0, 1.0}


def test_rotation_round_trips():
    g = torch.Generator().manual_seed(0)
    aa = torch.randn(64, 3, generator=g)
    R = rb.rodrigues(aa)
    assert torch.allclose(rb.rodrigues(rb.rotmat_to_aa(R)), R, atol=2e-6)           # R -> aa (tgm) -> Rodrigues == R
    x6 = rb.convert_to_6D_all(aa)
    assert torch.allclose(rb.gram_schmidt_6d(x6), rb.tgm_aa_to_rotmat(aa), atol=2e-6)
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(64, 3, 3), atol=1e-5)
