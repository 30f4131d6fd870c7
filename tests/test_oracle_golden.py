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


def test_infill_repr_and_reconstruction_match_reference():
    """oracle/ref_infill.py vs outputs of the reference's own get_local_markers_4chan / reconstruct_global_body
    (tests/golden/reference_golden_infill.npz, written by oracle/make_golden.py), plus the round trip body -> repr -> body."""
    import os
    from oracle import ref_infill as ri
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden_infill.npz')))
    for tag in ('a', 'b'):
        rep, rot0 = ri.get_local_markers_4chan(g['body_' + tag], g['contact_' + tag])
        assert np.abs(rep - g['repr_' + tag]).max() < 1e-12 and np.abs(rot0 - g['rot0_' + tag]).max() < 1e-12
        glob = ri.reconstruct_global_body(g['packed_' + tag], g['rot0_' + tag])
        assert np.abs(glob - g['global_' + tag]).max() < 1e-12
        # size-independent property: reconstruction inverts the representation up to the floor shift, the frame-0 pelvis (x, y)
        # origin and the dropped last frame
        body = g['body_' + tag].astype(np.float64)
        body[:, :, 2] -= np.float64(g['body_' + tag][:, :, 2].min())
        body[:, :, 0:2] -= body[0, 0, 0:2].copy()
        assert np.abs(glob - body[:-1]).max() < 1e-6


def test_infill_mask_rows_and_padding():
    from oracle import ref_infill as ri
    masked, loss_rows = ri.mask_rows(208)
    assert masked.shape == (66,) and masked.min() == 9 and masked.max() == 185 and len(set(masked.tolist())) == 66
    assert loss_rows.shape == (210 - 66 - 5,) and loss_rows[0] == 0 and loss_rows[-1] == 204
    x = np.arange(4 * 208 * 30, dtype=np.float32).reshape(4, 208, 30) + 1.0
    p = ri.prepare_input(x)
    assert p.shape == (4, 210, 46)
    assert np.all(p[0, masked + 1, :] == 0.0) and np.all(p[0, 205:209, :] == 0.0)
    assert np.array_equal(p[1, 1:-1, 8:-8], x[1]) and np.array_equal(p[1, 0, 8:-8], x[1, 1]) and np.array_equal(p[2, 5, 0:8], x[2, 4, 8:0:-1])


def test_tgm_restatement_against_scipy():
    """torchgeometry 0.1.2 is absent from the reference tree and this image (parity unpinned); its restated
    conversions are at least checked here against an INDEPENDENT implementation (scipy.spatial.transform.Rotation) on the inputs the
    fitting path produces: all four branches of rotation_matrix_to_quaternion (large rotations about each axis), small angles, and
    the canonical-sign convention of quaternion_to_angle_axis (angle in [0, pi] after the aa -> R -> aa round trip)."""
    from scipy.spatial.transform import Rotation
    g = np.random.default_rng(5)
    aa = np.concatenate([g.standard_normal((200, 3)) * 0.8,                         # generic
                         g.standard_normal((50, 3)) * 1e-3,                         # near identity
                         np.eye(3)[g.integers(0, 3, 60)] * 3.0 + 0.05 * g.standard_normal((60, 3)),    # ~172 deg about x / y / z
                         -np.eye(3)[g.integers(0, 3, 60)] * 2.5 + 0.05 * g.standard_normal((60, 3))]).astype(np.float64)
    R_sp = Rotation.from_rotvec(aa).as_matrix()
    R_tgm = rb.tgm_aa_to_rotmat(torch.from_numpy(aa)).numpy()                       # angle_axis_to_rotation_matrix (utils/utils.py:89)
    assert np.abs(R_tgm - R_sp).max() < 3e-6          # tgm normalises the axis with (theta + 1e-6): an O(1e-6 / theta) deviation by design
    assert np.abs(rb.rodrigues(torch.from_numpy(aa)).numpy() - R_sp).max() < 1e-6   # lbs.py batch_rodrigues adds 1e-8 to the vector
    aa_back = rb.rotmat_to_aa(torch.from_numpy(R_sp)).numpy()                       # rotation_matrix_to_angle_axis (utils/utils.py:80)
    ref = Rotation.from_matrix(R_sp).as_rotvec()                                    # scipy: angle in [0, pi]
    assert np.abs(aa_back - ref).max() < 1e-7, np.abs(aa_back - ref).max()
    # and as rotations (insensitive to any sign convention)
    assert np.abs(Rotation.from_rotvec(aa_back).as_matrix() - R_sp).max() < 1e-9


def test_chamfer_restatement_against_kdtree():
    """The `chamfer` CUDA extension (ChamferDistancePytorch@719b0f1) is absent from the reference tree (parity unpinned): its published
    definition -- exact nearest neighbour, squared L2, both directions, gradient 2 g (x1 - x2) to both clouds -- is checked against an
    independent implementation (scipy cKDTree) and against autograd of the explicit formula."""
    from scipy.spatial import cKDTree
    from oracle import ref_priors as rp
    g = np.random.default_rng(11)
    for B, n, m in ((1, 1, 1), (2, 37, 513), (3, 300, 29)):
        x1 = torch.from_numpy(g.standard_normal((B, n, 3))).requires_grad_(True)
        x2 = torch.from_numpy(g.standard_normal((B, m, 3))).requires_grad_(True)
        d1, d2, i1, i2 = rp.chamfer(x1, x2)
        assert i1.dtype == torch.int32 and i2.dtype == torch.int32 and tuple(d1.shape) == (B, n) and tuple(d2.shape) == (B, m)
        for b in range(B):
            dd, ii = cKDTree(x2[b].detach().numpy()).query(x1[b].detach().numpy())
            assert np.array_equal(ii, i1[b].numpy()) and np.allclose(dd ** 2, d1[b].detach().numpy(), atol=1e-12)
            dd, ii = cKDTree(x1[b].detach().numpy()).query(x2[b].detach().numpy())
            assert np.array_equal(ii, i2[b].numpy()) and np.allclose(dd ** 2, d2[b].detach().numpy(), atol=1e-12)
        g1, g2 = torch.from_numpy(g.standard_normal((B, n))), torch.from_numpy(g.standard_normal((B, m)))
        ((d1 * g1).sum() + (d2 * g2).sum()).backward()
        # chamfer.cu NmDistanceGradKernel: x1 += 2 g1 (x1 - nn), nn -= same; and symmetrically for the second direction
        e1 = torch.zeros_like(x1)
        e2 = torch.zeros_like(x2)
        for b in range(B):
            a = 2 * g1[b, :, None] * (x1[b] - x2[b][i1[b].long()]).detach()
            e1[b] += a
            e2[b].index_add_(0, i1[b].long(), -a)
            c = 2 * g2[b, :, None] * (x2[b] - x1[b][i2[b].long()]).detach()
            e2[b] += c
            e1[b].index_add_(0, i2[b].long(), -c)
        assert torch.allclose(x1.grad, e1, atol=1e-12) and torch.allclose(x2.grad, e2, atol=1e-12)
