"""Pin the oracle against the REAL reference and write the golden fixtures under tests/golden/.

Runs only in the build container (imports /root/reference; the GPU box has no such path).
Usage:  python -m oracle.make_golden

What it does
  1. imports the vendored reference modules (human_body_prior/body_model/lbs.py with the arithmetic-
     neutral `.contiguous()` shim of SURVEY.md 8c, models/AE.py, models/AE_sep.py),
  2. evaluates them on seeded synthetic inputs (numpy default_rng -> reproducible on any box),
  3. asserts the oracle restatement (oracle/ref_body.py, oracle/ref_priors.py) agrees, and
  4. saves the REFERENCE outputs as fixtures so tests/test_oracle_golden.py can re-check the oracle
     anywhere (CPU CI and the GPU box) without the reference tree.
"""
import os
import sys
import numpy as np
import torch

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '..', 'tests', 'golden')


def rng_state_dict(shapes, seed, scale=0.05):
    g = np.random.default_rng(seed)
    return {k: (scale * g.standard_normal(s)).astype(np.float32) for k, s in shapes.items()}


AE_SHAPES = {}
_c = [(4, 32), (32, 64), (64, 128), (128, 256), (256, 256)]
for _i, (a, b) in enumerate(_c, 1):
    AE_SHAPES['enc_blc%d.main.0.weight' % _i] = (b, a, 3, 3)
    AE_SHAPES['enc_blc%d.main.0.bias' % _i] = (b,)
    AE_SHAPES['enc_blc%d.main.2.weight' % _i] = (b, b, 3, 3)
    AE_SHAPES['enc_blc%d.main.2.bias' % _i] = (b,)
_d = [(256, 256), (256, 128), (128, 64), (64, 32), (32, 1)]
for _i, (a, b) in enumerate(_d, 1):
    AE_SHAPES['dec_blc%d.deconv1.weight' % _i] = (a, b, 3, 3)
    AE_SHAPES['dec_blc%d.deconv1.bias' % _i] = (b,)
    AE_SHAPES['dec_blc%d.deconv2.weight' % _i] = (b, b, 3, 3)
    AE_SHAPES['dec_blc%d.deconv2.bias' % _i] = (b,)


def main():
    sys.path.insert(0, REF)
    from human_body_prior.body_model import lbs as ref_lbs
    from models.AE import AE as RefAE
    from models.AE_sep import Enc as RefEnc
    from oracle import synth, ref_body as rb, ref_priors as rp

    _v2j = ref_lbs.vertices2joints
    ref_lbs.vertices2joints = lambda Jr, v: _v2j(Jr, v).contiguous()      # torch>=2 stride shim, arithmetic-neutral

    gold = {}
    # ---- LBS: small model (full outputs) and full-size model (marker rows + joints) ----
    for tag, nv, B in (('small', 640, 5), ('full', synth.V, 3)):
        m_np = synth.make_smplx_model(0, n_verts=nv)
        m = rb.model_to_torch(m_np)
        g = np.random.default_rng(11 + nv)
        betas = (g.standard_normal((B, 20))).astype(np.float32)
        pose = (0.4 * g.standard_normal((B, 165))).astype(np.float32)
        pose[0] = 0.0                                                          # rest pose row (1e-8 path)
        bt, pt = torch.from_numpy(betas), torch.from_numpy(pose)
        v_ref, j_ref = ref_lbs.lbs(bt, pt, m['v_template'].unsqueeze(0).expand(B, -1, -1), m['shapedirs'],
                                   m['posedirs'], m['J_regressor'], m['parents'], m['lbs_weights'])
        v_or, j_or, _ = rb.lbs(bt, pt, m)
        err = float((v_ref - v_or).abs().max() / v_ref.abs().max())
        print('lbs %s: oracle vs reference rel err %.2e' % (tag, err))
        assert err < 2e-6 and float((j_ref - j_or).abs().max()) < 2e-6
        gold['lbs_%s_betas' % tag], gold['lbs_%s_pose' % tag] = betas, pose
        gold['lbs_%s_joints' % tag] = j_ref.numpy()
        if tag == 'small':
            gold['lbs_small_verts'] = v_ref.numpy()
        else:
            rows = synth.load_tables()['markers81']
            gold['lbs_full_rows'] = rows
            gold['lbs_full_verts_rows'] = v_ref[:, rows].numpy()
            gold['lbs_full_verts_sum'] = v_ref.double().sum((1, 2)).numpy()
        # gradient golden (reference autograd) on the small model only
        if tag == 'small':
            pt2 = pt.clone().requires_grad_(True)
            bt2 = bt.clone().requires_grad_(True)
            v2, j2 = ref_lbs.lbs(bt2, pt2, m['v_template'].unsqueeze(0).expand(B, -1, -1), m['shapedirs'],
                                 m['posedirs'], m['J_regressor'], m['parents'], m['lbs_weights'])
            gw = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(v2.shape)).astype(np.float32))
            gj = torch.from_numpy(np.random.default_rng(6).standard_normal(tuple(j2.shape)).astype(np.float32))
            ((v2 * gw).sum() + (j2 * gj).sum()).backward()
            gold['lbs_small_gpose'], gold['lbs_small_gbetas'] = pt2.grad.numpy(), bt2.grad.numpy()

    # ---- Enc with the shipped real weights ----
    enc = RefEnc(downsample=False, z_channel=64)
    enc.load_state_dict(torch.load(os.path.join(REF, 'runs/15217/Enc_last_model.pkl'), map_location='cpu'))
    enc.eval()
    sd = {k: torch.from_numpy(v) for k, v in synth.load_enc_weights().items()}
    for tag, shp in (('small', (2, 1, 21, 30)), ('full', (1, 1, 245, 134))):
        x = torch.from_numpy(np.random.default_rng(21).standard_normal(shp).astype(np.float32) * 0.5)
        x.requires_grad_(True)
        z_ref = enc(x)[0]
        (z_ref[..., 1:] - z_ref[..., :-1]).pow(2).mean().backward()
        z_or = rp.enc_forward(x.detach(), sd)
        err = float((z_ref - z_or).abs().max() / z_ref.abs().max())
        print('Enc %s: oracle vs reference rel err %.2e' % (tag, err))
        assert err < 1e-5
        gold['enc_%s_x' % tag] = x.detach().numpy()
        if tag == 'small':
            gold['enc_small_z'] = z_ref.detach().numpy()
        else:
            gold['enc_full_z_sub'] = z_ref.detach().numpy()[:, ::8, ::7, ::9].copy()
        gold['enc_%s_gx' % tag] = x.grad.numpy()
        gold['enc_%s_loss' % tag] = np.float64((z_ref[..., 1:] - z_ref[..., :-1]).pow(2).mean().item())

    # ---- AE: real weights (here only) + rng weights (travels) ----
    ae = RefAE(downsample=True, in_channel=4, kernel=3)
    real = torch.load(os.path.join(REF, 'runs/59547/AE_last_model.pkl'), map_location='cpu')
    ae.load_state_dict(real)
    x = torch.from_numpy(np.random.default_rng(31).standard_normal((1, 4, 210, 135)).astype(np.float32) * 0.5)
    r_ref, z_ref = ae(x)
    r_or, z_or = rp.ae_forward(x, real)
    err = float((r_ref - r_or).abs().max() / r_ref.abs().max())
    print('AE real weights: oracle vs reference rel err %.2e' % err)
    assert err < 1e-5 and float((z_ref - z_or).abs().max()) < 1e-4
    rsd = {k: torch.from_numpy(v) for k, v in rng_state_dict(AE_SHAPES, 41).items()}
    ae.load_state_dict(rsd)
    for tag, shp in (('small', (1, 4, 37, 45)), ('full', (1, 4, 210, 135))):
        x = torch.from_numpy(np.random.default_rng(32).standard_normal(shp).astype(np.float32) * 0.5)
        ae.zero_grad()
        r_ref, z_ref = ae(x)
        r_or, z_or = rp.ae_forward(x, rsd)
        assert float((r_ref - r_or).abs().max() / r_ref.abs().max()) < 1e-5
        gold['ae_%s_x' % tag] = x.numpy()
        gold['ae_%s_rec' % tag] = r_ref.detach().numpy()
        gold['ae_%s_z' % tag] = z_ref.detach().numpy()
        if tag == 'small':
            (r_ref[:, 0] - x[:, 0]).abs().mean().backward()
            gold['ae_small_gw_first'] = ae.enc_blc1.main[0].weight.grad.numpy()
            gold['ae_small_gw_last'] = ae.dec_blc5.deconv2.weight.grad.numpy()
            gold['ae_small_gb_mid'] = ae.dec_blc2.deconv1.bias.grad.numpy()

    np.savez_compressed(os.path.join(OUT, 'reference_golden.npz'), **gold)
    print('wrote', os.path.join(OUT, 'reference_golden.npz'),
          '%.1f KB' % (os.path.getsize(os.path.join(OUT, 'reference_golden.npz')) / 1024))


from lemo_b200.synth import synth_marker_clip      # noqa: E402  (data generator, shared with bench.py)


def make_infill_golden():
    """Pin oracle/ref_infill.py against the reference's own get_local_markers_4chan / reconstruct_global_body.  utils/utils.py imports
    torchgeometry (absent here) at module level for unrelated functions: an empty stub module lets the import through."""
    import types
    sys.modules.setdefault('torchgeometry', types.ModuleType('torchgeometry'))
    sys.path.insert(0, REF)
    import utils.utils as U
    from oracle import ref_infill as ri
    gold = {}
    for tag, seed in (('a', 5), ('b', 6)):
        body, contact = synth_marker_clip(seed)
        ref_repr, ref_rot0 = U.get_local_markers_4chan(body.copy(), contact.copy())
        my_repr, my_rot0 = ri.get_local_markers_4chan(body, contact)
        e = np.abs(my_repr - ref_repr).max()
        print('get_local_markers_4chan[%s]: oracle vs reference max abs err %.3e, rot0 err %.3e' % (tag, e, np.abs(my_rot0 - ref_rot0).max()))
        assert e < 1e-12 and np.abs(my_rot0 - ref_rot0).max() < 1e-12
        # inverse: zero reference + local part + (vx, vy, r) trajectory, as opt_amass_temp.py:306-312 assembles it
        T = ref_repr.shape[1]
        local = ref_repr[0, :, 0:-4].reshape(T, 68, 3)
        traj = np.stack([ref_repr[1, :, 0], ref_repr[2, :, 0], ref_repr[3, :, 0]], -1)[:, None]
        packed = np.concatenate([np.zeros((T, 1, 3)), local, traj], axis=1)
        ref_glob = U.reconstruct_global_body(packed.copy(), ref_rot0)
        my_glob = ri.reconstruct_global_body(packed, ref_rot0)
        e = np.abs(my_glob - ref_glob).max()
        print('reconstruct_global_body[%s]: oracle vs reference max abs err %.3e' % (tag, e))
        assert e < 1e-12
        gold.update({'body_' + tag: body, 'contact_' + tag: contact, 'repr_' + tag: ref_repr, 'rot0_' + tag: ref_rot0,
                     'packed_' + tag: packed, 'global_' + tag: ref_glob})
    np.savez_compressed(os.path.join(OUT, 'reference_golden_infill.npz'), **gold)


if __name__ == '__main__':
    if 'infill' in sys.argv[1:]:
        make_infill_golden()
    else:
        main()
        make_infill_golden()
