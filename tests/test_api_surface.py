"""CPU: the Python mirror keeps the reference's call surface on the fitting path -- names, shapes, state_dict keys,
and the loud failure (no CPU fallback) when a product entry point is given host tensors."""
import os

import numpy as np
import pytest
import torch

from oracle import synth

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'lemo_b200', 'assets')


@pytest.fixture(scope='module')
def body():
    import lemo_b200.smplx as smplx
    return smplx.create(synth.make_smplx_model(0), model_type='smplx', gender='male', ext='npz', num_pca_comps=12,
                        create_global_orient=True, create_body_pose=True, create_betas=True, create_left_hand_pose=True,
                        create_right_hand_pose=True, create_expression=True, create_jaw_pose=True, create_leye_pose=True,
                        create_reye_pose=True, create_transl=True, batch_size=3)


def test_smplx_create_surface(body):
    """smplx.create(...) as called at opt_amass_temp.py:73-87 / temp_prox/main_slide.py:160-179."""
    names = [n for n, _ in body.named_parameters()]
    assert set(names) == {'betas', 'global_orient', 'transl', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose',
                          'expression', 'body_pose'}
    shapes = {n: tuple(p.shape) for n, p in body.named_parameters()}
    assert shapes['betas'] == (3, 10) and shapes['body_pose'] == (3, 63) and shapes['left_hand_pose'] == (3, 12) and shapes['transl'] == (3, 3)
    assert body.get_num_verts() == 10475
    assert body.faces.shape == (20908, 3) and body.faces_tensor.dtype == torch.int64 and tuple(body.faces_tensor.shape) == (20908, 3)
    assert body.joint_mapper is None
    body.joint_mapper = lambda j: j[:, :5]
    assert body.joint_mapper is not None
    body.joint_mapper = None


def test_reset_params_semantics(body):
    """smplx semantics (fit_temp_loadprox_slide.py:147): named parameters take the value, everything else is zero-filled."""
    with torch.no_grad():
        body.transl.fill_(3.0)
    body.reset_params(betas=np.full((3, 10), 0.5, np.float32))
    assert float(body.betas.detach().sum()) == 15.0 and float(body.transl.detach().abs().sum()) == 0.0
    body.reset_params()
    assert float(body.betas.detach().abs().sum()) == 0.0


def test_only_smplx_model_type():
    import lemo_b200.smplx as smplx
    with pytest.raises(ValueError):
        smplx.create(synth.make_smplx_model(0), model_type='smpl')


def test_prior_state_dicts_load_the_shipped_checkpoints():
    """models.AE / models.AE_sep keep the reference's state_dict keys: runs/59547/AE_last_model.pkl and runs/15217/Enc_last_model.pkl
    (exported to assets/*.npz as plain arrays) load with strict=True."""
    from lemo_b200.models.AE import AE
    from lemo_b200.models.AE_sep import Enc
    for net, f, n_tensors, n_params in ((AE(downsample=True, in_channel=4, kernel=3), 'ae_infill_59547.npz', 40, 4114219),
                                        (Enc(downsample=False, z_channel=64), 'enc_smooth_15217.npz', 20, 286560)):
        w = np.load(os.path.join(ASSETS, f))
        sd = {k: torch.from_numpy(w[k]) for k in w.files}
        assert len(sd) == n_tensors and sum(v.numel() for v in sd.values()) == n_params
        net.load_state_dict(sd, strict=True)
        assert all(torch.equal(net.state_dict()[k].cpu(), sd[k]) for k in sd)


def test_product_entry_points_refuse_host_tensors(body):
    """No CPU fallback anywhere on the product path: every mirror raises instead of computing on the host."""
    from lemo_b200.models.AE_sep import Enc
    from lemo_b200.models.AE import AE
    from lemo_b200.vposer import VPoserDecoder
    from lemo_b200.fit import TemporalFitter, PerFrameFitter
    with pytest.raises(RuntimeError, match='CUDA devices only'):
        body(return_verts=True)
    enc = Enc(downsample=False, z_channel=64)
    with pytest.raises(RuntimeError, match='CUDA devices only'):
        enc(torch.zeros(1, 1, 245, 134))
    with pytest.raises(RuntimeError, match='CUDA devices only'):
        AE(downsample=True, in_channel=4, kernel=3)(torch.zeros(1, 4, 210, 135))
    vp = VPoserDecoder(synth.make_vposer_weights(1))
    with pytest.raises(RuntimeError, match='CUDA devices only'):
        vp.decode(torch.zeros(2, 32), output_type='aa')
    for cls in (TemporalFitter, PerFrameFitter):
        with pytest.raises(RuntimeError):
            cls(body, vp, 1, 8, enc=enc, device='cpu') if cls is TemporalFitter else cls(body, vp, 1, 8, device='cpu')


def test_shard_assignment_is_round_robin():
    from lemo_b200 import shard
    ids = [list(shard.assign(64, 8, r)) for r in range(8)]
    assert ids[3][:3] == [3, 11, 19] and sorted(sum(ids, [])) == list(range(64))          # sequence s -> rank s mod N
