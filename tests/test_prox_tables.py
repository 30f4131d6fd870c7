"""CPU: the PROX-path glue mirrors (lemo_b200.temp_prox.prior / misc_utils) against known-answer vectors produced by the REFERENCE's own
modules (tests/golden/reference_golden_prox.npz, written by tools/export_assets.py prox from /root/reference/temp_prox/{prior,misc_utils}.py),
and the invariants of the exported index tables that the reference's loss relies on."""
import os

import numpy as np
import pytest
import torch

from lemo_b200.temp_prox import misc_utils as mu
from lemo_b200.temp_prox import prior as pr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden_prox.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def test_priors_match_reference_outputs(gold):
    pose, pose_g = torch.from_numpy(gold['pose']), torch.from_numpy(gold['pose_g'])
    assert np.array_equal(pr.SMPLifyAnglePrior()(pose).numpy(), gold['angle'])                       # same torch ops: bit-exact
    assert np.array_equal(pr.SMPLifyAnglePrior()(pose_g, with_global_pose=True).numpy(), gold['angle_g'])
    assert np.array_equal(pr.L2Prior()(pose).numpy(), gold['l2'])
    assert np.array_equal(pr.create_prior('angle')(pose).numpy(), gold['angle'])
    assert np.array_equal(pr.create_prior('l2')(pose).numpy(), gold['l2'])
    assert pr.create_prior(None)(pose) == 0.0 and pr.create_prior('none')(pose) == 0.0
    with pytest.raises(ValueError):
        pr.create_prior('laplace')


def test_gmm_prior_matches_reference_outputs():
    """MaxMixturePrior (prior.py:100-231) on a synthetic 8-component mixture: merged and per-component likelihoods equal the
    outputs of the reference class (tests/golden/reference_golden_gmm.npz, tools/make_gmm_golden.py), and create_prior('gmm') builds it."""
    g = np.load(os.path.join(os.path.dirname(GOLD), 'reference_golden_gmm.npz'))
    gmm = dict(means=g['means'], covars=g['covars'], weights=g['weights'])
    pose = torch.from_numpy(g['pose'])
    for merged, key in ((True, 'nll_merged'), (False, 'nll_full')):
        m = pr.create_prior('gmm', gmm=gmm, use_merged=merged)
        assert isinstance(m, pr.MaxMixturePrior) and m.num_gaussians == 8 and m.random_var_dim == 69
        out = m(pose, torch.zeros(12, 10)).numpy()
        assert np.allclose(out, g[key], rtol=1e-5, atol=1e-4), np.abs(out - g[key]).max()
    assert np.allclose(m.get_mean().numpy(), g['mean_pose'], atol=1e-6)


def test_oracle_angle_term_matches_reference(gold):
    """The restatement inside oracle/ref_prox.py (and lemo_b200's SMPLifyLoss) uses `full_pose[:, 3:66][:, idx - 3] * sgn`."""
    pose = torch.from_numpy(gold['pose'])
    idx = torch.tensor([55, 58, 12, 15]) - 3
    sgn = torch.tensor([1., -1., -1., -1.])
    assert np.array_equal(torch.exp(pose[:, idx] * sgn).numpy(), gold['angle'])


def test_robustifiers_match_reference_outputs(gold):
    res = torch.from_numpy(gold['res'])
    assert np.array_equal(mu.GMoF(rho=100)(res).numpy(), gold['gmof100'])
    assert np.array_equal(mu.GMoF_unscaled(rho=0.5)(res).numpy(), gold['gmof_unscaled'])


def test_joint_mapper_and_openpose_maps(gold):
    m = mu.smpl_to_openpose('smplx', use_hands=True, use_face=True, use_face_contour=False, openpose_format='coco25')
    assert m.dtype == np.int64 and m.shape == (118,) and m.min() >= 0 and m.max() < 127          # temp_prox/main_slide.py:160-179
    j = torch.arange(127.).view(1, 127, 1).expand(2, 127, 3)
    out = mu.JointMapper(m)(j)
    assert np.array_equal(out.numpy(), gold['mapped'])                                             # integer gather: bit-exact
    assert mu.JointMapper()(j) is j
    sizes = {(fmt, h, f, c): len(mu.smpl_to_openpose('smplx', bool(h), bool(f), bool(c), fmt))
             for fmt in ('coco25', 'coco19') for h in (0, 1) for f in (0, 1) for c in (0, 1)}
    assert sizes[('coco25', 1, 1, 0)] == 118 and sizes[('coco25', 1, 1, 1)] == 135 and sizes[('coco25', 0, 0, 0)] == 25
    assert sizes[('coco19', 1, 1, 0)] == 112 and sizes[('coco19', 0, 0, 0)] == 19
    with pytest.raises(ValueError):
        mu.smpl_to_openpose('smplx', openpose_format='coco17')
    with pytest.raises(ValueError):
        mu.smpl_to_openpose('smpl')


def test_prox_vertex_tables():
    t = mu.prox_tables()
    fric, con = t['friction_ids'], t['contact_ids']
    assert fric.shape == (307,) and con.shape == (1121,)                     # fit_temp_loadprox_slide.py:349-362
    assert fric.min() >= 0 and max(fric.max(), con.max()) < 10475
    assert len(set(fric.tolist())) == 307                                    # L_Leg, R_Leg, gluteus are disjoint vertex sets
    assert set(fric.tolist()) <= set(con.tolist())                           # contact parts include the three friction parts


def test_create_loss_dispatch():
    from lemo_b200.temp_prox import fitting_temp_slide as fs
    with pytest.raises(ValueError):
        fs.create_loss('camera_init')
    assert fs.create_loss.__doc__


def test_fitting_monitor_closure_protocol(capsys):
    """FittingMonitor with the reference's signatures (fitting_temp_slide.py:137-217) on a toy closure (host tensors, so the eager
    `optimizer.step(closure)` protocol is what runs): `maxiters` steps inside the context manager, the float of the last finite loss is
    returned, a non-finite loss stops the run with the reference's message (:197-203)."""
    from lemo_b200.temp_prox.fitting_temp_slide import FittingMonitor
    target = torch.arange(20.).view(20, 1).expand(20, 3).clone()
    p = torch.zeros(20, 3, requires_grad=True)
    opt = torch.optim.Adam([p], lr=0.1)
    calls = []

    def closure():
        opt.zero_grad()
        loss = ((p - target) ** 2).sum()
        loss.backward()
        calls.append(float(loss))
        return loss
    with FittingMonitor(summary_steps=1, maxiters=25, ftol=2e-9, gtol=1e-5, model_type='smplx', unknown_kwarg=1) as monitor:
        final = monitor.run_fitting(opt, closure, [p], None, use_vposer=True, pose_embedding=None, vposer=None)
    assert 'total steps:' in capsys.readouterr().out
    assert len(calls) == 25 and isinstance(final, float) and final == calls[-1] and calls[-1] < calls[0]
    assert monitor.last_path.startswith('eager')
    q = torch.zeros(4, requires_grad=True)
    n_calls = []
    r = FittingMonitor(maxiters=40).run_fitting(torch.optim.Adam([q], lr=0.1), lambda: (n_calls.append(1), (q * float('nan')).sum())[1], [q], None)
    assert 'NaN loss value, stopping!' in capsys.readouterr().out and len(n_calls) == 1 and r is None
    q = torch.zeros(4, requires_grad=True)
    FittingMonitor(maxiters=3).run_fitting(torch.optim.Adam([q], lr=0.1), lambda: (q * 0).sum() + float('inf'), [q], None)
    assert 'Infinite loss value, stopping!' in capsys.readouterr().out


def test_smplify_loss_reference_signature():
    """SMPLifyLoss(**the kwargs fit_temp_loadprox_slide.py:431-485 passes) constructs, exposes the reference's weight buffers, accepts
    reset_loss_weights (:548-562) with floats and tensors, and reports which configurations the fused driver covers."""
    import inspect
    from lemo_b200.temp_prox.fitting_temp_slide import SMPLifyLoss, FittingMonitor, create_loss
    from lemo_b200.temp_prox.prior import create_prior
    ref_init = ['search_tree', 'pen_distance', 'tri_filtering_module', 'body_pose_prior', 'shape_prior', 'expr_prior', 'angle_prior', 'jaw_prior',
                'use_joints_conf', 'use_face', 'use_hands', 'left_hand_prior', 'right_hand_prior', 'interpenetration', 'dtype', 'data_weight',
                'body_pose_weight', 'shape_weight', 'bending_prior_weight', 'hand_prior_weight', 'expr_prior_weight', 'jaw_prior_weight',
                'coll_loss_weight', 's2m', 'm2s', 'rho_s2m', 'rho_m2s', 's2m_weight', 'm2s_weight', 'head_mask', 'body_mask', 'sdf_penetration',
                'voxel_size', 'grid_min', 'grid_max', 'sdf', 'sdf_normals', 'sdf_penetration_weight', 'R', 't', 'contact', 'contact_loss_weight',
                'contact_verts_ids', 'smooth_acc', 'smooth_acc_weight', 'smooth_vel', 'smooth_vel_weight', 'use_motion_smooth_prior',
                'motion_prior_smooth_weight', 'motion_smooth_model', 'use_friction', 'friction_normal_weight', 'friction_tangent_weight',
                'contact_fric_verts_ids', 'use_motion_infill_prior', 'motion_infill_rec_weight', 'motion_infill_contact_weight',
                'motion_infill_model', 'infill_pretrain_weights', 'device']
    assert list(inspect.signature(SMPLifyLoss.__init__).parameters)[1:-1] == ref_init
    fwd = list(inspect.signature(SMPLifyLoss.forward).parameters)[1:]
    assert fwd == ['body_model', 'body_model_output', 'smplx_joints', 'camera', 'gt_joints', 'joints_conf', 'marker_mask', 'body_model_faces',
                   'joint_weights', 'use_vposer', 'pose_embedding', 'scan_tensor', 'scan_point_num', 'scene_v', 'opt_step', 'kwargs']
    assert list(inspect.signature(FittingMonitor.run_fitting).parameters)[1:] == ['optimizer', 'closure', 'params', 'body_model', 'use_vposer',
                                                                                  'pose_embedding', 'vposer', 'kwargs']
    cl = list(inspect.signature(FittingMonitor.create_fitting_closure).parameters)[1:]
    assert cl == ['optimizer', 'body_model', 'camera', 'gt_joints', 'loss', 'joints_conf', 'marker_mask', 'joint_weights', 'return_verts',
                  'return_full_pose', 'use_vposer', 'vposer', 'pose_embedding', 'scan_tensor', 'scan_point_num', 'scene_v', 'create_graph',
                  'writer', 'first_batch_flag', 'kwargs']
    loss = create_loss(loss_type='smplify', rho=100, vposer=None, pose_embedding=None, body_pose_prior=create_prior('l2'),
                       shape_prior=create_prior('l2'), angle_prior=create_prior('angle'), expr_prior=create_prior('l2'),
                       left_hand_prior=create_prior('l2'), right_hand_prior=create_prior('l2'), jaw_prior=create_prior('l2'),
                       interpenetration=False, sdf_penetration=True, sdf=torch.zeros(2, 1, 4, 4, 4), grid_min=torch.zeros(2, 1, 3),
                       grid_max=torch.ones(2, 1, 3), R=torch.eye(3), t=torch.zeros(1, 3), contact=True, contact_verts_ids=np.arange(5),
                       use_friction=True, contact_fric_verts_ids=np.arange(3), use_motion_smooth_prior=True, motion_smooth_model=None,
                       device='cpu')
    assert tuple(loss.sdf.shape) == (4, 4, 4) and tuple(loss.grid_min.shape) == (3,)          # one volume kept, not B replicas
    loss.reset_loss_weights({'data_weight': 2.0, 'contact_loss_weight': torch.tensor(0.5), 'hand_weight': 1.0, 'not_an_attribute': 3})
    assert float(loss.data_weight) == 2.0 and float(loss.contact_loss_weight) == 0.5
    assert loss.weight_dict()['data_weight'] == 2.0 and loss.weight_dict()['friction_normal_weight'] == 0.0
    assert loss.fusable(use_vposer=True)[0]
    assert not loss.fusable(use_vposer=False)[0]
    loss2 = create_loss(smooth_vel=True, smooth_vel_weight=1.0, angle_prior=create_prior('angle'), interpenetration=False, R=torch.eye(3), t=torch.zeros(1, 3))
    assert not loss2.fusable(use_vposer=True)[0]
    with pytest.raises(ValueError):
        create_loss(loss_type='camera_init')




def test_create_optimizer_kinds():
    """optimizers/optim_factory.py:26-65: (optimizer, False) for every type the reference accepts, ValueError otherwise."""
    import torch
    from lemo_b200.temp_prox.optimizers import create_optimizer
    p = [torch.nn.Parameter(torch.zeros(3))]
    kinds = {'adam': torch.optim.Adam, 'lbfgs': torch.optim.LBFGS, 'lbfgsls': torch.optim.LBFGS, 'rmsprop': torch.optim.RMSprop,
             'sgd': torch.optim.SGD}
    for k, cls in kinds.items():
        opt, flag = create_optimizer(p, optim_type=k, lr=0.005, maxiters=30)
        assert isinstance(opt, cls) and flag is False
    assert create_optimizer(p, optim_type='lbfgsls', lr=1.0, maxiters=30)[0].param_groups[0]['line_search_fn'] == 'strong_wolfe'
    a = create_optimizer(p, optim_type='adam', lr=0.005, beta1=0.9, beta2=0.999)[0].param_groups[0]
    assert a['lr'] == 0.005 and tuple(a['betas']) == (0.9, 0.999)
    with pytest.raises(ValueError):
        create_optimizer(p, optim_type='adagrad')
