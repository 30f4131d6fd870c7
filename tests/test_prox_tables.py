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


def test_fitting_monitor_adam_loop_semantics(capsys):
    """FittingMonitor.run_fitting (fitting_temp_slide.py:169-313, Adam branch) on a toy problem: `maxiters` steps, the first 15 % of
    the batch keeps its values when erase_first is set (:281-288), and a non-finite loss stops the run with the reference's message (:197-203)."""
    from lemo_b200.temp_prox.fitting_temp_slide import FittingMonitor
    target = torch.arange(20.).view(20, 1).expand(20, 3).clone()
    p = torch.zeros(20, 3, requires_grad=True)
    opt = torch.optim.Adam([p], lr=0.1)
    calls = []

    def closure():
        calls.append(1)
        return ((p - target) ** 2).sum()
    loss = FittingMonitor(maxiters=25, erase_first=True).run_fitting(opt, closure, [p])
    assert len(calls) == 25 and torch.isfinite(loss)
    assert float(p[:3].detach().abs().sum()) == 0.0 and float(p[3:].detach().abs().sum()) > 0.0       # int(20 * 0.15) = 3 frames frozen
    q = torch.zeros(4, requires_grad=True)
    n_calls = []
    FittingMonitor(maxiters=40, check_every=10).run_fitting(torch.optim.Adam([q], lr=0.1),
                                                            lambda: (n_calls.append(1), (q * float('nan')).sum())[1], [q])
    assert 'NaN loss value, stopping!' in capsys.readouterr().out and len(n_calls) == 10
    q = torch.zeros(4, requires_grad=True)                 # (the NaN run above poisoned the old one, as it does in the reference)
    FittingMonitor(maxiters=3).run_fitting(torch.optim.Adam([q], lr=0.1), lambda: (q * 0).sum() + float('inf'), [q])
    assert 'Infinite loss value, stopping!' in capsys.readouterr().out
    with pytest.raises(RuntimeError, match='capturable'):
        FittingMonitor(maxiters=10, use_cuda_graph=True).run_fitting(torch.optim.Adam([q], lr=0.1), lambda: q.sum(), [q])
