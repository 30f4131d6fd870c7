"""PROX stage-2 window (BASELINE config 4) through the REFERENCE'S OWN CALL SURFACE -- create_loss(**kwargs of
fit_temp_loadprox_slide.py:431-485), FittingMonitor as a context manager, create_fitting_closure, run_fitting(optimizer, closure,
params, body_model, pose_embedding=, vposer=, use_vposer=) -- against the oracle restatement (oracle/ref_prox.py):

  * closure level: every loss term and every parameter gradient, for the fused device driver (lemo_fit_prox_eval) AND the eager
    closure (autograd over the lemo operators), at a reduced size and at the full config-4 size (B=100, 256^3 SDF, 100 000 scene
    points);
  * loop level: 20 closure steps with the first-15 % freeze, fused vs eager vs the oracle's torch.optim.Adam loop.
"""
import numpy as np
import pytest
import torch

from oracle import synth, ref_body as rb, ref_prox
from gpu_common import DEV, smplx_module, vposer_module, enc_module, oracle_ctx, rel

pytestmark = pytest.mark.gpu
PKEYS = ref_prox.PKEYS
BODY_KEYS = ['transl', 'global_orient', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'expression', 'betas']
# oracle term -> loss_dict key of the product
TERMS = {'joint': 'joint_loss', 'pprior': 'pprior_loss', 'shape': 'shape_loss', 'angle': 'angle_prior_loss', 'hand': 'hand_prior_loss',
         'expr': 'expression_loss', 'jaw': 'jaw_prior_loss', 'sdf': 'sdf_penetration_loss', 'fric_t': 'loss_fric_tangent',
         'fric_n': 'loss_fric_normal', 'contact': 'contact_loss', 'smooth': 'motion_prior_smooth_loss'}


def _reference_call_surface(B, P, cfg, maxiters, first_batch_flag, lr=0.005):
    """What fit_temp_loadprox_slide.fit_single_frame does (:431-559), with the drop-in modules.  Returns everything the test needs."""
    import lemo_b200.smplx as smplx
    from lemo_b200.temp_prox import fitting_temp_slide as fitting
    from lemo_b200.temp_prox.camera import PerspectiveCamera
    from lemo_b200.temp_prox.prior import create_prior
    from lemo_b200.temp_prox.misc_utils import JointMapper
    w = cfg['w']
    body_model = smplx.create(synth.make_smplx_model(0), model_type='smplx', gender='male', ext='npz', num_pca_comps=12, batch_size=B,
                              joint_mapper=JointMapper(cfg['joint_map']), create_body_pose=False).to(DEV)
    vposer = vposer_module()
    Rc, tc, fx, fy, cc = cfg['camera']
    camera = PerspectiveCamera(rotation=Rc[None].repeat(B, 1, 1), translation=tc[None].repeat(B, 1), focal_length_x=fx, focal_length_y=fy,
                               batch_size=B, center=cc[None].repeat(B, 1)).to(DEV)
    D = cfg['sdf'].shape[-1]
    Rw, tw = cfg['cam2world']
    loss = fitting.create_loss(loss_type='smplify', joint_weights=cfg['joint_weights'].to(DEV), rho=100, use_joints_conf=True, use_face=True,
                               use_hands=True, vposer=vposer, pose_embedding=None, body_pose_prior=create_prior('l2'),
                               shape_prior=create_prior('l2'), angle_prior=create_prior('angle'), expr_prior=create_prior('l2'),
                               left_hand_prior=create_prior('l2'), right_hand_prior=create_prior('l2'), jaw_prior=create_prior('l2'),
                               interpenetration=False, s2m=False, m2s=False, sdf_penetration=True,
                               grid_min=cfg['grid_min'].to(DEV).repeat(B, 1).unsqueeze(1), grid_max=cfg['grid_max'].to(DEV).repeat(B, 1).unsqueeze(1),
                               sdf=cfg['sdf'].to(DEV).view(1, 1, D, D, D),          # (the reference's caller repeats it B times: a view is enough)
                               R=Rw.to(DEV), t=tw.to(DEV).view(1, 3), contact=True, contact_verts_ids=cfg['contact_ids'].numpy(),
                               dtype=torch.float32, use_motion_smooth_prior=True, motion_smooth_model=enc_module(), use_friction=True,
                               contact_fric_verts_ids=cfg['fric_ids'].numpy(), use_motion_infill_prior=False, device=DEV).to(DEV)
    body_model.reset_params(**{k: P[k] for k in BODY_KEYS})
    body_model.betas.requires_grad = False
    pose_embedding = torch.from_numpy(P['pose_embedding']).float().to(DEV).requires_grad_(True)
    final_params = [p for p in body_model.parameters() if p.requires_grad] + [pose_embedding]
    optimizer = torch.optim.Adam(final_params, lr=lr, betas=(0.9, 0.999))
    loss.reset_loss_weights(dict(data_weight=w['data'], body_pose_weight=w['body_pose'], shape_weight=w.get('shape', 0.0),
                                 bending_prior_weight=3.17 * w['body_pose'], hand_prior_weight=w['hand_prior'], expr_prior_weight=w['expr'],
                                 jaw_prior_weight=w['jaw'], sdf_penetration_weight=w['sdf'], contact_loss_weight=w['contact'],
                                 motion_prior_smooth_weight=w['smooth'], friction_normal_weight=w['fric_n'], friction_tangent_weight=w['fric_t']))
    monitor = fitting.FittingMonitor(maxiters=maxiters, model_type='smplx')
    closure = monitor.create_fitting_closure(optimizer, body_model, camera=camera, gt_joints=cfg['gt_joints'].to(DEV),
                                             joints_conf=cfg['joints_conf'].to(DEV), marker_mask=None, joint_weights=cfg['joint_weights'].to(DEV),
                                             loss=loss, create_graph=False, use_vposer=True, vposer=vposer, pose_embedding=pose_embedding,
                                             scan_tensor=None, scan_point_num=None, scene_v=cfg['scene_v'].to(DEV).unsqueeze(0),
                                             return_verts=True, return_full_pose=True, writer=None, first_batch_flag=first_batch_flag)
    return dict(body_model=body_model, vposer=vposer, loss=loss, monitor=monitor, closure=closure, optimizer=optimizer,
                final_params=final_params, pose_embedding=pose_embedding)


def _params_of(s):
    d = {k: getattr(s['body_model'], k).detach().cpu().numpy().copy() for k in BODY_KEYS}
    d['pose_embedding'] = s['pose_embedding'].detach().cpu().numpy().copy()
    return d


@pytest.mark.parametrize('B,D,m_scene,tol_t,tol_g', [(24, 32, 3000, 2e-3, 5e-3), (100, 256, 100000, 2e-3, 5e-3)])
def test_closure_terms_and_gradients(B, D, m_scene, tol_t, tol_g):
    P, cfg = synth.make_prox_problem(B, D=D, m_scene=m_scene, seed=1 if B == 24 else 3)
    cfg['w']['shape'] = 0.5
    c32 = oracle_ctx(torch.float32)
    Pt = {k: torch.from_numpy(v).requires_grad_(k in PKEYS) for k, v in P.items()}
    tot_ref, T_ref = ref_prox.s2_loss(Pt, c32, cfg)
    tot_ref.backward()
    assert float(T_ref['sdf']) > 0 and float(T_ref['contact']) > 0 and float(T_ref['fric_t']) > 0 and float(T_ref['fric_n']) > 0
    erase_n = int(B * 0.15)
    # ---- eager closure through the reference surface (first_batch_flag=False: the closure erases the first 15 % of every gradient)
    s = _reference_call_surface(B, P, cfg, maxiters=1, first_batch_flag=False)
    total = s['closure'](backward=True)
    ld = s['closure'].last_loss_dict
    for k, name in TERMS.items():
        a, b = float(ld[name]), float(T_ref[k])
        assert abs(a - b) <= tol_t * abs(b) + 1e-7, ('eager', k, a, b)
    assert abs(float(total) - float(tot_ref)) <= tol_t * abs(float(tot_ref))
    g_eager = {k: (s['pose_embedding'] if k == 'pose_embedding' else getattr(s['body_model'], k)).grad for k in PKEYS}
    # ---- fused driver: one closure evaluation
    fit = s['monitor']._fitter(s['closure'].lemo_spec, s['body_model'], s['pose_embedding'], s['vposer'])
    fit.set_weights(s['loss'].weight_dict(), erase_n, True)
    Pd = {k: getattr(s['body_model'], k) for k in BODY_KEYS}
    Pd['pose_embedding'] = s['pose_embedding']
    fit.set_window(Pd, cfg['gt_joints'], cfg['joints_conf'], cfg['joint_weights'])
    fit.eval()
    lf, gf = fit.losses(), fit.grads()
    for k, name in TERMS.items():
        a, b = float(lf[name]), float(T_ref[k])
        assert abs(a - b) <= tol_t * abs(b) + 1e-7, ('fused', k, a, b)
    assert abs(float(lf['total_loss']) - float(tot_ref)) <= tol_t * abs(float(tot_ref))
    for k in PKEYS:
        ref = Pt[k].grad.clone()
        ref[:erase_n] = 0
        assert float(gf[k][:erase_n].abs().max()) == 0.0 and float(g_eager[k][:erase_n].abs().max()) == 0.0
        assert rel(gf[k], ref) < tol_g, ('fused', k, rel(gf[k], ref))
        assert rel(g_eager[k], ref) < tol_g, ('eager', k, rel(g_eager[k], ref))


def _record(key, val):
    import json, os
    try:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
        os.makedirs(out, exist_ok=True)
        p = os.path.join(out, 'prox_parity.json')
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[key] = val
        json.dump(d, open(p, 'w'), indent=1, sort_keys=True)
    except Exception:
        pass


def test_run_fitting_20_steps_with_freeze_fused_eager_oracle(monkeypatch):
    """20 closure steps incl. the first-15 % freeze: run_fitting on the fused driver and on the eager closure vs the oracle's
    torch.optim.Adam loop.  The loss curve is compared step by step (the fused driver is stepped with resume=True, which must equal one
    20-step call): tight while the trajectories are still together, looser at the end (L1 keypoints, SDF clamp, friction thresholds and
    LeakyReLU kinks make the loop sensitive to rounding, as in the AMASS loops -- tests/test_gpu_loops_baseline.py)."""
    B, n_it = 24, 20
    P, cfg = synth.make_prox_problem(B, D=32, m_scene=3000, seed=2)
    cfg['w']['shape'] = 0.5
    c32 = oracle_ctx(torch.float32)
    tr = []
    P_ref, last_ref = ref_prox.fit_window(P, c32, cfg, n_it, lr=0.005, first_batch_flag=False, trace=tr)
    curve_ref = [sum(t.values()) for t in tr]
    res = {}
    for mode in ('fused', 'eager'):
        monkeypatch.setenv('LEMO_PROX_FUSED', '1' if mode == 'fused' else '0')
        s = _reference_call_surface(B, P, cfg, maxiters=n_it, first_batch_flag=False)
        with s['monitor'] as monitor:
            final = monitor.run_fitting(s['optimizer'], s['closure'], s['final_params'], s['body_model'], pose_embedding=s['pose_embedding'],
                                        vposer=s['vposer'], use_vposer=True)
        assert monitor.last_path.startswith(mode), monitor.last_path
        assert monitor.steps == n_it
        res[mode] = (_params_of(s), final)
    # the fused driver stepped one closure at a time (resume) = the same 20 steps; collects the loss curve
    monkeypatch.setenv('LEMO_PROX_FUSED', '1')
    s = _reference_call_surface(B, P, cfg, maxiters=n_it, first_batch_flag=False)
    fit = s['monitor']._fitter(s['closure'].lemo_spec, s['body_model'], s['pose_embedding'], s['vposer'])
    fit.set_weights(s['loss'].weight_dict(), int(B * 0.15), True)
    Pd = {k: getattr(s['body_model'], k) for k in BODY_KEYS}
    Pd['pose_embedding'] = s['pose_embedding']
    fit.set_window(Pd, cfg['gt_joints'], cfg['joints_conf'], cfg['joint_weights'])
    curve = []
    for it in range(n_it):
        fit.run(1, 0.005, resume=it > 0)
        curve.append(float(fit.losses()['total_loss']))
    stepped = {k: v.cpu().numpy() for k, v in fit.params().items()}
    dev = [abs(a - b) / abs(b) for a, b in zip(curve, curve_ref)]
    _record('window_B24_20steps', dict(loss_curve_oracle=curve_ref, loss_curve_fused=curve, rel_dev=dev, final_fused=res['fused'][1],
                                       final_eager=res['eager'][1], final_oracle=last_ref,
                                       params_max_abs_dev={m: {k: float(np.abs(res[m][0][k] - P_ref[k]).max()) for k in PKEYS} for m in res}))
    print(dev)
    assert max(dev[:3]) < 2e-3, dev                       # together at the start (closure-level parity)
    assert max(dev) < 2e-2, dev                           # and still at the same loss level after 20 steps
    assert curve_ref[-1] < curve_ref[0] and curve[-1] < curve[0]
    erase_n = int(B * 0.15)
    for mode, (Pm, final) in res.items():
        assert abs(final - last_ref) < 5e-2 * abs(last_ref), (mode, final, last_ref)
        for k in PKEYS:
            assert np.array_equal(Pm[k][:erase_n], P[k][:erase_n]), (mode, k)          # frozen frames keep their parameters bit for bit
            # 20 Adam steps of lr .005 move a parameter by <= 0.1
            assert np.abs(Pm[k] - P_ref[k]).max() < 2e-2, (mode, k, np.abs(Pm[k] - P_ref[k]).max())
        assert np.abs(Pm['transl'] - P['transl']).max() > 1e-2                           # the free frames did move
    for k in PKEYS:                                        # chunked (resume) run == single run, bit for bit
        assert np.array_equal(stepped[k], res['fused'][0][k]), k


def test_fused_window_is_bitwise_reproducible():
    """Two identical fused runs give bitwise identical parameters: every reduction of the closure has a fixed order (per-CTA partials
    added in CTA order, gather-form adjoints, K slices of the dX GEMM added in slice order) -- Adam turns any order-dependent rounding
    of a near-zero gradient into an O(lr) parameter difference, so this matters for reproducible fits."""
    B = 24
    P, cfg = synth.make_prox_problem(B, D=32, m_scene=3000, seed=5)
    outs = []
    for _ in range(2):
        s = _reference_call_surface(B, P, cfg, maxiters=8, first_batch_flag=True)
        s['monitor'].run_fitting(s['optimizer'], s['closure'], s['final_params'], s['body_model'], pose_embedding=s['pose_embedding'],
                                 vposer=s['vposer'], use_vposer=True)
        assert s['monitor'].last_path == 'fused'
        outs.append(_params_of(s))
    for k in PKEYS:
        assert np.array_equal(outs[0][k], outs[1][k]), (k, np.abs(outs[0][k] - outs[1][k]).max())


def test_unsupported_terms_fall_back_to_the_eager_closure():
    B = 12
    P, cfg = synth.make_prox_problem(B, D=16, m_scene=500, seed=7)
    s = _reference_call_surface(B, P, cfg, maxiters=2, first_batch_flag=True)
    lbfgs_like = torch.optim.SGD(s['final_params'], lr=1e-4)
    s['monitor'].run_fitting(lbfgs_like, s['closure'], s['final_params'], s['body_model'], pose_embedding=s['pose_embedding'],
                             vposer=s['vposer'], use_vposer=True)
    assert s['monitor'].last_path.startswith('eager'), s['monitor'].last_path


def test_scan_terms_match_oracle():
    """s2m / m2s (reference fitting_temp_slide.py:638-670) on the lemo Chamfer kernels vs the brute-force oracle: values
    and the gradient on the vertices.  bs = 3 exercises the reference's batch behaviour (frame-0 vertices, see ref_prox.scan_terms)."""
    import types
    from lemo_b200.temp_prox import fitting_temp_slide as fitting
    g = torch.Generator().manual_seed(11)
    bs, V, N = 3, 700, 400
    verts = torch.randn(bs, V, 3, generator=g) * 0.4
    scan = verts[:, torch.randperm(V, generator=g)[:N]] + 0.05 * torch.randn(bs, N, 3, generator=g)
    scan_num = torch.tensor([400, 333, 250])
    vis = torch.rand(bs, V, generator=g) > 0.4
    body_mask = torch.rand(V, generator=g) > 0.2
    loss = fitting.create_loss(loss_type='smplify', s2m=True, m2s=True, rho_s2m=0.2, rho_m2s=0.5, s2m_weight=1.7, m2s_weight=0.6,
                               body_mask=body_mask.numpy(), interpenetration=False, sdf_penetration=False, contact=False,
                               use_motion_smooth_prior=False, use_friction=False, use_motion_infill_prior=False, device=DEV).to(DEV)
    v_dev = verts.to(DEV).requires_grad_(True)
    s2m, m2s = loss._scan_terms(types.SimpleNamespace(vertices=v_dev), None, scan.to(DEV), scan_num, vis.to(DEV))
    (s2m + m2s).backward()
    v_ref = verts.double().requires_grad_(True)
    r1, r2 = ref_prox.scan_terms(v_ref, scan.double(), scan_num, vis, body_mask, 0.2, 0.5, 1.7, 0.6)
    (r1 + r2).backward()
    assert abs(float(s2m) - float(r1)) < 1e-5 * max(1.0, abs(float(r1))), (float(s2m), float(r1))
    assert abs(float(m2s) - float(r2)) < 1e-5 * max(1.0, abs(float(r2))), (float(m2s), float(r2))
    assert rel(v_dev.grad.cpu().double(), v_ref.grad) < 1e-4
    assert float(v_dev.grad[1:].abs().max()) == 0.0          # the reference's pairing: only frame 0 receives gradient


def test_lbfgs_line_search_drives_the_eager_closure():
    """optim_type 'lbfgsls' (reference optimizers/optim_factory.py:50): the closure protocol -- zero_grad, forward, backward, return the
    loss -- under an optimizer that re-evaluates it several times per step; the loss must go down and the fused driver must stay out."""
    from lemo_b200.temp_prox.optimizers import create_optimizer
    B = 12
    P, cfg = synth.make_prox_problem(B, D=16, m_scene=500, seed=9)
    s = _reference_call_surface(B, P, cfg, maxiters=3, first_batch_flag=True)
    opt, _ = create_optimizer(s['final_params'], optim_type='lbfgsls', lr=1.0, maxiters=4)
    l0 = float(s['closure'](backward=False))
    final = s['monitor'].run_fitting(opt, s['closure'], s['final_params'], s['body_model'], pose_embedding=s['pose_embedding'],
                                     vposer=s['vposer'], use_vposer=True)
    assert s['monitor'].last_path.startswith('eager'), s['monitor'].last_path
    assert np.isfinite(final) and final < l0, (final, l0)
