"""GPU parity of the fused fitting drivers against the oracle's restatement of the reference loops."""
import numpy as np
import pytest
import torch

from oracle import synth, ref_body as rb, ref_loops as rl
from gpu_common import DEV, smplx_module, vposer_module, enc_module, oracle_ctx, rel, rel_q

pytestmark = pytest.mark.gpu


def _problem(s, T, ctx):
    clean, init, contact = synth.make_sequence(s, T=T)
    with torch.no_grad():
        v, _ = rb.gen_body_mesh(torch.from_numpy(clean).to(ctx.dtype), ctx.smplx, ctx.vposer)
    return init, v[:, ctx.m67].float().numpy(), contact


def _fitter(S, T, **kw):
    from lemo_b200.fit import TemporalFitter
    return TemporalFitter(smplx_module(), vposer_module(), S, T, enc=enc_module(), device=DEV, **kw)


_CONV_MODES = {'simt': 0, 'pair': 1, 'wt': 8192}    # fp32 CUDA cores | tcgen05 pair kernel | tcgen05 weights-in-TMEM kernel (default)


@pytest.mark.parametrize('graph,conv', [(False, 'wt'), (True, 'wt'), (True, 'pair'), (True, 'simt')])
def test_temporal_first_iteration_gradients(graph, conv):
    from lemo_b200 import _lib
    _lib.call('lemo_debug_set_conv_tc', _CONV_MODES[conv])
    T, S = 119, 2
    c32, c64 = oracle_ctx(torch.float32), oracle_ctx(torch.float64)
    fit = _fitter(S, T, use_cuda_graph=graph)
    probs = [_problem(s, T, c32) for s in range(S)]
    for s, (init, mrec, con) in enumerate(probs):
        fit.set_sequence(s, init, mrec, con)
    fit.run(n_iters=1)
    st = fit.state()
    p72, losses = fit.results()
    for s, (init, mrec, con) in enumerate(probs):
        tr32, tr64 = [], []
        ref72, _ = rl.fit_temp(init, mrec, con, c32, n_iters=1, faithful=False, trace=tr32)
        rl.fit_temp(init, mrec, con, c64, n_iters=1, faithful=False, trace=tr64)
        sl = slice(s * T, (s + 1) * T)
        for k in ('g_transl', 'g_rot6d', 'g_other'):
            e32 = rel(tr32[0][k], tr64[0][k])
            e = rel(st[k][sl], tr64[0][k])
            # The smoothness term differentiates Enc features along time (|dz| ~ 1e-2 |z|), which amplifies rounding noise: the fp32
            # reference arithmetic itself is only ~6e-5 accurate on these gradients.  Budget: 5x that for the fp32 CUDA-core conv
            # path, 12x (measured 6.5x, 4e-4 of max|g|) for the bf16x3 tensor-core path whose per-activation rounding is 2^-18.
            # Kink-aware (as in test_gpu_priors.py): a pre-activation within rounding of the LeakyReLU kink flips sign with ANY change of
            # summation order upstream and moves the gradient on a (2L+1)^2 patch, so the tight budget is asserted on the 99 %
            # quantile and the maximum gets 4x the budget (measured with tools/diag_fit_grad.sh: ~10 of 6664 g_other entries at 21x e32
            # for the pair kernel once the VPoser GEMMs changed their K order -- q99 1.5e-4, median 1.5e-6, other kernels unchanged).
            budget = max((5 if conv == 'simt' else 12) * e32, 1e-4)
            eq = rel_q(st[k][sl], tr64[0][k], 0.99)
            assert eq < budget and e < 4 * budget, (k, s, e, eq, e32)
        for i, k in enumerate(['loss', 'rec', 'vposer', 'shape', 'hand', 'contact', 'smooth']):
            want = tr64[0][k]
            assert abs(float(losses[s, i]) - want) <= 2e-4 * abs(want) + 1e-9, (k, float(losses[s, i]), want)
        assert rel(p72[s], ref72) < 2e-5                    # parameters of the (first) forward incl. tgm axis-angle
    _lib.call('lemo_debug_set_conv_tc', -1)


def test_temporal_loop_tracks_oracle():
    T, n_it = 60, 12
    c32 = oracle_ctx(torch.float32)
    fit = _fitter(1, T)
    init, mrec, con = _problem(4, T, c32)
    fit.set_sequence(0, init, mrec, con)
    fit.run(n_iters=n_it, lr0=0.01, lr1=0.005, lr_switch=6)
    p72, losses = fit.results()
    tr = []
    ref72, _ = rl.fit_temp(init, mrec, con, c32, n_iters=n_it, lr_switch=6, faithful=True, trace=tr)
    assert np.abs(p72[0].cpu().numpy() - ref72).max() < 2e-3           # Adam steps are O(lr): allow a few ulps of drift per step
    assert abs(float(losses[0, 0]) - tr[-1]['loss']) < 2e-2 * abs(tr[-1]['loss'])
    assert tr[-1]['loss'] < tr[0]['loss']


def test_temporal_deterministic_across_slots_and_runs():
    """Same sequence in two slots / two runs -> bitwise identical parameters (multi-GPU determinism contract)."""
    T = 40
    c32 = oracle_ctx(torch.float32)
    init, mrec, con = _problem(2, T, c32)
    outs = []
    for _ in range(2):
        fit = _fitter(2, T)
        for s in range(2):
            fit.set_sequence(s, init, mrec, con)
        fit.run(n_iters=5)
        outs.append(fit.state())
    # third run: the batched loader (lemo_fit_set_sequences) must stage exactly what the per-sequence loader stages
    fit = _fitter(2, T)
    fit.set_sequences(np.stack([init, init]), np.stack([mrec, mrec]), np.stack([con, con]))
    fit.run(n_iters=5)
    outs.append(fit.state())
    a, b, c = outs
    for k in ('transl', 'rot6d', 'other'):
        assert torch.equal(a[k][:T], a[k][T:])
        assert torch.equal(a[k], b[k])
        assert torch.equal(a[k], c[k])


def test_perframe_tracks_oracle():
    """Per-frame stage (B=1 chain, warm start, lr .1/.01).  The L1 marker loss has sign() gradients and Adam takes O(lr)
    steps, so trajectories of two fp32 implementations separate after a few steps; parity is asserted step-exactly on a short
    run and on the loss level / decrease on the reference-length schedule."""
    from lemo_b200.fit import PerFrameFitter
    c32 = oracle_ctx(torch.float32)
    clean, _, _ = synth.make_sequence(1, T=3)
    with torch.no_grad():
        v, _ = rb.gen_body_mesh(torch.from_numpy(clean), c32.smplx, c32.vposer)
    mrec = v[:, c32.m67].numpy()
    for T, n_it, tol in ((2, 3, 3e-3), (3, 30, None)):
        fit = PerFrameFitter(smplx_module(), vposer_module(), 2, T, device=DEV)
        for s in range(2):
            fit.set_sequence(s, clean[0, 6:16], mrec[:T])
        fit.run(n_iters=n_it)
        p72, losses = fit.results()
        tr = []
        ref = rl.fit_perframe(mrec[:T], clean[0, 6:16], c32, n_frames=T, n_iters=n_it, trace=tr)
        assert torch.equal(p72[0], p72[1])                        # two slots, same problem -> bitwise identical
        if tol is not None:
            assert np.abs(p72[0].cpu().numpy() - ref).max() < tol, np.abs(p72[0].cpu().numpy() - ref).max()
        else:
            assert abs(float(losses[0, 0]) - tr[-1]) < 0.3 * abs(tr[-1]) + 1e-3, (float(losses[0, 0]), tr[-1])
            assert float(losses[0, 0]) < 0.5 * tr[0]


def test_perframe_persistent_kernel_vs_graph_path():
    """The persistent cluster kernel (default) and the per-step CUDA-graph path run the same loop: same parameters after a short run
    (they differ only in the blend arithmetic -- fp32 in the persistent kernel, TF32 tensor cores in the graph path -- and in summation
    order), the same loss level after a full frame, and every sequence slot of the persistent kernel is bitwise identical."""
    from lemo_b200 import _lib
    from lemo_b200.fit import PerFrameFitter
    c32 = oracle_ctx(torch.float32)
    clean, _, _ = synth.make_sequence(2, T=4)
    with torch.no_grad():
        v, _ = rb.gen_body_mesh(torch.from_numpy(clean), c32.smplx, c32.vposer)
    mrec = v[:, c32.m67].numpy()
    res = {}
    for mode in (1, 0):
        _lib.call('lemo_debug_set_perframe', mode)
        out = {}
        for n_it in (4, 100):
            fit = PerFrameFitter(smplx_module(), vposer_module(), 3, 4, device=DEV)
            for s in range(3):
                fit.set_sequence(s, clean[0, 6:16], mrec)
            fit.run(n_iters=n_it)
            p72, losses = fit.results()
            out[n_it] = (p72.cpu().numpy(), losses.cpu().numpy(), fit.kernel_launches())
        res[mode] = out
    _lib.call('lemo_debug_set_perframe', -1)
    assert res[1][100][2] == 1 and res[0][100][2] > 1000                     # ONE launch for 4 frames x 100 steps vs one graph per step
    a, b = res[1][4][0], res[0][4][0]
    assert np.array_equal(a[0], a[1]) and np.array_equal(a[0], a[2])         # slots are independent and identical
    assert np.abs(a - b).max() < 3e-3, np.abs(a - b).max()                   # 4 steps of lr 0.1: still together
    ref = rl.fit_perframe(mrec, clean[0, 6:16], c32, n_frames=4, n_iters=4)
    assert np.abs(a[0] - ref).max() < 3e-3, np.abs(a[0] - ref).max()
    la, lb = res[1][100][1][0], res[0][100][1][0]
    assert abs(la[0] - lb[0]) < 0.3 * abs(lb[0]) + 1e-3 and la[0] < 0.05     # same loss level at the end of the last frame
    assert np.allclose(la[2:5], lb[2:5], rtol=0.3, atol=1e-3)


def test_infill_pool_equals_single_stage():
    """InfillPool (clips of a batch fine-tuned concurrently on their own streams / handles) returns, clip by clip, exactly what one
    InfillStage returns: the fine-tune is deterministic (fixed-order split-K and weight-gradient reductions)."""
    from lemo_b200.infill import InfillStage, InfillPool, body_repr, load_infill_prior, load_infill_stats
    st64 = load_infill_stats()
    clips = []
    for s in range(3):
        b68, c68 = synth.synth_marker_clip(40 + s, T=40)
        clips.append(body_repr(torch.from_numpy(b68).to(DEV), torch.from_numpy(c68).to(DEV), stats=st64, device=DEV))
    single = InfillStage(load_infill_prior(), device=DEV, stats=st64, finetune_steps=6)
    want = [single.run(c, r) for c, r in clips]
    again = [single.run(c, r) for c, r in clips]
    pool = InfillPool(load_infill_prior(), n_streams=2, device=DEV, stats=st64, finetune_steps=6)
    got = pool.run_many([c for c, _ in clips], [r for _, r in clips])
    torch.cuda.synchronize()
    for w, a2, g in zip(want, again, got):
        assert torch.equal(w[0], a2[0])                                      # bitwise repeatable
        assert torch.equal(w[0], g[0]) and torch.equal(w[1], g[1])


def test_full_size_property_rest_pose_zero_loss():
    """Size-independent property at BASELINE's full size: targets generated from the init => rec loss 0, grads of rec term vanish,
    and the loss decreases monotonically-ish afterwards."""
    T = 119
    c32 = oracle_ctx(torch.float32)
    clean, _, con = synth.make_sequence(0, T=T)
    with torch.no_grad():
        v, _ = rb.gen_body_mesh(torch.from_numpy(clean), c32.smplx, c32.vposer)
    fit = _fitter(1, T, weights=dict(w_smooth=0.0, w_contact=0.0, w_vposer=0.0, w_hand=0.0))
    fit.set_sequence(0, clean, v[:, c32.m67].numpy(), con)
    fit.run(n_iters=1)
    _, losses = fit.results()
    assert float(losses[0, 1]) < 1e-5          # mean |marker - target|: TF32 rounding of Wt on ~1 m coordinates
