"""GPU parity of the infill pre-stage against oracle/ref_infill.py and the golden vectors recorded from the
reference's own get_local_markers_4chan / reconstruct_global_body."""
import os
import numpy as np
import pytest
import torch

from oracle import ref_infill as ri, ref_priors as rp
from oracle.make_golden import rng_state_dict, AE_SHAPES, synth_marker_clip
from gpu_common import DEV

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden_infill.npz')


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(GOLD))


def test_body_repr_matches_reference_golden(gold):
    from lemo_b200.infill import body_repr, load_infill_stats
    from lemo_b200.utils.utils import get_local_markers_4chan
    for tag in ('a', 'b'):
        rep, rot0 = get_local_markers_4chan(gold['body_' + tag], gold['contact_' + tag], device=DEV)
        ref = gold['repr_' + tag]                                                  # [4, T-1, 208] float64 from the reference
        assert tuple(rep.shape) == ref.shape
        # double arithmetic on the device, float32 on output: the only error is the final rounding
        assert np.abs(rep.cpu().numpy().astype(np.float64) - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())
        assert abs(float(rot0[0]) - float(gold['rot0_' + tag][0])) < 1e-12
        assert np.array_equal(rep[0, :, -4:].cpu().numpy(), gold['contact_' + tag][:-1])         # labels pass through bit-exactly
        # normalised wire format [4,208,T-1]
        st = load_infill_stats()
        img, _ = body_repr(gold['body_' + tag], gold['contact_' + tag], stats=st, device=DEV)
        want = ref.copy()
        want[0] = (want[0] - st[0:208]) / st[208:416]
        want[1:3] = (want[1:3] - st[416]) / st[417]
        want[3] = (want[3] - st[418]) / st[419]
        want = want.transpose(0, 2, 1)
        assert tuple(img.shape) == want.shape
        assert np.abs(img.cpu().numpy() - want).max() <= 2e-6 * np.abs(want).max()


def test_reconstruct_global_body_matches_reference_golden(gold):
    from lemo_b200.utils.utils import reconstruct_global_body
    for tag in ('a', 'b'):
        packed32 = gold['packed_' + tag].astype(np.float32)
        out = reconstruct_global_body(packed32, gold['rot0_' + tag], device=DEV).cpu().numpy()
        want = ri.reconstruct_global_body(packed32, gold['rot0_' + tag])           # oracle on the same float32-rounded inputs
        assert np.abs(out - want).max() < 2e-6
        assert np.abs(out - gold['global_' + tag]).max() < 2e-5                   # and the reference's float64-input output


def test_body_repr_round_trip_full_clip():
    """size-independent property at a long clip: representation -> reconstruction returns the markers (floor- and origin-shifted)."""
    from lemo_b200.infill import body_repr
    from lemo_b200.utils.utils import reconstruct_global_body
    body, contact = synth_marker_clip(17, T=600)
    rep, rot0 = body_repr(body, contact, stats=None, device=DEV)
    T = rep.shape[2]
    packed = torch.zeros(T, 70, 3, device=DEV)
    packed[:, 1:69] = rep[0, :204].t().reshape(T, 68, 3)
    packed[:, 69, 0], packed[:, 69, 1], packed[:, 69, 2] = rep[1, 0], rep[2, 0], rep[3, 0]
    glob = reconstruct_global_body(packed, rot0, device=DEV).cpu().numpy()
    want = body.astype(np.float64)
    want[:, :, 2] -= np.float64(body[:, :, 2].min())
    want[:, :, 0:2] -= want[0, 0, 0:2].copy()
    assert np.abs(glob - want[:-1]).max() < 5e-5


def test_prepare_input_bit_exact():
    from lemo_b200.infill import InfillStage
    from lemo_b200.models.AE import AE
    g = np.random.default_rng(3)
    clip = g.standard_normal((4, 208, 119)).astype(np.float32)
    stage = InfillStage(AE(downsample=True, in_channel=4, kernel=3), device=DEV, finetune_steps=0)
    _, xp, mask = stage.prepare(clip)
    assert np.array_equal(xp[0].cpu().numpy(), ri.prepare_input(clip))
    _, loss_rows = ri.mask_rows(208)
    assert np.array_equal(torch.nonzero(mask > 0.5).flatten().cpu().numpy(), loss_rows)


def test_infill_stage_matches_oracle_pipeline():
    """mask -> pad -> 3 fine-tune steps -> inference -> crop / labels / de-normalise / reconstruct, against the reference op sequence
    on CPU (torch AE restatement + numpy float64 post-processing).  Random AE weights, synthetic clip."""
    from lemo_b200.infill import InfillStage, body_repr, load_infill_stats
    from lemo_b200.models.AE import AE
    body, contact = synth_marker_clip(9, T=41)
    stats = load_infill_stats()
    clip, rot0 = body_repr(body, contact, stats=stats, device=DEV)
    ae = AE(downsample=True, in_channel=4, kernel=3)
    w0 = rng_state_dict(AE_SHAPES, 41)
    ae.load_state_dict(w0)
    stage = InfillStage(ae, device=DEV, finetune_steps=3, lr=3e-6)
    m_rec, con, m_in, losses = stage.run(clip, rot0, return_losses=True)
    # ---- oracle
    clip_np = clip.cpu().numpy()
    xp = torch.from_numpy(ri.prepare_input(clip_np))[None]
    _, loss_rows = ri.mask_rows(208)
    sd = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in w0.items()}
    opt = torch.optim.Adam(list(sd.values()), lr=3e-6)
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        rec, _ = rp.ae_forward(xp, sd)
        loss = (rec[:, 0] - xp[:, 0])[:, loss_rows].abs().mean()
        loss.backward()
        opt.step()
        ref_losses.append(float(loss.detach()))
    with torch.no_grad():
        rec, _ = rp.ae_forward(xp, sd)
    st = ri.load_stats()
    want_rec, want_con, want_in = ri.finalize(rec[0, 0].numpy(), clip_np, st, rot0.cpu().numpy())
    assert np.allclose(losses.cpu().numpy(), ref_losses, rtol=2e-4)
    # the un-infilled path involves no network: float32 rounding of the output only
    assert np.abs(m_in.cpu().numpy() - want_in).max() < 2e-6 * max(1.0, np.abs(want_in).max())
    # infilled markers: AE forward parity (1e-5 of the image range, x Xstd_local) through the rigid reconstruction
    scale = np.abs(want_rec).max()
    assert np.abs(m_rec.cpu().numpy() - want_rec).max() < 1e-4 * scale
    flips = np.sum(con.cpu().numpy() != want_con)
    assert flips <= 1                                                               # a logit within 1e-5 of zero may flip
    # fine-tuning must not leak into the next clip: the weights are restored before every run
    # (the weight-gradient kernels accumulate with float atomics, so two runs agree to rounding, not bitwise)
    m2, _, _ = stage.run(clip, rot0)
    assert torch.allclose(m2, m_rec, atol=1e-5 * scale)
