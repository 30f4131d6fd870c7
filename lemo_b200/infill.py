"""Infill pre-stage of the temporal fitting script, on device (opt_amass_temp.py:141-325).

    clip of world-space markers  --body_repr-->  [4,208,T-1] normalised image  --run-->  infilled world-space markers [T-1,67,3]
                                                                                         + contact labels [T-1,4]

`run` = mask + reflect pad, 60 self-supervised fine-tune steps of the infill AE (fused, models/AE.py), inference, crop, contact labels,
de-normalisation and reconstruct_global_body -- the clip never leaves the GPU (the reference goes GPU -> numpy float64 -> GPU here).
The outputs are what `TemporalFitter.set_sequence(s, init72, markers_rec, contact)` consumes.
"""
import os
import numpy as np
import torch

from . import _lib
from .models.AE import AE

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')
D_ROWS = 208


def load_infill_stats():
    """The six normalisation statistics of preprocess_stats_infill_local_markers_4chan.npz as one float64 vector of 420 values:
    Xmean_local[208], Xstd_local[208], Xmean_global_xy, Xstd_global_xy, Xmean_global_r, Xstd_global_r (include/lemo_b200.h)."""
    t = np.load(os.path.join(_ASSETS, 'lemo_tables.npz'))
    return np.concatenate([t['infill_Xmean_local'], t['infill_Xstd_local'],
                           [t['infill_Xmean_global_xy'], t['infill_Xstd_global_xy'], t['infill_Xmean_global_r'], t['infill_Xstd_global_r']]]
                          ).astype(np.float64)


def load_infill_prior():
    """AE(downsample=True, in_channel=4, kernel=3) with the reference's shipped infill-prior weights (runs/59547/AE_last_model.pkl)."""
    ae = AE(downsample=True, in_channel=4, kernel=3)
    ae.load_state_dict(dict(np.load(os.path.join(_ASSETS, 'ae_infill_59547.npz'))))
    return ae


def _f32(a, device):
    t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
    return t.to(device, torch.float32).contiguous()


def body_repr(body, contact, stats=None, device='cuda'):
    """body [T,68,3] (pelvis joint + 67 SSM2 markers, z up), contact [T,4] -> (clip_img [4,208,T-1] float32 on `device`,
    rot_0_pivot double[1] on `device`).  stats: 420 float64 values (load_infill_stats()) or None for the un-normalised representation.
    utils/utils.py:209-265 + loader/optimize_loader_amass_new.py:359-361,376-377."""
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
    b, c = _f32(body, device), _f32(contact, device)
    T = b.shape[0]
    assert tuple(b.shape) == (T, 68, 3) and tuple(c.shape) == (T, 4), 'expected body [T,68,3] and contact [T,4]'
    st = None if stats is None else torch.as_tensor(np.asarray(stats, np.float64)).to(device)
    rep = torch.empty(4, D_ROWS, T - 1, device=device)
    rot0 = torch.empty(1, device=device, dtype=torch.float64)
    ws = torch.empty(8 * T, device=device, dtype=torch.float64)
    _lib.call('lemo_repr_local_markers_4chan', _lib.ptr(b), _lib.ptr(c), T, None if st is None else _lib.ptr(st), _lib.ptr(rep),
              _lib.ptr(rot0), _lib.ptr(ws), _lib.cur_stream(device))
    for t in (b, c, ws) + (() if st is None else (st,)):
        t.record_stream(torch.cuda.current_stream(device))
    return rep, rot0


class InfillStage:
    """The per-clip loop body of opt_amass_temp.py:152-215 and :262-325 for one GPU.

        stage = InfillStage(load_infill_prior(), device='cuda:0')
        markers_rec, contact, markers_in = stage.run(clip_img, rot_0_pivot)         # clip_img [4,208,T] from body_repr / the loader
    """

    def __init__(self, ae: AE, device='cuda', stats=None, finetune_steps=60, lr=3e-6):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        self.ae = ae.to(self.device)
        self.steps, self.lr = finetune_steps, lr
        self.stats = torch.as_tensor(np.asarray(load_infill_stats() if stats is None else stats, np.float64)).to(self.device)
        self._pristine = self.ae.flat.detach().clone()          # `infill_model.load_state_dict(weights)` before every clip (:160)

    def prepare(self, clip_img):
        """clip_img [4,208,T] -> (x_pad [1,4,210,T+16], loss row indices into the padded image)"""
        x = _f32(clip_img, self.device)
        assert x.dim() == 3 and x.shape[0] == 4 and x.shape[1] == D_ROWS, 'expected clip_img [4,208,T]'
        T = x.shape[2]
        xp = torch.empty(1, 4, D_ROWS + 2, T + 16, device=self.device)
        mask = torch.empty(D_ROWS + 2, device=self.device)
        _lib.call('lemo_infill_prepare_input', _lib.ptr(x), D_ROWS, T, _lib.ptr(xp), _lib.ptr(mask), None, _lib.cur_stream(self.device))
        return x, xp, mask

    def run(self, clip_img, rot_0_pivot, return_losses=False):
        """-> (markers_rec [T,67,3], contact_lbl_rec [T,4], markers_input [T,67,3]) float32 on the device; T = clip_img.shape[-1]."""
        x, xp, mask = self.prepare(clip_img)
        T = x.shape[2]
        with torch.no_grad():
            self.ae.flat.data.copy_(self._pristine)
            rows = torch.nonzero(mask > 0.5).flatten()
            losses = self.ae.finetune(xp, rows, steps=self.steps, lr=self.lr) if self.steps > 0 else None
            rec, _ = self.ae(xp)                                   # eval-mode forward on the same padded input (:206-209)
        rot0 = torch.as_tensor(rot_0_pivot, dtype=torch.float64).reshape(1).to(self.device) if not torch.is_tensor(rot_0_pivot) \
            else rot_0_pivot.to(self.device, torch.float64).reshape(1)
        m_rec = torch.empty(T, 67, 3, device=self.device)
        m_in = torch.empty(T, 67, 3, device=self.device)
        con = torch.empty(T, 4, device=self.device)
        ws = torch.empty(8 * T, device=self.device, dtype=torch.float64)
        rec = rec.contiguous()
        _lib.call('lemo_infill_finalize', _lib.ptr(rec), _lib.ptr(x), _lib.ptr(self.stats), _lib.ptr(rot0), D_ROWS, T, _lib.ptr(m_rec),
                  _lib.ptr(con), _lib.ptr(m_in), _lib.ptr(ws), _lib.cur_stream(self.device))
        for t in (rec, x, rot0, ws):
            t.record_stream(torch.cuda.current_stream(self.device))
        return (m_rec, con, m_in, losses) if return_losses else (m_rec, con, m_in)


class InfillPool:
    """S InfillStages with their own AE weights, scratch and CUDA stream: the fine-tune of one clip under-fills a B200 (deep AE levels
    are 16x16 planes), so the clips of a batch are fine-tuned CONCURRENTLY -- each stage's captured step graph replays on its own
    stream and the SMs interleave them.  Same per-clip arithmetic as InfillStage.run (bitwise: every clip has its own handle)."""

    def __init__(self, ae: AE, n_streams=8, device='cuda', stats=None, finetune_steps=60, lr=3e-6):
        self.device = torch.device(device)
        sd = ae.state_dict()
        self.stages, self.streams = [], []
        for _ in range(max(1, n_streams)):
            m = AE(downsample=True, in_channel=ae.in_channel, kernel=3)
            m.load_state_dict(sd)
            self.stages.append(InfillStage(m, device=self.device, stats=stats, finetune_steps=finetune_steps, lr=lr))
            self.streams.append(torch.cuda.Stream(device=self.device))

    def run_many(self, clip_imgs, rot0s):
        """clip_imgs: sequence of [4,208,T] tensors; rot0s: sequence of rot_0_pivot values.  -> list of (markers_rec, contact, markers_in),
        in input order.  The caller's current stream waits for all of them.

        Measured (tools/ab_infill_pool.py): 49 ms per clip for 4, 8, 12, 16 or 24 clips in flight, and the same when every stage is driven
        from its own host thread -- the pool is bound by the rate at which the device accepts graph nodes (~190 per fine-tune step,
        ~4 us each across all streams), not by SM time and not by the issuing thread.  Fewer, fatter nodes per step is what moves it."""
        cur = torch.cuda.current_stream(self.device)
        outs = [None] * len(clip_imgs)
        for st in self.streams:
            st.wait_stream(cur)
        for i, (clip, rot0) in enumerate(zip(clip_imgs, rot0s)):
            k = i % len(self.stages)
            with torch.cuda.stream(self.streams[k]):
                outs[i] = self.stages[k].run(clip, rot0)
        for st in self.streams:
            cur.wait_stream(st)
        for o in outs:
            for t in o:
                t.record_stream(cur)
        return outs
