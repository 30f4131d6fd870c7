"""Fused fitting drivers: the Adam inner loops of the reference scripts, on device.

    temporal stage   opt_amass_temp.py:329-455      -> TemporalFitter.run(...)
    per-frame stage  opt_amass_perframe.py:293-361  -> PerFrameFitter.run(...)

A fitter holds S independent sequences of T frames side by side on one GPU.  One `run()` is a single C-ABI call
(lemo_fit_run / lemo_fit_run_perframe): every iteration -- 6D->R, VPoser decode, SMPL-X on the 253 loss rows,
marker / smoothness (Enc) / contact-velocity / prior losses, backward, Adam with the script's LR schedule --
is enqueued on the current stream (optionally replayed as a CUDA graph) with no host synchronisation.
"""
import ctypes as C
import os
import numpy as np
import torch

from . import _lib
from .smplx import SMPLX
from .vposer import VPoserDecoder
from .models.AE_sep import Enc

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'assets')


def load_tables():
    """Index tables + normalisation statistics exported from the reference's data files (tools/export_assets.py)."""
    return dict(np.load(os.path.join(_ASSETS, 'lemo_tables.npz')))


def load_smooth_prior():
    """Enc with the reference's shipped smoothness-prior weights (runs/15217/Enc_last_model.pkl)."""
    enc = Enc(downsample=False, z_channel=64)
    enc.load_state_dict(dict(np.load(os.path.join(_ASSETS, 'enc_smooth_15217.npz'))))
    return enc


# argparse defaults of the scripts (opt_amass_temp.py:46-51, opt_amass_perframe.py:40-43)
TEMP_WEIGHTS = dict(w_rec=1.0, w_contact=0.03, w_smooth=1e6, w_vposer=0.02, w_shape=0.01, w_hand=0.01)
PERFRAME_WEIGHTS = dict(w_rec=1.0, w_contact=0.0, w_smooth=0.0, w_vposer=0.02, w_shape=0.01, w_hand=0.01)


class _Fitter:
    MODE = 0

    def __init__(self, smplx_model: SMPLX, vposer: VPoserDecoder, n_seq, n_frames, enc: Enc = None, device='cuda',
                 weights=None, tables=None, use_cuda_graph=True):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', idx)
        self.S, self.T = n_seq, n_frames
        self.B = n_seq * n_frames if self.MODE == 0 else n_seq
        tables = tables or load_tables()
        w = dict(TEMP_WEIGHTS if self.MODE == 0 else PERFRAME_WEIGHTS)
        w.update(weights or {})
        self.weights = w
        self._keep = []
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        f32 = lambda a: np.ascontiguousarray(a, np.float32)
        cfg = _lib.LemoFitConfigC()
        cfg.mode, cfg.n_seq, cfg.n_frames = self.MODE, n_seq, n_frames
        for k, v in w.items():
            setattr(cfg, k, float(v))
        cfg.vel_thres, cfg.fps = 0.1, 30.0
        m67, m81 = i32(tables['markers67']), i32(tables['markers81'])
        foot = [i32(tables[k]) for k in ('left_heel', 'right_heel', 'left_toe', 'right_toe')]
        mean, std = f32(tables['smooth_Xmean']), f32(tables['smooth_Xstd'])
        self._keep += [m67, m81, mean, std] + foot
        cfg.h_markers67, cfg.h_markers81 = m67.ctypes.data, m81.ctypes.data
        for p in range(4):
            cfg.h_foot_ids[p] = foot[p].ctypes.data
            cfg.n_foot[p] = foot[p].shape[0]
        cfg.h_smooth_mean, cfg.h_smooth_std = mean.ctypes.data, std.ctypes.data
        cfg.use_cuda_graph = 1 if use_cuda_graph else 0
        with torch.cuda.device(idx):
            self._dmodel = smplx_model.device_model(self.device)
            self._vp = vposer.handle(self.device, self.B, private=True)      # private scratch: baked into this fitter's graph
            self._enc = None
            if self.MODE == 0 and enc is not None and w['w_smooth'] > 0:
                self._enc = enc.net(self.device, n_seq, 245, n_frames - 1 + 16, private=True)
            h = C.c_void_p()
            _lib.call('lemo_fit_create', self._dmodel.handle, self._vp.handle, None,
                      self._enc.handle if self._enc else None, C.byref(cfg), idx, C.byref(h))
        self.handle = h
        self.iters_run = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().lemo_fit_destroy(self.handle)
        except Exception:
            pass

    def _dev(self, a, shape):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).to(self.device, torch.float32).contiguous()
        assert tuple(t.shape) == tuple(shape), 'expected shape %s, got %s' % (shape, tuple(t.shape))
        return t

    def results(self):
        """(params72 [S,T,72] of the last forward -- what the scripts np.save --, losses [S,8])."""
        p = torch.empty(self.S, self.T, 72, device=self.device)
        l = torch.empty(self.S, 8, device=self.device)
        _lib.call('lemo_fit_get', self.handle, _lib.ptr(p), _lib.ptr(l), _lib.cur_stream(self.device))
        return p, l

    def state(self):
        """Raw optimisation state after the last step + gradients of the last iteration (for parity tests)."""
        B, d = self.B, self.device
        out = [torch.empty(B, n, device=d) for n in (3, 6, 56, 3, 6, 56)]
        _lib.call('lemo_fit_get_state', self.handle, *[_lib.ptr(t) for t in out], _lib.cur_stream(d))
        return dict(zip(['transl', 'rot6d', 'other', 'g_transl', 'g_rot6d', 'g_other'], out))

    def kernel_launches(self):
        return int(_lib.lib().lemo_fit_kernel_launches(self.handle))


class TemporalFitter(_Fitter):
    """opt_amass_temp.py:329-455 for S sequences at once."""
    MODE = 0

    def set_sequence(self, s, init72, markers_rec, contact, sync=True):
        """init72 [T,72] (per-frame stage result), markers_rec [T,67,3] (infilled targets), contact [T,4]."""
        T = self.T
        a, b, c = self._dev(init72, (T, 72)), self._dev(markers_rec, (T, 67, 3)), self._dev(contact, (T, 4))
        _lib.call('lemo_fit_set_sequence', self.handle, s, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.cur_stream(self.device))
        if sync:
            torch.cuda.current_stream(self.device).synchronize()      # inputs may be freed by the caller after return
        else:
            for t in (a, b, c):
                t.record_stream(torch.cuda.current_stream(self.device))

    def set_sequences(self, init72, markers_rec, contact):
        """All S sequences at once from [S,T,72], [S,T,67,3], [S,T,4] (host or device); no host synchronisation."""
        S, T = self.S, self.T
        a, b, c = self._dev(init72, (S, T, 72)), self._dev(markers_rec, (S, T, 67, 3)), self._dev(contact, (S, T, 4))
        _lib.call('lemo_fit_set_sequences', self.handle, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.cur_stream(self.device))
        for t in (a, b, c):
            t.record_stream(torch.cuda.current_stream(self.device))

    def run(self, n_iters=100, lr0=0.01, lr1=0.005, lr_switch=60):
        """total_steps=100, lr .01 -> .005 after step 60 (opt_amass_temp.py:343-352).  Asynchronous."""
        _lib.call('lemo_fit_run', self.handle, n_iters, lr0, lr1, lr_switch, _lib.cur_stream(self.device))
        self.iters_run += n_iters


class PerFrameFitter(_Fitter):
    """opt_amass_perframe.py:293-361 for S sequences at once (each a chain of T warm-started B=1 problems)."""
    MODE = 1

    def set_sequence(self, s, betas, markers_rec):
        T = self.T
        init = torch.zeros(1, 72)
        init[0, 6:16] = torch.as_tensor(np.asarray(betas), dtype=torch.float32).view(10)
        a, b = self._dev(init, (1, 72)), self._dev(markers_rec, (T, 67, 3))
        _lib.call('lemo_fit_set_sequence', self.handle, s, _lib.ptr(a), _lib.ptr(b), None, _lib.cur_stream(self.device))
        torch.cuda.current_stream(self.device).synchronize()

    def run(self, n_iters=100):
        _lib.call('lemo_fit_run_perframe', self.handle, n_iters, _lib.cur_stream(self.device))
        self.iters_run += n_iters * self.T
