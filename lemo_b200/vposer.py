"""VPoser decoder behind the reference's `vposer_model.decode(Z, output_type='aa')` call
(human_body_prior/train/vposer_smpl.py:107-121; used at utils/utils.py:148,163)."""
import ctypes as C
import numpy as np
import torch
import torch.nn as nn

from . import _lib

_KEYS = ['dec_fc1_w', 'dec_fc1_b', 'dec_fc2_w', 'dec_fc2_b', 'dec_out_w', 'dec_out_b']
_SD_KEYS = ['bodyprior_dec_fc1.weight', 'bodyprior_dec_fc1.bias', 'bodyprior_dec_fc2.weight', 'bodyprior_dec_fc2.bias',
            'bodyprior_dec_out.weight', 'bodyprior_dec_out.bias']


class _Handle:
    def __init__(self, weights, max_batch, device_index):
        h = C.c_void_p()
        w = [np.ascontiguousarray(weights[k], np.float32) for k in _KEYS]
        _lib.call('lemo_vposer_create', *[_lib.ptr(a) for a in w], max_batch, device_index, C.byref(h))
        self.handle, self.max_batch, self.stamp = h, max_batch, 0

    def __del__(self):
        try:
            _lib.lib().lemo_vposer_destroy(self.handle)
        except Exception:
            pass


class _Decode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hnd, z, want_aa):
        z = z.contiguous().float()
        B = z.shape[0]
        R = torch.empty(B, 21, 9, device=z.device)
        aa = torch.empty(B, 63, device=z.device) if want_aa else None
        _lib.call('lemo_vposer_decode', hnd.handle, _lib.ptr(z), B, _lib.ptr(R), _lib.ptr(aa), _lib.cur_stream(z.device))
        hnd.stamp += 1
        ctx.hnd, ctx.stamp, ctx.want_aa = hnd, hnd.stamp, want_aa
        if want_aa:
            ctx.save_for_backward(z, R)
        else:
            ctx.save_for_backward(z)
            aa = torch.empty(0, device=z.device)
            ctx.mark_non_differentiable(aa)
        return R, aa

    @staticmethod
    def backward(ctx, gR, gaa):
        z = ctx.saved_tensors[0]
        hnd, B = ctx.hnd, z.shape[0]
        st = _lib.cur_stream(z.device)
        if ctx.want_aa and gaa is not None:
            # aa = tgm.rotation_matrix_to_angle_axis(R): adjoint through the selected quaternion branch, as autograd does in the
            # reference (vposer_smpl.py:152-161); the gradient reaches z exactly as in `vposer.decode(z,'aa') -> body_model(body_pose=)`
            R = ctx.saved_tensors[1]
            dR = torch.empty_like(R)
            _lib.call('lemo_rotmat_to_aa_backward', _lib.ptr(R), _lib.ptr(gaa.contiguous().float()), B * 21, _lib.ptr(dR), st)
            gR = dR if gR is None else gR + dR
        if gR is None:
            return None, torch.zeros_like(z), None
        if hnd.stamp != ctx.stamp:      # activations were overwritten by a later decode: recompute ours
            R = torch.empty(B, 21, 9, device=z.device)
            _lib.call('lemo_vposer_decode', hnd.handle, _lib.ptr(z), B, _lib.ptr(R), None, st)
            hnd.stamp += 1
            ctx.stamp = hnd.stamp
        dz = torch.empty_like(z)
        _lib.call('lemo_vposer_decode_backward', hnd.handle, _lib.ptr(z), B, _lib.ptr(gR.contiguous().float()), _lib.ptr(dz), st)
        return None, dz, None


class VPoserDecoder(nn.Module):
    """decode-only VPoser (32 -> 512 -> 512 -> 126 -> 6D Gram-Schmidt).  `weights`: dict with dec_fc1_w/... arrays
    (nn.Linear layout) or a state_dict with bodyprior_dec_* keys."""
    latentD = 32
    num_joints = 21

    def __init__(self, weights):
        super().__init__()
        if 'bodyprior_dec_fc1.weight' in weights:
            weights = {k: (weights[s].detach().cpu().numpy() if torch.is_tensor(weights[s]) else weights[s]) for k, s in zip(_KEYS, _SD_KEYS)}
        self._w = {k: np.ascontiguousarray(weights[k], np.float32) for k in _KEYS}
        self._handles = {}
        self._dummy = nn.Parameter(torch.zeros(1), requires_grad=False)   # lets .to(device) / device queries work

    def handle(self, device, B, private=False):
        """Scratch handle for batch B on `device`.  private=True returns a fresh handle that is NOT cached: a fused fitter bakes the
        handle's activation buffers into its CUDA graph and replays it on its own stream, so it must not share them with other
        fitters or with plain decode() calls of the same batch size."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if private:
            with torch.cuda.device(idx):
                return _Handle(self._w, B, idx)
        key = (idx, B)
        if key not in self._handles:
            with torch.cuda.device(idx):
                self._handles[key] = _Handle(self._w, B, idx)
        return self._handles[key]

    def decode(self, Zin, output_type='matrot'):
        assert output_type in ['matrot', 'aa']
        if Zin.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        R, aa = _Decode.apply(self.handle(Zin.device, Zin.shape[0]), Zin, output_type == 'aa')
        if output_type == 'aa':
            # differentiable like the reference's decode(Z,'aa') (adjoint of the tgm conversion: lemo_rotmat_to_aa_backward)
            return aa.view(Zin.shape[0], 1, 21, 3)
        return R.view(Zin.shape[0], 1, 21, 9)

    def decode_matrot(self, Zin):
        return self.decode(Zin, 'matrot')
