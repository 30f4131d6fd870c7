"""Motion-smoothness prior encoder, drop-in for the reference's models/AE_sep.py `Enc`
(models/AE_sep.py:77-99; used with downsample=False, z_channel=64 at opt_amass_temp.py:137-142)."""
import ctypes as C
from collections import OrderedDict
import numpy as np
import torch
import torch.nn as nn

from .. import _lib

_CH = [(1, 32), (32, 32), (32, 64), (64, 64), (64, 64), (64, 64), (64, 64), (64, 64), (64, 64), (64, 64)]


def enc_state_keys():
    keys = []
    for blk in range(1, 6):
        for li in (0, 2):
            keys += ['enc_blc%d.main.%d.weight' % (blk, li), 'enc_blc%d.main.%d.bias' % (blk, li)]
    return keys


class _Net:
    def __init__(self, kind, in_ch, flat_w, max_n, H, W, device_index, with_backward=True):
        h = C.c_void_p()
        _lib.call('lemo_convnet_create', kind, in_ch, _lib.ptr(flat_w), flat_w.size, max_n, H, W, 1 if with_backward else 0,
                  device_index, C.byref(h))
        self.handle, self.max_n, self.H, self.W, self.stamp = h, max_n, H, W, 0

    def __del__(self):
        try:
            _lib.lib().lemo_convnet_destroy(self.handle)
        except Exception:
            pass


class _EncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x):
        x = x.contiguous().float()
        N, _, H, W = x.shape
        z = torch.empty(N, 64, H, W, device=x.device)
        _lib.call('lemo_enc_forward', net.handle, _lib.ptr(x), N, _lib.ptr(z), _lib.cur_stream(x.device))
        net.stamp += 1
        ctx.net, ctx.stamp = net, net.stamp
        ctx.save_for_backward(x)
        return z

    @staticmethod
    def backward(ctx, gz):
        (x,) = ctx.saved_tensors
        net, N = ctx.net, x.shape[0]
        st = _lib.cur_stream(x.device)
        if net.stamp != ctx.stamp:
            z = torch.empty(N, 64, x.shape[2], x.shape[3], device=x.device)
            _lib.call('lemo_enc_forward', net.handle, _lib.ptr(x), N, _lib.ptr(z), st)
            net.stamp += 1
        dx = torch.empty_like(x)
        _lib.call('lemo_enc_backward_input', net.handle, _lib.ptr(gz.contiguous().float()), N, _lib.ptr(dx), st)
        return None, dx


class Enc(nn.Module):
    """Enc(downsample=False, z_channel=64): 10x (conv3x3 pad1 + LeakyReLU 0.2) at full resolution.
    Weights are frozen on the fitting path (opt_amass_temp.py:141-142), so only the input gradient exists."""

    def __init__(self, downsample=True, z_channel=64):
        super().__init__()
        if downsample or z_channel != 64:
            raise NotImplementedError('the fitting path uses Enc(downsample=False, z_channel=64) (opt_amass_temp.py:137)')
        self._params = nn.ParameterDict()
        self._keys = enc_state_keys()
        for k, (ci, co) in zip(range(10), _CH):
            w, b = self._keys[2 * k], self._keys[2 * k + 1]
            self._params[w.replace('.', '/')] = nn.Parameter(torch.zeros(co, ci, 3, 3), requires_grad=False)
            self._params[b.replace('.', '/')] = nn.Parameter(torch.zeros(co), requires_grad=False)
        self._nets = {}

    # state_dict with the reference's key names (runs/15217/Enc_last_model.pkl loads unchanged)
    def state_dict(self, *a, **k):
        return OrderedDict((key, self._params[key.replace('.', '/')].data) for key in self._keys)

    def load_state_dict(self, sd, strict=True):
        for key in self._keys:
            v = sd[key]
            v = torch.as_tensor(np.asarray(v)) if not torch.is_tensor(v) else v
            self._params[key.replace('.', '/')].data.copy_(v)
        self._nets = {}
        return self

    def _flat(self):
        return np.concatenate([self._params[k.replace('.', '/')].detach().cpu().numpy().ravel() for k in self._keys]).astype(np.float32)

    def net(self, device, N, H, W, private=False):
        """private=True: a fresh, un-cached handle (own activation planes) for a fused fitter that bakes them into its CUDA graph."""
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if private:
            with torch.cuda.device(idx):
                return _Net(0, 1, self._flat(), N, H, W, idx)
        key = (idx, N, H, W)
        if key not in self._nets:
            with torch.cuda.device(idx):
                self._nets[key] = _Net(0, 1, self._flat(), N, H, W, idx)
        return self._nets[key]

    def forward(self, input):
        if input.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        N, _, H, W = input.shape
        z = _EncFn.apply(self.net(input.device, N, H, W), input)
        s = input.size()
        return z, s, torch.Size([N, 32, H, W]), torch.Size([N, 64, H, W]), torch.Size([N, 64, H, W]), torch.Size([N, 64, H, W])
