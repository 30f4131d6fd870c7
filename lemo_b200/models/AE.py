"""Motion-infilling prior, drop-in for the reference's models/AE.py `AE` (models/AE.py:79-108; built as
AE(downsample=True, in_channel=4, kernel=3) at opt_amass_perframe.py:112, opt_amass_temp.py:132).

The 40 state_dict tensors (enc_blc{1-5}.main.{0,2}.{weight,bias}, dec_blc{1-5}.deconv{1,2}.{weight,bias}) are views of ONE
flat device parameter `flat` in the reference's key order, so `optim.Adam(infill_model.parameters(), lr=3e-6)` and
`load_state_dict(weights)` from runs/59547/AE_last_model.pkl work unchanged while the CUDA side sees one contiguous vector.
"""
import ctypes as C
from collections import OrderedDict
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from .AE_sep import _Net


def ae_layout(in_channel=4):
    """[(key, shape)] in state_dict order."""
    ec = [in_channel, 32, 64, 128, 256, 256]
    dc = [256, 256, 128, 64, 32, 1]
    out = []
    for i in range(5):
        out += [('enc_blc%d.main.0.weight' % (i + 1), (ec[i + 1], ec[i], 3, 3)), ('enc_blc%d.main.0.bias' % (i + 1), (ec[i + 1],)),
                ('enc_blc%d.main.2.weight' % (i + 1), (ec[i + 1], ec[i + 1], 3, 3)), ('enc_blc%d.main.2.bias' % (i + 1), (ec[i + 1],))]
    for b in range(5):
        out += [('dec_blc%d.deconv1.weight' % (b + 1), (dc[b], dc[b + 1], 3, 3)), ('dec_blc%d.deconv1.bias' % (b + 1), (dc[b + 1],)),
                ('dec_blc%d.deconv2.weight' % (b + 1), (dc[b + 1], dc[b + 1], 3, 3)), ('dec_blc%d.deconv2.bias' % (b + 1), (dc[b + 1],))]
    return out


class _AEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, flat):
        x = x.contiguous().float()
        N, _, H, W = x.shape
        st = _lib.cur_stream(x.device)
        _lib.call('lemo_convnet_set_weights', net.handle, _lib.ptr(flat.detach().contiguous()), st)
        rec = torch.empty(N, 1, H, W, device=x.device)
        zs = net.z_shape
        z = torch.empty(N, 256, zs[0], zs[1], device=x.device)
        _lib.call('lemo_ae_forward', net.handle, _lib.ptr(x), N, _lib.ptr(rec), _lib.ptr(z), st)
        net.stamp += 1
        ctx.net, ctx.stamp = net, net.stamp
        ctx.save_for_backward(x, flat)
        ctx.mark_non_differentiable(z)
        return rec, z

    @staticmethod
    def backward(ctx, g_rec, _gz):
        x, flat = ctx.saved_tensors
        net, N = ctx.net, x.shape[0]
        st = _lib.cur_stream(x.device)
        if net.stamp != ctx.stamp:
            raise RuntimeError('AE handle was re-used by another forward before backward; call backward first')
        dw = torch.empty_like(flat)
        _lib.call('lemo_ae_backward_weights', net.handle, _lib.ptr(g_rec.contiguous().float()), N, _lib.ptr(dw), st)
        return None, None, dw


class AE(nn.Module):
    def __init__(self, downsample=True, in_channel=1, kernel=3):
        super().__init__()
        if not downsample or kernel != 3:
            raise NotImplementedError('the fitting path uses AE(downsample=True, kernel=3) (opt_amass_perframe.py:112)')
        self.in_channel = in_channel
        self._layout = ae_layout(in_channel)
        n = sum(int(np.prod(s)) for _, s in self._layout)
        self.flat = nn.Parameter(torch.zeros(n))
        self._nets = {}

    def _views(self):
        out, off = OrderedDict(), 0
        for k, s in self._layout:
            n = int(np.prod(s))
            out[k] = self.flat.data[off:off + n].view(s)
            off += n
        return out

    def state_dict(self, *a, **k):
        return OrderedDict((k_, v.clone()) for k_, v in self._views().items())

    def load_state_dict(self, sd, strict=True):
        with torch.no_grad():
            for k, v in self._views().items():
                s = sd[k]
                v.copy_(torch.as_tensor(np.asarray(s)) if not torch.is_tensor(s) else s)
        return self

    def net(self, device, N, H, W):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, N, H, W)
        if key not in self._nets:
            with torch.cuda.device(idx):
                net = _Net(1, self.in_channel, self.flat.detach().cpu().numpy().astype(np.float32), N, H, W, idx)
            h, w = H, W
            for _ in range(5):
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            net.z_shape = (h, w)
            self._nets[key] = net
        return self._nets[key]

    def forward(self, input):
        if input.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        N, _, H, W = input.shape
        return _AEFn.apply(self.net(input.device, N, H, W), input, self.flat)

    @torch.no_grad()
    def finetune(self, clip_img_input, row_ids, steps=60, lr=3e-6):
        """The self-supervised fine-tune loop of opt_amass_perframe.py:152-173 fused on device: `steps` x (forward, L1 on the
        rows `row_ids` of channel 0, weight backward, Adam).  Returns the per-step losses [steps] (device tensor)."""
        x = clip_img_input.contiguous().float()
        N, _, H, W = x.shape
        net = self.net(x.device, N, H, W)
        st = _lib.cur_stream(x.device)
        _lib.call('lemo_convnet_set_weights', net.handle, _lib.ptr(self.flat.detach().contiguous()), st)
        mask = torch.zeros(H, device=x.device)
        rows = row_ids.to(x.device).long() if torch.is_tensor(row_ids) else torch.as_tensor(np.asarray(row_ids), device=x.device).long()
        mask[rows] = 1.0
        # one C-ABI call for the whole loop: the step is captured once as a CUDA graph and replayed (lemo_ae_finetune_run); the mask /
        # loss buffers are kept on the handle so that the captured pointers stay valid from clip to clip
        if getattr(net, '_ft_mask', None) is None or net._ft_mask.shape[0] != H or net._ft_losses.shape[0] != steps:
            net._ft_mask = torch.zeros(H, device=x.device)
            net._ft_losses = torch.zeros(steps, device=x.device)
        net._ft_mask.copy_(mask)
        _lib.call('lemo_ae_finetune_run', net.handle, _lib.ptr(x), _lib.ptr(net._ft_mask), int(rows.numel()), N, float(lr), int(steps),
                  _lib.ptr(net._ft_losses), st)
        losses = net._ft_losses.clone()
        x.record_stream(torch.cuda.current_stream(x.device))
        _lib.call('lemo_convnet_get_weights', net.handle, _lib.ptr(self.flat.data), st)
        net.stamp += 1
        return losses
