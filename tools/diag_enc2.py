import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import synth
from lemo_b200.fit import load_smooth_prior
from lemo_b200 import _lib
dev = 'cuda:0'
enc = load_smooth_prior().to(dev)
sd = {k: torch.from_numpy(v).double() for k, v in synth.load_enc_weights().items()}
keys = [('enc_blc%d.main.%d' % (b, li)) for b in range(1, 6) for li in (0, 2)]
N, H, W = 2, 37, 53
x = torch.from_numpy((0.5 * np.random.default_rng(9).standard_normal((N, 1, H, W))).astype(np.float32))
gz = torch.from_numpy(np.random.default_rng(10).standard_normal((N, 64, H, W)).astype(np.float32))
h = x.double(); pres = []
for k in keys:
    pre = F.conv2d(h, sd[k + '.weight'], sd[k + '.bias'], padding=1); pre.requires_grad_(True); pre.retain_grad()
    pres.append(pre); h = F.leaky_relu(pre, 0.2)
# manual chain so each pre is a leaf: recompute properly with autograd
h = x.double().requires_grad_(True); acts = []; pre_list = []
hh = h
for k in keys:
    pre = F.conv2d(hh, sd[k + '.weight'], sd[k + '.bias'], padding=1); pre.retain_grad(); pre_list.append(pre)
    hh = F.leaky_relu(pre, 0.2)
(hh * gz.double()).sum().backward()
xg = x.to(dev)
z = enc(xg)[0]
net = enc.net(torch.device(dev), N, H, W)
print('fwd err', float((z.cpu().double() - hh.detach()).abs().max() / hh.abs().max()))
for l in range(9, -1, -1):
    C = pre_list[l].shape[1]
    out = torch.empty(N, C, H, W, device=dev)
    _lib.call('lemo_enc_debug_backward', net.handle, _lib.ptr(gz.to(dev).contiguous()), N, l, _lib.ptr(out), _lib.cur_stream())
    ref = pre_list[l].grad
    e = (out.cpu().double() - ref).abs()
    am = np.unravel_index(int(e.argmax()), e.shape)
    print('layer', l, 'C', C, 'rel err %.3e' % float(e.max() / ref.abs().max()), 'argmax', tuple(int(a) for a in am),
          'n_bad', int((e > 1e-4 * ref.abs().max()).sum()), 'min|pre| near argmax', float(pre_list[l][am].abs()))
    if float(e.max() / ref.abs().max()) > 1e-4:
        bad = (e > 1e-4 * ref.abs().max()).nonzero()
        print('  bad idx (first 10):', bad[:10].tolist())
        print('  mine', float(out.cpu()[am]), 'ref', float(ref[am]), 'pre', float(pre_list[l][am]), 'my act sign?')
