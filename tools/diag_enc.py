"""GPU diagnostic: where does Enc's input gradient differ from the fp64 oracle?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, ref_priors as rp
from lemo_b200.fit import load_smooth_prior

torch.set_printoptions(linewidth=200, precision=3, sci_mode=True)
dev = 'cuda:0'
enc = load_smooth_prior().to(dev)
sd64 = {k: torch.from_numpy(v).double() for k, v in synth.load_enc_weights().items()}
for (N, H, W) in ((1, 21, 30), (2, 37, 53)):
    x = torch.from_numpy((0.5 * np.random.default_rng(9).standard_normal((N, 1, H, W))).astype(np.float32))
    gz = torch.from_numpy(np.random.default_rng(10).standard_normal((N, 64, H, W)).astype(np.float32))
    xx = x.double().requires_grad_(True)
    z64 = rp.enc_forward(xx, sd64)
    (z64 * gz.double()).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    z = enc(xg)[0]
    (z * gz.to(dev)).sum().backward()
    ez = (z.detach().cpu().double() - z64.detach()).abs()
    print('fwd rel err', float(ez.max() / z64.abs().max()), 'argmax', np.unravel_index(int(ez.argmax()), ez.shape))
    e = (xg.grad.cpu().double() - xx.grad).abs()[:, 0]
    print('bwd rel err', float(e.max() / xx.grad.abs().max()))
    print('err by column (max over rows):', (e.max(1).values[0] / xx.grad.abs().max()).numpy().round(6))
    print('err by row (max over cols):', (e.max(2).values[0] / xx.grad.abs().max()).numpy().round(6))
# single-channel probes: gradient only through one output channel / one pixel
N, H, W = 1, 21, 30
x = torch.from_numpy((0.5 * np.random.default_rng(9).standard_normal((N, 1, H, W))).astype(np.float32))
for (c, y, xq) in ((0, 10, 15), (5, 0, 0), (7, 20, 29), (63, 10, 0)):
    gz = torch.zeros(N, 64, H, W); gz[0, c, y, xq] = 1.0
    xx = x.double().requires_grad_(True)
    (rp.enc_forward(xx, sd64) * gz.double()).sum().backward()
    xg = x.to(dev).requires_grad_(True)
    (enc(xg)[0] * gz.to(dev)).sum().backward()
    e = (xg.grad.cpu().double() - xx.grad).abs()[0, 0]
    print('probe', (c, y, xq), 'rel err', float(e.max() / xx.grad.abs().max()), 'at', np.unravel_index(int(e.argmax()), e.shape))
