"""GPU experiment: tensor-core conv variants (per-tap TMA boxes vs row-reuse with/without descriptor base_offset)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lemo_b200 import _lib
from lemo_b200.fit import load_smooth_prior
dev = 'cuda:0'
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_golden.npz')))
enc = load_smooth_prior().to(dev)
x = torch.from_numpy(g['enc_full_x']).to(dev).repeat(8, 1, 1, 1).contiguous()
ref = torch.from_numpy(g['enc_full_z_sub'])
net = enc.net(torch.device(dev), 8, 245, 134)
for mode, name in ((0, 'simt fp32'), (2, 'tc per-tap'), (3, 'tc row-reuse'), (4, 'tc + streamed weights, 3 stages'), (5, 'tc + stacked [Whi;Wlo] N=128'), (1, 'tc + 8 epilogue warps (default)')):
    _lib.call('lemo_debug_set_conv_tc', mode)
    z = enc(x)[0]
    torch.cuda.synchronize()
    err = float((z[:1, ::8, ::7, ::9].cpu() - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, 8, 0, 3, _lib.cur_stream())
    torch.cuda.synchronize(); e0.record()
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, 8, 0, 20, _lib.cur_stream())
    e1.record(); torch.cuda.synchronize()
    print('%-28s z rel err %.2e   64->64 layer %.1f us' % (name, err, e0.elapsed_time(e1) * 50), flush=True)
