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
MODES = ((0, 'simt fp32'), (2, 'tc per-tap'), (3, 'tc row-reuse'), (4, 'tc + streamed weights, 3 stages'), (5, 'tc + stacked [Whi;Wlo] N=128'),
         (6, 'tc + 8 epilogue warps'), (1, 'tc pair, lean epilogue, 8 warps (default)'), (7, 'tc pair, lean epilogue, 16 warps'))
# timing experiments on the pair kernel (mode = 8 + bit mask; results are garbage by construction): 1 no TMA loads, 2 no MMAs, 4 no epilogue stores,
# 8 skip the N=64 MMA, 16 no kx descriptor offset, 32 one N=256 MMA instead of the pair, 64 skip the N=128 MMA
EXPERIMENTS = ((8 + 1, 'no loads'), (8 + 2, 'no MMA'), (8 + 4, 'no stores'), (8 + 5, 'MMA only (no loads, no stores)'), (8 + 6, 'loads only'),
               (8 + 3, 'stores only'), (8 + 7, 'nothing (pipeline skeleton)'))
# 8 skip the N=64 MMA, 16 no kx descriptor offset, 32 one N=256 MMA instead of the pair, 64 skip the N=128 MMA, 128 / 256 issue every step 2x / 4x,
# 512 rotate the destination over 4 accumulator blocks (independent accumulate chains instead of one dependent chain)
TENSOR = ((8 + 5 + 256, 'MMA only 4x: N=128 + N=64, one chain'), (8 + 13 + 256, 'MMA only 4x: N=128 only, one chain'), (8 + 69 + 256, 'MMA only 4x: N=64 only, one chain'),
          (8 + 37 + 256, 'MMA only 4x: N=256, one chain'),
          (8 + 5 + 256 + 512, 'MMA only 4x: N=128 + N=64, 4 chains'), (8 + 13 + 256 + 512, 'MMA only 4x: N=128 only, 4 chains'),
          (8 + 69 + 256 + 512, 'MMA only 4x: N=64 only, 4 chains'), (8 + 37 + 256 + 512, 'MMA only 4x: N=256, 2 chains'), (8 + 21 + 256 + 512, 'MMA only 4x: pair, 4 chains, no kx offset'))
if '--tensor' in sys.argv: MODES, EXPERIMENTS = (), TENSOR
for mode, name in (MODES + EXPERIMENTS if ('--experiments' in sys.argv or '--tensor' in sys.argv) else MODES):
    _lib.call('lemo_debug_set_conv_tc', mode)
    z = enc(x)[0]
    torch.cuda.synchronize()
    err = float((z[:1, ::8, ::7, ::9].cpu() - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, 8, 0, 3, _lib.cur_stream())
    torch.cuda.synchronize(); e0.record()
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, 8, 0, 20, _lib.cur_stream())
    e1.record(); torch.cuda.synchronize()
    print('%-40s z rel err %.2e   64->64 layer %.1f us' % (name, err, e0.elapsed_time(e1) * 50), flush=True)
