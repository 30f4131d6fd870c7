"""GPU experiment: the weights-in-TMEM conv kernel (k_conv_tc_wt, mode 8192 + bits) against the pair kernel (mode 1) and fp32 SIMT (0).

Parity: Enc forward vs the REAL reference's golden z, input gradient vs the golden gradient (median, kink-robust) -- for every mode.
Timing: one 64->64 layer (forward and input gradient), S = 8, CUDA events, plus the ablations of the wt kernel.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lemo_b200 import _lib
from lemo_b200.fit import load_smooth_prior

dev = 'cuda:0'
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = dict(np.load(os.path.join(root, 'tests', 'golden', 'reference_golden.npz')))
enc = load_smooth_prior().to(dev)
S = 8
net = enc.net(torch.device(dev), S, 245, 134)


def q_rel(a, b, q):
    a, b = a.double().cpu().numpy(), np.asarray(b, np.float64)
    return float(np.quantile(np.abs(a - b), q) / np.abs(b).max())


def parity(mode):
    _lib.call('lemo_debug_set_conv_tc', mode)
    out = []
    for tag in ('small', 'full'):
        x = torch.from_numpy(g['enc_%s_x' % tag]).to(dev).requires_grad_(True)
        z = enc(x)[0]
        loss = (z[..., 1:] - z[..., :-1]).pow(2).mean()
        loss.backward()
        torch.cuda.synchronize()
        zr = z if tag == 'small' else z[:, ::8, ::7, ::9]
        ref = g['enc_small_z'] if tag == 'small' else g['enc_full_z_sub']
        out.append('%s: z %.2e loss %.2e gx med %.2e max %.2e' % (
            tag, q_rel(zr.detach(), ref, 1.0), abs(float(loss) - float(g['enc_%s_loss' % tag])) / float(g['enc_%s_loss' % tag]),
            q_rel(x.grad, g['enc_%s_gx' % tag], 0.5), q_rel(x.grad, g['enc_%s_gx' % tag], 1.0)))
    return ' | '.join(out)


def timing(mode, backward):
    _lib.call('lemo_debug_set_conv_tc', mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, S, backward, 3, _lib.cur_stream())
    torch.cuda.synchronize(); e0.record()
    _lib.call('lemo_convnet_profile_layer', net.handle, 5, S, backward, 20, _lib.cur_stream())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 50          # us per launch


x8 = torch.from_numpy(g['enc_full_x']).to(dev).repeat(S, 1, 1, 1).contiguous()
WT = 8192
_lib.call('lemo_debug_set_conv_tc', WT)
enc(x8)                                       # sizes the S = 8 buffers
torch.cuda.synchronize()
only = [int(a) for a in sys.argv[1:] if a.isdigit()]
if '--timeline' in sys.argv:
    for bits, name in ((0, 'full kernel'), (10, 'MMA only'), (14 + 16, 'skeleton, no epilogue work')):
        timing(WT, 0)
        print('---- timeline of CTA 0, %s (globaltimer ns since kernel entry)' % name, flush=True)
        _lib.call('lemo_debug_set_conv_tc', WT + 128 + bits)
        _lib.call('lemo_convnet_profile_layer', net.handle, 5, S, 0, 1, _lib.cur_stream())
        torch.cuda.synchronize()
if '--experiments' in sys.argv:
    for mode, name in ((1, 'tc pair'), (WT, 'tc weights-in-TMEM, 2 epilogue groups'), (WT + 64, 'tc weights-in-TMEM, 1 epilogue group')):
        print('%-44s 64->64 layer: forward %.1f us, input gradient %.1f us' % (name, timing(mode, 0), timing(mode, 1)), flush=True)
    # bit 1 (2) no epilogue stores, bit 2 (4) no MMAs, bit 3 (8) no TMA loads, bit 4 (16) epilogue = hand-shakes only, bit 5 (32) TMEM load but no arithmetic
    for g in (0, 64):
        for bits, name in ((2, 'no epilogue stores'), (4, 'no MMAs'), (8, 'no TMA loads'), (6, 'loads only'), (10, 'MMA only'), (12, 'epilogue only'),
                           (14, 'pipeline skeleton'), (16, 'no epilogue work'), (32, 'epilogue = TMEM load only'), (10 + 16, 'MMA only, no epilogue work'),
                           (10 + 32, 'MMA only, epilogue = TMEM load only'), (14 + 16, 'skeleton, no epilogue work'), (14 + 32, 'skeleton, TMEM load only')):
            print('wt[%d grp] ablation %-36s forward %.1f us' % (1 if g else 2, name, timing(WT + g + bits, 0)), flush=True)
for mode, name in ((0, 'simt fp32'), (1, 'tc pair'), (WT, 'tc weights-in-TMEM'), (WT + 64, 'tc weights-in-TMEM, 1 group')):
    if mode in only:
        print('%-44s %s' % (name, parity(mode)), flush=True)
_lib.call('lemo_debug_set_conv_tc', -1)
