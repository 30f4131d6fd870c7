"""GPU experiment: does splitting the S sequences of a GPU into concurrent groups (separate fit handles / streams) hide the
latency-bound body-model kernels of one group behind the tensor-core conv stack of the other?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lemo_b200.smplx as smplx
from lemo_b200.vposer import VPoserDecoder
from lemo_b200.fit import TemporalFitter, load_smooth_prior
from lemo_b200.utils.utils import gen_body_mesh_v1
from oracle import synth
dev = torch.device('cuda:0')
T, S = 120, 8
body = smplx.create(synth.make_smplx_model(0), batch_size=T).to(dev)
vp = VPoserDecoder(synth.make_vposer_weights(1)).to(dev)
enc = load_smooth_prior().to(dev)
m67 = torch.from_numpy(synth.load_tables()['markers67']).long().to(dev)
seqs = []
for s in range(S):
    clean, init, contact = synth.make_sequence(s, T=T)
    with torch.no_grad():
        v = gen_body_mesh_v1(torch.from_numpy(clean).to(dev), body, vp)
    seqs.append((init, v[:, m67].cpu().numpy(), contact))
for groups in (1, 2, 4):
    per = S // groups
    fits = []
    for g in range(groups):
        f = TemporalFitter(body, vp, per, T, enc=load_smooth_prior().to(dev) if g else enc, device=dev)
        for i in range(per):
            f.set_sequence(i, *seqs[g * per + i])
        fits.append(f)
    for f in fits: f.run(n_iters=5)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in fits: f.run(n_iters=30)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('groups %d x %d seq: %.3f ms/step  -> %.0f seq-it/s' % (groups, per, dt / 30 * 1e3, S * 30 / dt), flush=True)
    del fits
