#!/usr/bin/env python
"""Run ONE pass of a secondary stage (for ncu launch lists): python tools/run_stage.py [infill] [perframe] [prox]."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lemo_b200 import _lib, synth                      # noqa: E402
_lib.build()
import lemo_b200.smplx as smplx                          # noqa: E402
from lemo_b200.vposer import VPoserDecoder               # noqa: E402

dev = torch.device('cuda', 0)
stages = sys.argv[1:] or ['infill']
model = synth.make_smplx_model(0)
body = smplx.create(model, model_type='smplx', gender='male', ext='npz', num_pca_comps=12, batch_size=120).to(dev)
vp = VPoserDecoder(synth.make_vposer_weights(1)).to(dev)

if 'infill' in stages:
    from lemo_b200.infill import InfillStage, body_repr, load_infill_prior, load_infill_stats
    body68, con68 = synth.synth_marker_clip(5, T=120)
    st64 = load_infill_stats()
    stage = InfillStage(load_infill_prior(), device=dev, stats=st64)
    clip, rot0 = body_repr(torch.from_numpy(body68).to(dev), torch.from_numpy(con68).to(dev), stats=st64, device=dev)
    torch.cuda.synchronize()
    print('INFILL BEGIN', flush=True)
    stage.run(clip, rot0)
    torch.cuda.synchronize()
    print('INFILL END', flush=True)

if 'perframe' in stages:
    from lemo_b200.fit import PerFrameFitter
    S, T = 8, 2
    pf = PerFrameFitter(body, vp, S, T, device=dev, use_cuda_graph=False)
    clean, _, _ = synth.make_sequence(0, T=T)
    for i in range(S):
        pf.set_sequence(i, clean[0, 6:16], np.zeros((T, 67, 3), np.float32))
    pf.run(n_iters=10)
    torch.cuda.synchronize()
    print('PERFRAME done', flush=True)

if 'prox' in stages:
    from lemo_b200.temp_prox.synthetic import make_window
    from lemo_b200.fit import load_smooth_prior
    fit, _, _ = make_window(body, vp, load_smooth_prior().to(dev), B=100, D=256, m_scene=100000, device=dev, use_cuda_graph=False)
    torch.cuda.synchronize()
    print('PROX BEGIN', flush=True)
    fit.run(3)
    torch.cuda.synchronize()
    print('PROX done', flush=True)
