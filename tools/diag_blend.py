"""GPU diagnostic: tcgen05 blend GEMM vs CUDA-core GEMM vs fp64 oracle; LBS forward timing in both modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from lemo_b200 import _lib
from oracle import synth, ref_body as rb
from gpu_common import smplx_module, model_np, rand_pose, rel
dev = 'cuda:0'
KEYS = ['transl', 'global_orient', 'betas', 'body_pose', 'left_hand_pose', 'right_hand_pose', 'expression', 'jaw_pose', 'leye_pose', 'reye_pose']
for nv, B in ((640, 5), (synth.V, 120), (synth.V, 300)):
    pose = rand_pose(B, 3)
    t = {k: torch.from_numpy(v).to(dev) for k, v in pose.items()}
    mod = smplx_module(nv)
    out = {}
    for mode in (0, 1):
        _lib.call('lemo_debug_set_blend_tc', mode)
        o = mod(return_verts=True, **t)
        torch.cuda.synchronize()
        out[mode] = o.vertices.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for _ in range(3): mod(return_verts=True, **t)
            torch.cuda.synchronize(); e0.record()
            for _ in range(10): mod(return_verts=True, **t)
            e1.record(); torch.cuda.synchronize()
        print('nv', nv, 'B', B, 'mode', 'tc' if mode else 'simt', 'smplx forward %.1f us' % (e0.elapsed_time(e1) * 100), flush=True)
    ref = rb.SMPLXRef(model_np(nv), dtype=torch.float64)
    v64, _, _ = ref(**{k: torch.from_numpy(pose[k]).double() for k in KEYS})
    print('   simt vs f64 %.2e   tc vs f64 %.2e   tc vs simt %.2e' % (rel(out[0], v64), rel(out[1], v64), rel(out[1], out[0])), flush=True)
