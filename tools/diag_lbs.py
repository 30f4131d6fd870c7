"""GPU diagnostic: full-mesh lemo_smplx_forward at B=120 -- eager and CUDA-graph-replayed time per back end, parity between them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lemo_b200 import _lib, synth
import lemo_b200.smplx as smplx


def smplx_module(n_verts):
    return smplx.create(synth.make_smplx_model(0, n_verts=n_verts), model_type='smplx', gender='male', ext='npz', num_pca_comps=12,
                        batch_size=1).to('cuda:0')


def rand_pose(B, seed, scale=0.3):
    g = np.random.default_rng(seed)
    f = lambda *s: (scale * g.standard_normal(s)).astype(np.float32)
    return dict(transl=f(B, 3), global_orient=f(B, 3), body_pose=f(B, 63), jaw_pose=f(B, 3), leye_pose=f(B, 3), reye_pose=f(B, 3),
                left_hand_pose=f(B, 12), right_hand_pose=f(B, 12), betas=g.standard_normal((B, 10)).astype(np.float32), expression=f(B, 10))


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
dev = 'cuda:0'
Bs = [int(x) for x in sys.argv[1:]] or [120]
for B in Bs:
    pose = {k: torch.from_numpy(v).to(dev) for k, v in rand_pose(B, 3).items()}
    mod = smplx_module(synth.V)
    outs = {}
    for name, bt, sk in (('tc blend + tc skin', 1, 1), ('tc blend + simt skin', 1, 0), ('simt blend + simt skin', 0, 0)):
        _lib.call('lemo_debug_set_blend_tc', bt)
        _lib.call('lemo_debug_set_skin_tc', sk)
        with torch.no_grad():
            o = mod(return_verts=True, **pose)
            torch.cuda.synchronize()
            outs[name] = o.vertices.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3): mod(return_verts=True, **pose)
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): mod(return_verts=True, **pose)
            e1.record(); torch.cuda.synchronize()
            eager = e0.elapsed_time(e1) * 50
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                mod(return_verts=True, **pose)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    og = mod(return_verts=True, **pose)
            for _ in range(3): g.replay()
            torch.cuda.synchronize(); e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            graph = e0.elapsed_time(e1) * 50
        alg = 4.0 * (512 * 3 * synth.V + synth.V * 55 + 3 * synth.V + B * 168) + 4.0 * B * 3 * synth.V
        print('B %d  %-24s eager %.1f us   graph %.1f us  (%.0f GB/s algorithmic)   graph-vs-eager output %.1e' %
              (B, name, eager, graph, alg / graph / 1e3, rel(og.vertices, outs[name])), flush=True)
    ks = list(outs)
    print('   tc-skin vs simt-skin %.2e   tc vs all-simt %.2e' % (rel(outs[ks[0]], outs[ks[1]]), rel(outs[ks[0]], outs[ks[2]])), flush=True)
_lib.call('lemo_debug_set_blend_tc', 1); _lib.call('lemo_debug_set_skin_tc', 1)
