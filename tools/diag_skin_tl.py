"""GPU diagnostic: run one full-mesh forward with LEMO_SKIN_TL=1 so k_skin_tc<TL> prints the timeline of CTA 0."""
import os, sys
os.environ['LEMO_SKIN_TL'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from oracle import synth
from gpu_common import smplx_module, rand_pose
B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
pose = {k: torch.from_numpy(v).to('cuda:0') for k, v in rand_pose(B, 3).items()}
mod = smplx_module(synth.V)
with torch.no_grad():
    for _ in range(3):
        mod(return_verts=True, **pose)
        torch.cuda.synchronize()
