"""GPU diagnostic (next round's first call): VPoser MLP + adjoint on the three GEMM back ends -- accuracy vs an fp64 oracle and time.
The mode is read once per process (LEMO_VPOSER), so run it once per mode:
    for m in simt tc tc64; do LEMO_VPOSER=$m python tools/diag_vposer_modes.py; done
Decision rule for making tc64 the default: R_body and dz errors within 2x of the simt row, and faster at B = 960."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from oracle import ref_body as rb
from gpu_common import DEV, vposer_w, rel

mode = os.environ.get('LEMO_VPOSER', 'simt')
from lemo_b200.vposer import VPoserDecoder
vp = VPoserDecoder(vposer_w()).to(DEV)
for B in (7, 120, 960):
    g = np.random.default_rng(B)
    z = torch.from_numpy(g.standard_normal((B, 32)).astype(np.float32))
    gR = torch.from_numpy(g.standard_normal((B * 21, 3, 3)).astype(np.float32))
    ref = rb.VPoserRef(vposer_w(), dtype=torch.float64)
    z64 = z.double().requires_grad_(True)
    R64 = ref.decode_matrot(z64)
    (R64 * gR.double()).sum().backward()
    zg = z.to(DEV).requires_grad_(True)
    R = vp.decode(zg, 'matrot')
    (R.view(B * 21, 3, 3) * gR.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gRd = gR.to(DEV)
    for _ in range(3):
        zz = z.to(DEV).requires_grad_(True); (vp.decode(zz, 'matrot').view(B * 21, 3, 3) * gRd).sum().backward()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20):
        zz = z.to(DEV).requires_grad_(True); (vp.decode(zz, 'matrot').view(B * 21, 3, 3) * gRd).sum().backward()
    e1.record(); torch.cuda.synchronize()
    print('LEMO_VPOSER=%-5s B %4d  R_body err %.2e  dz err %.2e  fwd+bwd %.1f us (eager, incl. torch glue)' %
          (mode, B, rel(R.view(B * 21, 3, 3), R64), rel(zg.grad, z64.grad), e0.elapsed_time(e1) * 50), flush=True)
