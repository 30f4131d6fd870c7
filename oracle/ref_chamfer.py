"""ctypes front end of oracle/csrc/chamfer_ref.c (exact brute-force Chamfer NN, pinned arithmetic).  TEST INFRASTRUCTURE ONLY.

The torch restatement oracle/ref_priors.py:chamfer materialises the [B,n,m] distance tensor and rounds the sum of squares without
FMA; this C restatement fixes the evaluation order bit for bit (see the header of the C file) and scales to BASELINE config 4
(100 x 1121 x 100 000) in seconds on the host cores.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, '_build', 'liboracle.so')
_lib = None


def build():
    src = os.path.join(_HERE, 'csrc', 'chamfer_ref.c')
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, '-B', '_build/liboracle.so'], check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.chamfer_ref_nn.restype = None
        L.chamfer_ref_nn.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def chamfer_nn(q, t):
    """q [B,n,3], t [B,m,3] or [m,3] / [1,m,3] (shared) float32 -> (dist [B,n] float32, idx [B,n] int32)."""
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    B, n = q.shape[:2]
    shared = t.ndim == 2 or t.shape[0] == 1 and B > 1
    m = t.shape[-2]
    dist = np.empty((B, n), np.float32)
    idx = np.empty((B, n), np.int32)
    L = lib()
    t_bs = 0 if shared else m * 3

    def one(b):          # ctypes releases the GIL: one batch element per host thread
        L.chamfer_ref_nn(q[b].ctypes.data, n * 3, n, t.ctypes.data + 4 * b * t_bs, t_bs, m, 1, dist[b].ctypes.data, idx[b].ctypes.data)
    if B * n * m < 1 << 22:
        for b in range(B):
            one(b)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(B, os.cpu_count() or 1)) as ex:
            list(ex.map(one, range(B)))
    return dist, idx


def chamfer(xyz1, xyz2):
    """Both directions, as chamferDist.forward returns them: dist1 [B,n], dist2 [B,m], idx1, idx2 (numpy)."""
    xyz1 = np.ascontiguousarray(xyz1, np.float32)
    xyz2 = np.ascontiguousarray(xyz2, np.float32)
    B = xyz1.shape[0]
    d1, i1 = chamfer_nn(xyz1, xyz2)
    x2 = np.broadcast_to(xyz2 if xyz2.ndim == 3 else xyz2[None], (B,) + xyz2.shape[-2:])
    d2, i2 = chamfer_nn(np.ascontiguousarray(x2), xyz1)
    return d1, d2, i1, i2
