"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol include/lemo_b200.h
declares, and its host-callable rotation math (the same __host__ __device__ code the kernels run) matches the oracle."""
import ctypes as C
import numpy as np
import torch

from lemo_b200 import _lib
from oracle import ref_body as rb


def test_exports_every_declared_symbol(built_lib):
    declared = _lib.header_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(built_lib, s)]
    assert not missing, missing
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert built_lib.lemo_version() == 100


def _host(fn, *arrays, out_shape):
    out = np.zeros(out_shape, np.float32)
    getattr(_lib.lib(), fn)(*[a.ctypes.data_as(C.c_void_p) for a in arrays], out.ctypes.data_as(C.c_void_p))
    return out


def test_host_rodrigues_and_adjoint(built_lib):
    g = np.random.default_rng(0)
    for scale in (1.0, 1e-3, 0.0):
        for _ in range(8):
            aa = (scale * g.standard_normal(3)).astype(np.float32)
            dR = g.standard_normal(9).astype(np.float32)
            t = torch.from_numpy(aa).double().requires_grad_(True)
            R = rb.rodrigues(t[None])[0]
            (R.reshape(9) * torch.from_numpy(dR).double()).sum().backward()
            assert np.abs(_host('lemo_host_rodrigues', aa, out_shape=9) - R.detach().numpy().reshape(9)).max() < 2e-6
            got = _host('lemo_host_rodrigues_bwd', aa, dR, out_shape=3)
            if scale > 0:
                assert np.abs(got - t.grad.numpy()).max() < 2e-4 * max(1.0, np.abs(t.grad.numpy()).max())


def test_host_gs6d_and_adjoint(built_lib):
    g = np.random.default_rng(1)
    for _ in range(16):
        x = g.standard_normal(6).astype(np.float32)
        dR = g.standard_normal(9).astype(np.float32)
        t = torch.from_numpy(x).double().requires_grad_(True)
        R = rb.gram_schmidt_6d(t[None])[0]
        (R.reshape(9) * torch.from_numpy(dR).double()).sum().backward()
        assert np.abs(_host('lemo_host_gs6d', x, out_shape=9) - R.detach().numpy().reshape(9)).max() < 2e-6
        assert np.abs(_host('lemo_host_gs6d_bwd', x, dR, out_shape=6) - t.grad.numpy()).max() < 1e-4 * max(1.0, np.abs(t.grad.numpy()).max())


def test_host_tgm_conversions(built_lib):
    g = np.random.default_rng(2)
    aa = np.concatenate([g.standard_normal((32, 3)), 3.0 * g.standard_normal((32, 3)), 1e-4 * g.standard_normal((4, 3))]).astype(np.float32)
    for a in aa:
        R = rb.tgm_aa_to_rotmat(torch.from_numpy(a)[None])[0]
        assert np.abs(_host('lemo_host_aa_to_rotmat_tgm', a, out_shape=9) - R.numpy().reshape(9)).max() < 2e-6
        Rr = rb.rodrigues(torch.from_numpy(a)[None])[0].numpy().reshape(9).astype(np.float32)
        want = rb.rotmat_to_aa(torch.from_numpy(Rr).view(1, 3, 3))[0].numpy()
        assert np.abs(_host('lemo_host_rotmat_to_aa', Rr, out_shape=3) - want).max() < 5e-6 * max(1.0, np.abs(want).max())


def test_host_rotmat_to_aa_adjoint_all_branches(built_lib):
    """lemo_host_rotmat_to_aa_bwd (the same HD code as the lemo_rotmat_to_aa_backward kernel) vs autograd through the oracle's restated
    torchgeometry conversion, on rotations that hit all four quaternion branches (angles up to pi)."""
    g = np.random.default_rng(3)
    seen = set()
    for i in range(400):
        aa = ((1.0, 3.0, 0.01, 3.1)[i % 4] * g.standard_normal(3)).astype(np.float32)
        if i % 4 == 3:
            aa = (aa / np.linalg.norm(aa) * 3.14).astype(np.float32)
        R = rb.rodrigues(torch.from_numpy(aa)[None])[0].numpy().astype(np.float32).reshape(9)
        t = R.reshape(3, 3).T
        seen.add((0 if t[0, 0] > t[1, 1] else 1) if t[2, 2] < 1e-6 else (2 if t[0, 0] < -t[1, 1] else 3))
        Rt = torch.from_numpy(R).double().requires_grad_(True)
        d = g.standard_normal(3).astype(np.float32)
        (rb.rotmat_to_aa(Rt.view(1, 3, 3))[0] * torch.from_numpy(d).double()).sum().backward()
        want = Rt.grad.numpy()
        got = _host('lemo_host_rotmat_to_aa_bwd', np.ascontiguousarray(R), d, out_shape=9)
        assert np.abs(got - want).max() < 1e-5 * max(1.0, np.abs(want).max())
    assert seen == {0, 1, 2, 3}


def test_tgm_known_answers_from_shipped_results(built_lib):
    """Known-answer test for the sign / branch convention of the restated torchgeometry conversions.  The reference ships ten result
    clips (res_opt_amass_{perframe,temp}/TotalCapture/body_params_opt_clip_*.npy, fixtures in tests/golden/seed_clips.npz) whose
    global-orient axis-angles were produced by the REAL torchgeometry 0.1.2 `rotation_matrix_to_angle_axis` (utils/utils.py:80) with
    angles up to pi, i.e. on the hard quaternion branches.  They must be fixed points of rotmat_to_aa(rodrigues(aa)) for the oracle's
    restatement and for the library's HD code."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'seed_clips.npz'))
    n, hard = 0, 0
    for k in z.files:
        if '_params_' not in k:
            continue
        aa = z[k][:, 3:6].astype(np.float32)
        ang = np.linalg.norm(aa, axis=1)
        hard += int((ang > 3.0).sum())
        R = rb.rodrigues(torch.from_numpy(aa).double())
        back = rb.rotmat_to_aa(R).numpy()
        assert np.abs(back - aa).max() < 5e-6, (k, np.abs(back - aa).max())
        R32 = R.float().numpy().reshape(-1, 9)
        for i in range(0, aa.shape[0], 7):
            got = _host('lemo_host_rotmat_to_aa', np.ascontiguousarray(R32[i]), out_shape=3)
            # fp32 R -> aa loses ~1e-7/sin(theta/2) near pi: 2e-4 absolute covers the angle-3.1416 frames, branch/sign errors are O(1)
            assert np.abs(got - aa[i]).max() < 2e-4, (k, i, got, aa[i])
        n += aa.shape[0]
    assert n == 10 * 119 and hard > 0


def test_no_cpu_fallback_in_product():
    """The product package must never import the oracle (parity claims depend on it)."""
    import os, re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'lemo_b200')
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith('.py'):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), os.path.join(dp, f)


def test_bad_arguments_return_status_and_message(built_lib):
    """Error contract of the boundary: a non-zero int status plus a thread-local message via lemo_last_error(),
    never an exception or a crash across the ABI -- checked on entry points of every subsystem with null / empty arguments (the
    argument checks run before any CUDA call, so this needs no GPU)."""
    L = built_lib
    for name in ('lemo_smplx_forward', 'lemo_smplx_backward', 'lemo_gather_rows', 'lemo_scatter_rows_add', 'lemo_vposer_decode',
                 'lemo_vposer_decode_backward', 'lemo_enc_forward', 'lemo_enc_backward_input', 'lemo_ae_forward', 'lemo_chamfer_forward',
                 'lemo_chamfer_backward', 'lemo_camera_project', 'lemo_sdf_sample', 'lemo_adam_step', 'lemo_fit_run', 'lemo_fit_run_perframe',
                 'lemo_fit_set_sequence', 'lemo_fit_get'):
        f = getattr(L, name)
        args = [None if t in (C.c_void_p,) or 'LP_' in t.__name__ or t.__name__.endswith('_p') else 0 for t in f.argtypes]
        status = f(*args)                                  # empty problem (B = n = 0) on null buffers
        assert status != 0, name
        msg = L.lemo_last_error().decode()
        assert msg and '@' in msg, (name, msg)             # "<what> (<failed condition>) @file:line"


def test_sparse_synthetic_model_has_smplx_like_skinning_weights():
    """synth.make_smplx_model(weights_nnz=4): at most 4 influences per vertex, on a joint and its ancestors, rows normalised, and a tile of
    256 consecutive vertices touches a handful of joints (what the compact skinning adjoint keys on); the default model stays dense."""
    import numpy as np
    from lemo_b200 import synth
    m = synth.make_smplx_model(0, weights_nnz=4)
    w = m['lbs_weights']
    nz = w != 0
    assert nz.sum(1).max() == 4 and nz.sum(1).min() >= 1
    assert np.allclose(w.sum(1), 1.0, atol=1e-6)
    par = m['parents']
    for v in (0, 1234, 5000, w.shape[0] - 1):
        js = np.nonzero(nz[v])[0]
        top = js.max()                                    # the home joint; the others are its ancestors
        chain, c = {int(top)}, int(top)
        while par[c] >= 0 and len(chain) < 4:
            c = int(par[c]); chain.add(c)
        assert set(js.tolist()) == chain
    ntile = (w.shape[0] + 255) // 256
    slots = sum(int(nz[t * 256:(t + 1) * 256].any(0).sum()) for t in range(ntile))
    assert slots * 5 < ntile * 55 * 2
    assert (synth.make_smplx_model(0)['lbs_weights'] != 0).all()


def _tree_tables_py(parents):
    """Independent construction of the tables of lemo_host_tree_tables (layout documented in include/lemo_b200.h)."""
    import numpy as np
    J = 55
    par = np.array(parents, np.int64).copy()
    par[0] = -1
    depth = np.zeros(J, np.int64)
    for j in range(1, J):
        depth[j] = depth[par[j]] + 1
    t = np.zeros(776, np.int64)
    t[0:J] = par
    order = sorted(range(J), key=lambda j: (depth[j], j))
    t[56:56 + J] = order
    md = int(depth.max())
    for lev in range(md + 2):
        t[112 + lev] = sum(1 for j in range(J) if depth[j] < lev)
    kids = [[c for c in range(J) if par[c] == j] for j in range(J)]
    k = 0
    for j in range(J):
        t[128 + j] = k
        for c in kids[j]:
            t[184 + k] = c
            k += 1
    t[128 + J] = k
    t[240:720] = -1
    if md > 14 or np.bincount(depth).max() > 32:
        return None, md
    for lev in range(md + 1):
        lv = [j for j in order if depth[j] == lev]
        for i, j in enumerate(lv):
            t[240 + lev * 32 + i] = j | ((par[j] + 1) << 8) | (t[128 + j] << 16) | (len(kids[j]) << 24)
    t[720:720 + J] = depth
    return t, md


def test_host_tree_tables_match_independent_construction():
    """The level-ordered kinematic-tree tables the chain kernels walk (host code, exported for this test): the real SMPL-X tree and random
    trees against an independent Python construction; structural properties; the error codes."""
    import ctypes as C
    import numpy as np
    from lemo_b200 import _lib, synth
    L = _lib.lib()
    g = np.random.default_rng(0)
    cases = [synth.PARENTS.astype(np.int32)]
    for _ in range(20):
        p = np.zeros(55, np.int32)
        for j in range(1, 55):
            p[j] = g.integers(0, max(1, (2 * j) // 3))          # bushy enough to stay within 14 levels, narrow enough for 32 per level (mostly)
        cases.append(p)
    checked = 0
    for p in cases:
        out = np.zeros(776, np.int32)
        md = C.c_int32(0)
        rc = L.lemo_host_tree_tables(p.ctypes.data, out.ctypes.data, 776, C.byref(md))
        ref, md_ref = _tree_tables_py(p)
        if ref is None:
            assert rc in (12, 13)
            continue
        assert rc == 0 and md.value == md_ref
        assert np.array_equal(out.astype(np.int64), ref), np.nonzero(out != ref)[0][:10]
        # every joint appears exactly once among the packed words, on its own level, after its parent's level
        words = out[240:720].reshape(15, 32)
        seen = {}
        for lev in range(md_ref + 1):
            for w in words[lev]:
                if w >= 0:
                    seen[int(w) & 255] = lev
        assert sorted(seen) == list(range(55))
        assert all(seen[j] == seen[int(p[j])] + 1 for j in range(1, 55))
        checked += 1
    assert checked >= 10
    # the real tree: 11 levels, at most 10 joints on one (the finger levels)
    out = np.zeros(776, np.int32); md = C.c_int32(0)
    assert L.lemo_host_tree_tables(cases[0].ctypes.data, out.ctypes.data, 776, C.byref(md)) == 0 and md.value == 10
    assert np.bincount(out[720:775]).max() == 10
    bad = cases[0].copy(); bad[5] = 7
    assert L.lemo_host_tree_tables(bad.ctypes.data, out.ctypes.data, 776, C.byref(md)) == 11
    chain = np.arange(-1, 54, dtype=np.int32)                    # a 55-deep chain: deeper than the level table
    assert L.lemo_host_tree_tables(chain.ctypes.data, out.ctypes.data, 776, C.byref(md)) == 12
    star = np.zeros(55, np.int32)                                # 54 children of the root: a level wider than a warp
    assert L.lemo_host_tree_tables(star.ctypes.data, out.ctypes.data, 776, C.byref(md)) == 13
    assert L.lemo_host_tree_tables(cases[0].ctypes.data, out.ctypes.data, 100, C.byref(md)) == 1
