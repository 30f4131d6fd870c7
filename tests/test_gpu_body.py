"""GPU parity of the SMPL-X path (through the C ABI) against the oracle: forward, adjoint, gathers."""
import numpy as np
import pytest
import torch

from oracle import synth, ref_body as rb
from gpu_common import DEV, model_np, smplx_module, rand_pose, rel

pytestmark = pytest.mark.gpu
KEYS = ['transl', 'global_orient', 'betas', 'body_pose', 'left_hand_pose', 'right_hand_pose', 'expression', 'jaw_pose',
        'leye_pose', 'reye_pose']


def _oracle(nv, pose, dtype, grad=False):
    ref = rb.SMPLXRef(model_np(nv), dtype=dtype)
    t = {k: torch.from_numpy(pose[k]).to(dtype).requires_grad_(grad) for k in KEYS}
    v, j, fp = ref(**t)
    return v, j, fp, t


@pytest.fixture(params=['tc', 'tc_blend_simt_skin', 'simt'])
def blend_mode(request):
    """Back ends of the two LBS contractions: 'tc' = tcgen05 blend GEMM (TF32; Wt rounded once, X split hi/lo) + tcgen05 skinning
    GEMM (TF32 3-term split, full meshes only) -- the default; 'tc_blend_simt_skin' = the CUDA-core skinning kernel after the
    tensor-core blend; 'simt' = both on fp32 CUDA cores."""
    from lemo_b200 import _lib
    _lib.call('lemo_debug_set_blend_tc', 0 if request.param == 'simt' else 1)
    _lib.call('lemo_debug_set_skin_tc', 1 if request.param == 'tc' else 0)
    yield 'simt' if request.param == 'simt' else 'tc'
    _lib.call('lemo_debug_set_blend_tc', 1)
    _lib.call('lemo_debug_set_skin_tc', 1)


@pytest.mark.parametrize('nv,B', [(640, 5), (synth.V, 3), (synth.V, 119)])
def test_forward_matches_oracle(nv, B, blend_mode):
    pose = rand_pose(B, 7 + B)
    pose['global_orient'][0] = 0.0            # exercises the 1e-8 Rodrigues path
    out = smplx_module(nv)(return_verts=True, return_full_pose=True, **{k: torch.from_numpy(v).to(DEV) for k, v in pose.items()})
    v, j, fp, _ = _oracle(nv, pose, torch.float32)
    assert out.vertices.shape == (B, nv, 3) and out.joints.shape == (B, 127, 3)
    assert rel(out.vertices, v) < 1e-4, rel(out.vertices, v)          # north_star tolerance: 1e-4 relative fp32
    assert rel(out.joints, j) < 1e-4
    assert rel(out.full_pose, fp) < 1e-5
    v64, j64, _, _ = _oracle(nv, pose, torch.float64)
    if blend_mode == 'simt':
        # tolerance budget: as close to fp64 truth as the fp32 reference arithmetic is (x4 slack)
        assert rel(out.vertices, v64) < max(4 * rel(v, v64), 2e-6), (rel(out.vertices, v64), rel(v, v64))
    else:
        # TF32 tensor-core blend: the only rounding is Wt -> TF32 (2^-12 relative, once, unbiased): measured 2.3e-5 of max|v|
        assert rel(out.vertices, v64) < 5e-5, rel(out.vertices, v64)


@pytest.mark.parametrize('B', [1, 13, 300])
def test_skin_tc_matches_cuda_core_skinning(B):
    """tcgen05 skinning (3-term TF32 split, fp32-grade) against the CUDA-core kernel on the same v_posed: ragged frame chunks
    (B not a multiple of 12), the ragged last vertex tile (10475 = 81 x 128 + 107) and CTAs that change weight tile mid-range."""
    from lemo_b200 import _lib
    pose = {k: torch.from_numpy(v).to(DEV) for k, v in rand_pose(B, 50 + B).items()}
    mod = smplx_module(synth.V)
    out = {}
    try:
        for on in (1, 0):
            _lib.call('lemo_debug_set_skin_tc', on)
            o = mod(return_verts=True, **pose)
            out[on] = (o.vertices.clone(), o.joints.clone())
    finally:
        _lib.call('lemo_debug_set_skin_tc', 1)
    assert torch.isfinite(out[1][0]).all()
    assert rel(out[1][0], out[0][0]) < 2e-6, rel(out[1][0], out[0][0])
    assert rel(out[1][1], out[0][1]) < 2e-6


def test_golden_reference_lbs(golden):
    """Against outputs of the REAL reference lbs() (tests/golden), not just the restatement."""
    B = golden['lbs_small_pose'].shape[0]
    pose, betas = golden['lbs_small_pose'], golden['lbs_small_betas']
    from lemo_b200 import smplx as sx
    m = dict(model_np(640))
    m['hands_meanl'] = np.zeros(45, np.float32); m['hands_meanr'] = np.zeros(45, np.float32)
    mod = sx.create(m, use_pca=False, num_pca_comps=45, batch_size=B).to(DEV)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    out = mod(global_orient=T(pose[:, :3]), body_pose=T(pose[:, 3:66]), jaw_pose=T(pose[:, 66:69]), leye_pose=T(pose[:, 69:72]),
              reye_pose=T(pose[:, 72:75]), left_hand_pose=T(pose[:, 75:120]), right_hand_pose=T(pose[:, 120:165]),
              betas=T(betas[:, :10]), expression=T(betas[:, 10:]), transl=torch.zeros(B, 3, device=DEV))
    assert rel(out.vertices, golden['lbs_small_verts']) < 1e-4
    assert rel(out.joints[:, :55], golden['lbs_small_joints']) < 1e-4


@pytest.mark.parametrize('nv,B', [(640, 4), (synth.V, 2)])
def test_backward_matches_oracle_autograd(nv, B, blend_mode):
    pose = rand_pose(B, 21)
    g = np.random.default_rng(3)
    gv = g.standard_normal((B, nv, 3)).astype(np.float32)
    gj = g.standard_normal((B, 127, 3)).astype(np.float32)
    t = {k: torch.from_numpy(v).to(DEV).requires_grad_(True) for k, v in pose.items()}
    out = smplx_module(nv)(return_verts=True, **t)
    ((out.vertices * torch.from_numpy(gv).to(DEV)).sum() + (out.joints * torch.from_numpy(gj).to(DEV)).sum()).backward()
    res = {}
    for dtype in (torch.float32, torch.float64):
        v, j, _, tt = _oracle(nv, pose, dtype, grad=True)
        ((v * torch.from_numpy(gv).to(dtype)).sum() + (j * torch.from_numpy(gj).to(dtype)).sum()).backward()
        res[dtype] = {k: tt[k].grad for k in KEYS}
    for k in KEYS:
        e32 = rel(res[torch.float32][k], res[torch.float64][k])
        e = rel(t[k].grad, res[torch.float64][k])
        assert e < max(4 * e32, 2e-5 if blend_mode == 'simt' else 1e-4), (k, e, e32)   # tc: v_posed carries the TF32 rounding of Wt


def test_rotation_matrix_override_equals_aa_path():
    """R_global / R_body inputs (used by the fused fit) give the same mesh as the aa inputs they came from."""
    B, nv = 6, 640
    pose = rand_pose(B, 5)
    t = {k: torch.from_numpy(v).to(DEV) for k, v in pose.items()}
    mod = smplx_module(nv)
    a = mod(**t)
    Rg = rb.rodrigues(torch.from_numpy(pose['global_orient'])).reshape(B, 9).to(DEV)
    Rb = rb.rodrigues(torch.from_numpy(pose['body_pose']).reshape(-1, 3)).reshape(B, 21, 9).to(DEV)
    t2 = dict(t); t2.pop('global_orient'); t2.pop('body_pose')
    b = mod(R_global=Rg, R_body=Rb, return_full_pose=True, **t2)
    assert rel(b.vertices, a.vertices) < 2e-6
    assert rel(b.full_pose[:, :66], torch.from_numpy(np.concatenate([pose['global_orient'], pose['body_pose']], 1))) < 2e-5


def test_gather_rows_bit_exact():
    from lemo_b200 import _lib
    tab = synth.load_tables()
    B, V = 3, synth.V
    src = torch.randn(B, V, 3, device=DEV)
    for key in ('markers67', 'markers81', 'left_heel', 'right_toe'):
        idx = torch.from_numpy(tab[key]).to(DEV)
        out = torch.empty(B, idx.numel(), 3, device=DEV)
        _lib.call('lemo_gather_rows', _lib.ptr(src), _lib.ptr(idx), B, V, idx.numel(), _lib.ptr(out), _lib.cur_stream())
        assert torch.equal(out, src[:, idx.long()])            # integer indexing: bit-exact
    g = torch.randn(B, 81, 3, device=DEV)
    dst = torch.zeros(B, V, 3, device=DEV)
    idx = torch.from_numpy(tab['markers81']).to(DEV)
    _lib.call('lemo_scatter_rows_add', _lib.ptr(g), _lib.ptr(idx), B, V, 81, _lib.ptr(dst), _lib.cur_stream())
    ref = torch.zeros(B, V, 3, device=DEV); ref[:, idx.long()] = g
    assert torch.equal(dst, ref)


def test_errors_are_reported_not_thrown():
    from lemo_b200 import _lib
    with pytest.raises(RuntimeError, match='batch exceeds'):
        mod = smplx_module(640)
        body = mod._body(torch.device(DEV), 2)
        pose = rand_pose(4, 1)
        t = {k: torch.from_numpy(v).to(DEV) for k, v in pose.items()}
        from lemo_b200.smplx import _pose_struct
        import ctypes as C
        ps = _pose_struct({**{k: None for k in ['R_global', 'R_body']}, **t}, True, False)
        v = torch.empty(4, 640, 3, device=DEV)
        _lib.call('lemo_smplx_forward', body.handle, C.byref(ps), 4, _lib.ptr(v), None, None, _lib.cur_stream())
    with pytest.raises(RuntimeError, match='CUDA devices only'):
        import lemo_b200.smplx as sx
        sx.create(model_np(640), batch_size=1)(transl=torch.zeros(1, 3))


def test_sparse_skinning_adjoint_matches_dense_and_oracle():
    """Full mesh with SMPL-X-like sparse skinning weights (4 influences per vertex): the adjoint over the non-zeros (k_skin_bwd_sp_*) against
    the dense 55-wide kernel on the same model (summation order only) and against the oracle's autograd in float64."""
    import lemo_b200.smplx as smplx
    from lemo_b200 import _lib
    nv, B = synth.V, 5
    model = synth.make_smplx_model(0, n_verts=nv, weights_nnz=4)
    assert (model['lbs_weights'] != 0).sum(1).max() == 4
    body = smplx.create(model, model_type='smplx', gender='male', ext='npz', num_pca_comps=12, batch_size=B).to(DEV)
    pose = rand_pose(B, 31)
    g = np.random.default_rng(5)
    gv = torch.from_numpy(g.standard_normal((B, nv, 3)).astype(np.float32)).to(DEV)
    gj = torch.from_numpy(g.standard_normal((B, 127, 3)).astype(np.float32)).to(DEV)
    grads = {}
    for mode in (1, 0):
        _lib.call('lemo_debug_set_skin_sparse', mode)
        t = {k: torch.from_numpy(v).to(DEV).requires_grad_(True) for k, v in pose.items()}
        out = body(return_verts=True, **t)
        ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
        grads[mode] = {k: t[k].grad.clone() for k in KEYS}
    _lib.call('lemo_debug_set_skin_sparse', 1)
    ref = rb.SMPLXRef(model, dtype=torch.float64)
    tt = {k: torch.from_numpy(pose[k]).double().requires_grad_(True) for k in KEYS}
    v, j, _ = ref(**tt)
    ((v * gv.cpu().double()).sum() + (j * gj.cpu().double()).sum()).backward()
    for k in KEYS:
        assert rel(grads[1][k], grads[0][k]) < 2e-5, (k, rel(grads[1][k], grads[0][k]))
        assert rel(grads[1][k], tt[k].grad) < 1e-4, (k, rel(grads[1][k], tt[k].grad))
    # same inputs twice: bitwise identical (fixed summation order)
    t = {k: torch.from_numpy(v).to(DEV).requires_grad_(True) for k, v in pose.items()}
    out = body(return_verts=True, **t)
    ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
    for k in KEYS:
        assert torch.equal(t[k].grad, grads[1][k]), k
