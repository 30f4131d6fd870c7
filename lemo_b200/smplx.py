"""Drop-in for the `smplx` package surface the reference fitting loops use.

    import lemo_b200.smplx as smplx
    body_model = smplx.create(model_path, model_type='smplx', gender='male', ext='npz', num_pca_comps=12,
                              create_global_orient=True, ..., batch_size=B).to('cuda')
    out = body_model(return_verts=True, **body_params_dict)        # out.vertices [B,10475,3], out.joints [B,127,3]

Reference call sites: opt_amass_perframe.py:66-80, opt_amass_temp.py:73-87, utils/utils.py:141-169,
temp_prox/main_slide.py:160-179, temp_prox/fitting_temp_slide.py:248-258.  Semantics follow smplx==0.1.26
`SMPLX.forward`.  All arithmetic runs in liblemo_b200.so (hand-written sm_100a
kernels) through the C ABI; torch only owns the tensors and the stream.
"""
import ctypes as C
import os
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib

ModelOutput = namedtuple('ModelOutput', ['vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose',
                                         'expression', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'transl'])
ModelOutput.__new__.__defaults__ = (None,) * len(ModelOutput._fields)

# smplx vertex_ids['smplx']: nose, eyes, ears, toes/heels, finger tips
EXTRA_JOINT_VIDS = np.array([9120, 9929, 9448, 616, 6, 5770, 5780, 8846, 8463, 8474, 8635,
                             5361, 4933, 5058, 5169, 5286, 8079, 7669, 7794, 7905, 8022], np.int32)


def _load_model_dict(model_path, model_type, gender, ext):
    if isinstance(model_path, dict):
        return model_path
    if os.path.isdir(model_path):
        model_path = os.path.join(model_path, model_type, 'SMPLX_%s.%s' % (gender.upper(), ext))
    if not os.path.exists(model_path):
        raise FileNotFoundError('SMPL-X model file not found: %s' % model_path)
    return dict(np.load(model_path, allow_pickle=True))


def _normalise(d, num_pca_comps, flat_hand_mean):
    """smplx .npz keys (or the synthetic dict of oracle/synth.py) -> contiguous fp32/int32 arrays for the C ABI."""
    f32 = lambda a: np.ascontiguousarray(np.asarray(a), np.float32)
    i32 = lambda a: np.ascontiguousarray(np.asarray(a), np.int32)
    V = np.asarray(d['v_template']).shape[0]
    out = {'v_template': f32(d['v_template'])}
    sd = np.asarray(d['shapedirs'])
    if sd.shape[-1] >= 310:     # SMPL-X v1.1 files: 300 shape + 100 expression components; expression dirs start at 300 (smplx body_models)
        sd = np.concatenate([sd[:, :, :10], sd[:, :, 300:310]], axis=-1)
    out['shapedirs'] = f32(sd[:, :, :20])
    pd = np.asarray(d['posedirs'])
    if pd.ndim == 3:                                   # [V,3,486] -> [486, 3V]  (body_model.py:126-128)
        pd = pd.reshape(-1, pd.shape[-1]).T
    out['posedirs'] = f32(pd)
    out['J_regressor'] = f32(d['J_regressor'].todense() if hasattr(d['J_regressor'], 'todense') else d['J_regressor'])
    out['lbs_weights'] = f32(d['lbs_weights'] if 'lbs_weights' in d else d['weights'])
    if 'parents' in d:
        parents = i32(d['parents']).copy()
    else:
        parents = i32(np.asarray(d['kintree_table'])[0]).copy()
    parents[0] = -1
    out['parents'] = parents
    out['hand_l'] = f32(np.asarray(d['hands_componentsl'])[:num_pca_comps])
    out['hand_r'] = f32(np.asarray(d['hands_componentsr'])[:num_pca_comps])
    pm = np.zeros(165, np.float32)
    if not flat_hand_mean:
        pm[75:120] = np.asarray(d['hands_meanl'], np.float32)
        pm[120:165] = np.asarray(d['hands_meanr'], np.float32)
    out['pose_mean'] = pm
    out['faces'] = i32(d['faces'] if 'faces' in d else d['f'])
    out['extra_joint_vids'] = i32(d['extra_joint_vids']) if 'extra_joint_vids' in d else np.minimum(EXTRA_JOINT_VIDS, V - 1).astype(np.int32)
    out['lmk_faces_idx'] = i32(d['lmk_faces_idx'])
    out['lmk_bary'] = f32(d['lmk_bary_coords'])
    return out


class DeviceModel:
    """Owns a LemoModel handle (immutable model tensors on one CUDA device)."""

    def __init__(self, arrays, device_index):
        self.arrays = arrays
        self.device_index = device_index
        a = arrays
        desc = _lib.LemoModelDescC()
        desc.n_verts = a['v_template'].shape[0]
        desc.n_faces = a['faces'].shape[0]
        desc.num_pca_comps = a['hand_l'].shape[0]
        desc.n_extra_joints = a['extra_joint_vids'].shape[0]
        desc.n_landmarks = a['lmk_faces_idx'].shape[0]
        for field, key in (('h_v_template', 'v_template'), ('h_shapedirs', 'shapedirs'), ('h_posedirs', 'posedirs'),
                           ('h_J_regressor', 'J_regressor'), ('h_lbs_weights', 'lbs_weights'), ('h_parents', 'parents'),
                           ('h_hand_comp_l', 'hand_l'), ('h_hand_comp_r', 'hand_r'), ('h_pose_mean', 'pose_mean'),
                           ('h_extra_joint_vids', 'extra_joint_vids'), ('h_faces', 'faces'),
                           ('h_lmk_faces_idx', 'lmk_faces_idx'), ('h_lmk_bary', 'lmk_bary')):
            setattr(desc, field, a[key].ctypes.data)
        h = C.c_void_p()
        _lib.call('lemo_model_create', C.byref(desc), device_index, C.byref(h))
        self.handle = h
        self.n_verts = desc.n_verts

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().lemo_model_destroy(self.handle)
        except Exception:
            pass


class _Body:
    """LemoBody handle: forward/backward scratch for a fixed maximum batch."""

    def __init__(self, dmodel, max_batch, with_backward=True):
        self.dmodel = dmodel
        self.max_batch = max_batch
        h = C.c_void_p()
        _lib.call('lemo_body_create', dmodel.handle, max_batch, 1 if with_backward else 0, C.byref(h))
        self.handle = h
        self.stamp = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().lemo_body_destroy(self.handle)
        except Exception:
            pass


_ARG_ORDER = ['transl', 'global_orient', 'body_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'left_hand_pose',
              'right_hand_pose', 'betas', 'expression', 'R_global', 'R_body']


def _pose_struct(tensors, use_pca, betas_shared):
    p = _lib.LemoPoseC()
    for k in _ARG_ORDER:
        t = tensors.get(k)
        setattr(p, k, t.data_ptr() if t is not None else None)
    p.betas_shared = 1 if betas_shared else 0
    p.use_pca = 1 if use_pca else 0
    return p


class _SMPLXFunction(torch.autograd.Function):
    """verts, joints, full_pose = SMPL-X(pose parts).  Forward/backward = lemo_smplx_forward/backward."""

    @staticmethod
    def forward(ctx, body, use_pca, want_joints, *args):
        tensors = {k: (a.contiguous().float() if a is not None else None) for k, a in zip(_ARG_ORDER, args)}
        ref = next(t for t in tensors.values() if t is not None)
        B = ref.shape[0]
        dev = ref.device
        betas_shared = tensors['betas'] is not None and tensors['betas'].shape[0] == 1 and B > 1
        V = body.dmodel.n_verts
        verts = torch.empty(B, V, 3, device=dev, dtype=torch.float32)
        joints = torch.empty(B, 127, 3, device=dev, dtype=torch.float32) if want_joints else None
        full_pose = torch.empty(B, 165, device=dev, dtype=torch.float32)
        pose = _pose_struct(tensors, use_pca, betas_shared)
        _lib.call('lemo_smplx_forward', body.handle, C.byref(pose), B, _lib.ptr(verts), _lib.ptr(joints),
                  _lib.ptr(full_pose), _lib.cur_stream(dev))
        body.stamp += 1
        ctx.body, ctx.use_pca, ctx.B, ctx.stamp, ctx.betas_shared = body, use_pca, B, body.stamp, betas_shared
        ctx.tensors = tensors
        ctx.want_joints = want_joints
        ctx.mark_non_differentiable(full_pose)
        if joints is None:
            joints = torch.empty(0, device=dev)
        return verts, joints, full_pose

    @staticmethod
    def backward(ctx, g_verts, g_joints, _g_fp):
        body, B, tensors = ctx.body, ctx.B, ctx.tensors
        dev = next(t for t in tensors.values() if t is not None).device
        pose = _pose_struct(tensors, ctx.use_pca, ctx.betas_shared)
        st = _lib.cur_stream(dev)
        if body.stamp != ctx.stamp:
            # another forward ran on this handle since ours (the reference calls the model twice per iteration,
            # opt_amass_temp.py:357,364): rebuild the saved state for OUR inputs before taking the adjoint
            V = body.dmodel.n_verts
            scratch = torch.empty(B, V, 3, device=dev, dtype=torch.float32)
            _lib.call('lemo_smplx_forward', body.handle, C.byref(pose), B, _lib.ptr(scratch), None, None, st)
            body.stamp += 1
            ctx.stamp = body.stamp
        grads = _lib.LemoPoseGradC()
        outs = []
        for i, k in enumerate(_ARG_ORDER):
            t = tensors[k]
            if t is not None and ctx.needs_input_grad[3 + i]:
                g = torch.zeros_like(t)
                setattr(grads, k, g.data_ptr())
                outs.append(g)
            else:
                outs.append(None)
        gv = g_verts.contiguous().float() if g_verts is not None else None
        gj = g_joints.contiguous().float() if (g_joints is not None and ctx.want_joints) else None
        _lib.call('lemo_smplx_backward', body.handle, C.byref(pose), B, _lib.ptr(gv), _lib.ptr(gj), C.byref(grads), st)
        return (None, None, None) + tuple(outs)


class SMPLX(nn.Module):
    NUM_BODY_JOINTS = 21
    NUM_JOINTS = 55

    def __init__(self, arrays, batch_size=1, num_pca_comps=12, use_pca=True, joint_mapper=None, dtype=torch.float32,
                 gender='neutral', create_body_pose=True):
        super().__init__()
        self._arrays = arrays
        self.batch_size = batch_size
        self.num_pca_comps = num_pca_comps
        self.use_pca = use_pca
        self.joint_mapper = joint_mapper
        self.gender = gender
        self.dtype = dtype
        hand_dim = num_pca_comps if use_pca else 45
        z = lambda n: nn.Parameter(torch.zeros(batch_size, n, dtype=dtype), requires_grad=True)
        self.betas = z(10)
        self.global_orient = z(3)
        if create_body_pose:
            self.body_pose = z(63)
        else:       # smplx create_body_pose=False (temp_prox/main_slide.py: `create_body_pose=not use_vposer`): not a parameter
            self.register_buffer('body_pose', torch.zeros(batch_size, 63, dtype=dtype))
        self.left_hand_pose = z(hand_dim)
        self.right_hand_pose = z(hand_dim)
        self.jaw_pose = z(3)
        self.leye_pose = z(3)
        self.reye_pose = z(3)
        self.expression = z(10)
        self.transl = z(3)
        self.faces = arrays['faces']
        self.register_buffer('faces_tensor', torch.from_numpy(arrays['faces'].astype(np.int64)))
        self._dmodels = {}
        self._bodies = {}

    def get_num_verts(self):
        return self._arrays['v_template'].shape[0]

    @torch.no_grad()
    def reset_params(self, **params_dict):
        """smplx semantics: named parameters take the given value, everything else is zero-filled."""
        for name, p in self.named_parameters():
            if name in params_dict:
                p[:] = torch.as_tensor(np.asarray(params_dict[name]) if not torch.is_tensor(params_dict[name])
                                       else params_dict[name], dtype=p.dtype, device=p.device).reshape(p.shape)
            else:
                p.fill_(0)

    def device_model(self, device):
        if device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback); move the module with .to("cuda")')
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._dmodels:
            with torch.cuda.device(idx):
                self._dmodels[idx] = DeviceModel(self._arrays, idx)
        return self._dmodels[idx]

    def _body(self, device, B):
        dm = self.device_model(device)
        key = (dm.device_index, B)
        if key not in self._bodies:
            with torch.cuda.device(dm.device_index):
                self._bodies[key] = _Body(dm, B)
        return self._bodies[key]

    def forward(self, betas=None, global_orient=None, body_pose=None, left_hand_pose=None, right_hand_pose=None,
                transl=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None, return_verts=True,
                return_full_pose=False, R_global=None, R_body=None, **kwargs):
        pick = lambda v, p: p if v is None else v
        vals = dict(transl=pick(transl, self.transl), global_orient=pick(global_orient, self.global_orient),
                    body_pose=pick(body_pose, self.body_pose), jaw_pose=pick(jaw_pose, self.jaw_pose),
                    leye_pose=pick(leye_pose, self.leye_pose), reye_pose=pick(reye_pose, self.reye_pose),
                    left_hand_pose=pick(left_hand_pose, self.left_hand_pose),
                    right_hand_pose=pick(right_hand_pose, self.right_hand_pose), betas=pick(betas, self.betas),
                    expression=pick(expression, self.expression), R_global=R_global, R_body=R_body)
        if R_global is not None:
            vals['global_orient'] = None
        if R_body is not None:
            vals['body_pose'] = None
        B = max(v.shape[0] for v in vals.values() if v is not None)
        for k, v in vals.items():
            if v is not None and v.shape[0] != B and not (k == 'betas' and v.shape[0] == 1):
                raise AssertionError('%s has batch %d, expected %d (smplx modules are built for a fixed batch_size)' % (k, v.shape[0], B))
        dev = vals['transl'].device
        body = self._body(dev, B)
        verts, joints, full_pose = _SMPLXFunction.apply(body, self.use_pca, True, *[vals[k] for k in _ARG_ORDER])
        if self.joint_mapper is not None:
            joints = self.joint_mapper(joints)
        return ModelOutput(vertices=verts if return_verts else None, joints=joints,
                           full_pose=full_pose if return_full_pose else None, betas=vals['betas'],
                           global_orient=vals['global_orient'], body_pose=vals['body_pose'], expression=vals['expression'],
                           left_hand_pose=vals['left_hand_pose'], right_hand_pose=vals['right_hand_pose'],
                           jaw_pose=vals['jaw_pose'], transl=vals['transl'])


def create(model_path, model_type='smplx', gender='neutral', ext='npz', num_pca_comps=12, use_pca=True,
           flat_hand_mean=False, batch_size=1, joint_mapper=None, dtype=torch.float32, **kwargs):
    """smplx.create(...) for model_type='smplx'.  `model_path` may also be a dict of model arrays (synthetic models).
    create_body_pose=False keeps body_pose out of .parameters() (the PROX script passes `create_body_pose=not use_vposer`); the other
    create_* flags are accepted and ignored: those parameters always exist, as in the reference's calls."""
    if model_type != 'smplx':
        raise ValueError('only model_type="smplx" is on the LEMO fitting path')
    arrays = _normalise(_load_model_dict(model_path, model_type, gender, ext), num_pca_comps if use_pca else 45, flat_hand_mean)
    return SMPLX(arrays, batch_size=batch_size, num_pca_comps=num_pca_comps, use_pca=use_pca, joint_mapper=joint_mapper,
                 dtype=dtype, gender=gender, create_body_pose=bool(kwargs.get('create_body_pose', True)))
