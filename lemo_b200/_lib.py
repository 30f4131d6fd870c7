"""ctypes binding of liblemo_b200.so (the C ABI declared in include/lemo_b200.h).

There is no CPU fallback: every op in this package goes through this library, and loading fails loudly
if the library has not been built (`python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.path.join(_HERE, '_build', 'liblemo_b200.so')
HEADER = os.path.join(_HERE, '..', 'include', 'lemo_b200.h')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-ldl']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into lemo_b200/_build/liblemo_b200.so (in-tree, so it travels)."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get('LEMO_NVCC_EXTRA', '').split() + (['-Xptxas', '-v'] if verbose else []) + ['-o', LIB_PATH] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB_PATH


class LemoModelDescC(C.Structure):
    _fields_ = [('n_verts', C.c_int32), ('n_faces', C.c_int32), ('num_pca_comps', C.c_int32),
                ('n_extra_joints', C.c_int32), ('n_landmarks', C.c_int32),
                ('h_v_template', C.c_void_p), ('h_shapedirs', C.c_void_p), ('h_posedirs', C.c_void_p),
                ('h_J_regressor', C.c_void_p), ('h_lbs_weights', C.c_void_p), ('h_parents', C.c_void_p),
                ('h_hand_comp_l', C.c_void_p), ('h_hand_comp_r', C.c_void_p), ('h_pose_mean', C.c_void_p),
                ('h_extra_joint_vids', C.c_void_p), ('h_faces', C.c_void_p), ('h_lmk_faces_idx', C.c_void_p),
                ('h_lmk_bary', C.c_void_p)]


POSE_FIELDS = ['transl', 'global_orient', 'body_pose', 'jaw_pose', 'leye_pose', 'reye_pose', 'left_hand_pose',
               'right_hand_pose', 'betas', 'expression', 'R_global', 'R_body']


class LemoPoseC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in POSE_FIELDS] + [('betas_shared', C.c_int32), ('use_pca', C.c_int32)]


class LemoPoseGradC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in POSE_FIELDS]


class LemoFitConfigC(C.Structure):
    _fields_ = [('mode', C.c_int32), ('n_seq', C.c_int32), ('n_frames', C.c_int32),
                ('w_rec', C.c_float), ('w_vposer', C.c_float), ('w_shape', C.c_float), ('w_hand', C.c_float),
                ('w_contact', C.c_float), ('w_smooth', C.c_float), ('vel_thres', C.c_float), ('fps', C.c_float),
                ('h_markers67', C.c_void_p), ('h_markers81', C.c_void_p), ('h_foot_ids', C.c_void_p * 4),
                ('n_foot', C.c_int32 * 4), ('h_smooth_mean', C.c_void_p), ('h_smooth_std', C.c_void_p),
                ('use_cuda_graph', C.c_int32)]


class LemoProxWeightsC(C.Structure):
    _fields_ = [(k, C.c_float) for k in ('data_weight', 'body_pose_weight', 'shape_weight', 'bending_prior_weight', 'hand_prior_weight',
                                         'expr_prior_weight', 'jaw_prior_weight', 'sdf_penetration_weight', 'contact_loss_weight',
                                         'motion_prior_smooth_weight', 'friction_normal_weight', 'friction_tangent_weight')] + \
               [('use_joints_conf', C.c_int32)]


class LemoProxConfigC(C.Structure):
    _fields_ = [('n_frames', C.c_int32), ('n_joints_mapped', C.c_int32), ('h_joint_map', C.c_void_p),
                ('cam_R', C.c_float * 9), ('cam_t', C.c_float * 3), ('fx', C.c_float), ('fy', C.c_float), ('cx', C.c_float), ('cy', C.c_float),
                ('R', C.c_float * 9), ('t', C.c_float * 3),
                ('sdf', C.c_void_p), ('sdf_dim', C.c_int32), ('grid_min', C.c_float * 3), ('grid_max', C.c_float * 3),
                ('sdf_penetration', C.c_int32), ('use_friction', C.c_int32), ('contact', C.c_int32), ('use_motion_smooth_prior', C.c_int32),
                ('h_fric_ids', C.c_void_p), ('n_fric', C.c_int32), ('h_contact_ids', C.c_void_p), ('n_contact', C.c_int32),
                ('h_markers81', C.c_void_p), ('scene_v', C.c_void_p), ('n_scene', C.c_int32),
                ('h_smooth_mean', C.c_void_p), ('h_smooth_std', C.c_void_p), ('weights', LemoProxWeightsC), ('use_cuda_graph', C.c_int32)]


PROX_PARAMS = ['transl', 'global_orient', 'pose_embedding', 'left_hand_pose', 'right_hand_pose', 'jaw_pose', 'leye_pose', 'reye_pose',
               'expression']


class LemoProxWindowC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PROX_PARAMS + ['betas', 'gt_joints', 'joints_conf', 'joint_weights']]


class LemoProxParamsOutC(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PROX_PARAMS]


_P, _I, _L, _F, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); restype int means "status code"
SIGNATURES = {
    'lemo_last_error': (C.c_char_p, []),
    'lemo_version': (C.c_int, []),
    'lemo_debug_check': (C.c_int, []),
    'lemo_model_create': (C.c_int, [C.POINTER(LemoModelDescC), C.c_int, C.POINTER(_P)]),
    'lemo_model_select_rows': (C.c_int, [_P, _P, _I, C.POINTER(_P)]),
    'lemo_model_destroy': (C.c_int, [_P]),
    'lemo_model_num_verts': (C.c_int, [_P]),
    'lemo_body_create': (C.c_int, [_P, _I, _I, C.POINTER(_P)]),
    'lemo_body_destroy': (C.c_int, [_P]),
    'lemo_smplx_forward': (C.c_int, [_P, C.POINTER(LemoPoseC), _I, _P, _P, _P, _P]),
    'lemo_smplx_backward': (C.c_int, [_P, C.POINTER(LemoPoseC), _I, _P, _P, C.POINTER(LemoPoseGradC), _P]),
    'lemo_debug_set_blend_tc': (C.c_int, [_I]),
    'lemo_debug_set_skin_tc': (C.c_int, [_I]),
    'lemo_debug_set_skin_sparse': (C.c_int, [_I]),
    'lemo_host_tree_tables': (C.c_int, [_P, _P, _I, _P]),
    'lemo_gather_rows': (C.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    'lemo_scatter_rows_add': (C.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    'lemo_rot6d_to_rotmat': (C.c_int, [_P, _I, _P, _P]),
    'lemo_rot6d_to_rotmat_backward': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_rotmat_to_aa': (C.c_int, [_P, _I, _P, _P]),
    'lemo_rotmat_to_aa_backward': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_aa_to_rot6d': (C.c_int, [_P, _I, _P, _P]),
    'lemo_rodrigues': (C.c_int, [_P, _I, _P, _P]),
    'lemo_rodrigues_backward': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_vposer_create': (C.c_int, [_P, _P, _P, _P, _P, _P, _I, C.c_int, C.POINTER(_P)]),
    'lemo_vposer_destroy': (C.c_int, [_P]),
    'lemo_vposer_decode': (C.c_int, [_P, _P, _I, _P, _P, _P]),
    'lemo_vposer_decode_backward': (C.c_int, [_P, _P, _I, _P, _P, _P]),
    'lemo_convnet_create': (C.c_int, [_I, _I, _P, _L, _I, _I, _I, _I, C.c_int, C.POINTER(_P)]),
    'lemo_convnet_destroy': (C.c_int, [_P]),
    'lemo_convnet_set_weights': (C.c_int, [_P, _P, _P]),
    'lemo_convnet_get_weights': (C.c_int, [_P, _P, _P]),
    'lemo_convnet_num_weights': (_L, [_P]),
    'lemo_enc_forward': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_enc_backward_input': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_debug_set_conv_tc': (C.c_int, [_I]),
    'lemo_enc_debug_backward': (C.c_int, [_P, _P, _I, _I, _P, _P]),
    'lemo_convnet_profile_layer': (C.c_int, [_P, _I, _I, _I, _I, _P]),
    'lemo_ae_forward': (C.c_int, [_P, _P, _I, _P, _P, _P]),
    'lemo_ae_backward_weights': (C.c_int, [_P, _P, _I, _P, _P]),
    'lemo_ae_finetune_step': (C.c_int, [_P, _P, _P, _I, _I, _D, _I, _P, _P]),
    'lemo_ae_finetune_run': (C.c_int, [_P, _P, _P, _I, _I, _D, _I, _P, _P]),
    'lemo_chamfer_forward': (C.c_int, [_P, _I, _I, _P, _I, _L, _P, _P, _P, _P, _P]),
    'lemo_chamfer_backward': (C.c_int, [_P, _I, _I, _P, _I, _L, _P, _P, _P, _P, _P, _P, _P]),
    'lemo_scene_create': (C.c_int, [_P, _I, C.POINTER(_P)]),
    'lemo_scene_destroy': (C.c_int, [_P]),
    'lemo_scene_query': (C.c_int, [_P, _P, _I, _I, _P, _P, _P]),
    'lemo_camera_project': (C.c_int, [_P, _L, _P, _P, _F, _F, _F, _F, _P, _P]),
    'lemo_camera_project_backward': (C.c_int, [_P, _L, _P, _P, _F, _F, _F, _F, _P, _P, _P]),
    'lemo_rigid_transform': (C.c_int, [_P, _L, _P, _P, _I, _P, _P]),
    'lemo_sdf_sample': (C.c_int, [_P, _L, _P, _I, _P, _P, _P, _P]),
    'lemo_sdf_sample_backward': (C.c_int, [_P, _L, _P, _I, _P, _P, _P, _P, _P]),
    'lemo_adam_step': (C.c_int, [_P, _P, _P, _P, _L, _D, _D, _D, _D, _I, _P]),
    'lemo_fit_create': (C.c_int, [_P, _P, _P, _P, C.POINTER(LemoFitConfigC), C.c_int, C.POINTER(_P)]),
    'lemo_fit_destroy': (C.c_int, [_P]),
    'lemo_repr_local_markers_4chan': (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P]),
    'lemo_reconstruct_global_body': (C.c_int, [_P, _P, _I, _P, _P, _P]),
    'lemo_infill_prepare_input': (C.c_int, [_P, _I, _I, _P, _P, _P, _P]),
    'lemo_infill_finalize': (C.c_int, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    'lemo_fit_set_sequence': (C.c_int, [_P, _I, _P, _P, _P, _P]),
    'lemo_fit_set_sequences': (C.c_int, [_P, _P, _P, _P, _P]),
    'lemo_fit_run': (C.c_int, [_P, _I, _F, _F, _I, _P]),
    'lemo_fit_run_perframe': (C.c_int, [_P, _I, _P]),
    'lemo_fit_get': (C.c_int, [_P, _P, _P, _P]),
    'lemo_fit_get_state': (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    'lemo_fit_kernel_launches': (_L, [_P]),
    'lemo_debug_set_perframe': (C.c_int, [_I]),
    'lemo_fit_prox_create': (C.c_int, [_P, _P, _P, C.POINTER(LemoProxConfigC), C.c_int, C.POINTER(_P)]),
    'lemo_fit_prox_destroy': (C.c_int, [_P]),
    'lemo_fit_prox_set_weights': (C.c_int, [_P, C.POINTER(LemoProxWeightsC), _I, _P]),
    'lemo_fit_prox_set_window': (C.c_int, [_P, C.POINTER(LemoProxWindowC), _P]),
    'lemo_fit_prox_run': (C.c_int, [_P, _I, _F, _I, _P]),
    'lemo_fit_prox_eval': (C.c_int, [_P, _P]),
    'lemo_fit_prox_get': (C.c_int, [_P, C.POINTER(LemoProxParamsOutC), C.POINTER(LemoProxParamsOutC), _P, _P]),
    'lemo_fit_prox_kernel_launches': (_L, [_P]),
    'lemo_host_rodrigues': (None, [_P, _P]),
    'lemo_host_rodrigues_bwd': (None, [_P, _P, _P]),
    'lemo_host_gs6d': (None, [_P, _P]),
    'lemo_host_gs6d_bwd': (None, [_P, _P, _P]),
    'lemo_host_rotmat_to_aa': (None, [_P, _P]),
    'lemo_host_rotmat_to_aa_bwd': (None, [_P, _P, _P]),
    'lemo_host_aa_to_rotmat_tgm': (None, [_P, _P]),
}
STATUS_FUNCS = {k for k, (r, _) in SIGNATURES.items() if r is C.c_int and k not in ('lemo_version', 'lemo_model_num_verts', 'lemo_debug_check')}

_lib = None
_DEBUG_CHECK = os.environ.get('LEMO_DEBUG_CHECK', '0') == '1'     # synchronise + check the CUDA error state after every C-ABI call


def lib():
    """The loaded library (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('lemo_b200: %s is missing -- run __graft_entry__.build() (nvcc, sm_100a). '
                               'There is no CPU fallback.' % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)            # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def call(name, *args):
    """Call a status-returning entry point; raise RuntimeError(lemo_last_error()) on failure."""
    L = lib()
    r = getattr(L, name)(*args)
    if name in STATUS_FUNCS and r != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, r, L.lemo_last_error().decode()))
    if _DEBUG_CHECK and L.lemo_debug_check() != 0:
        raise RuntimeError('after %s: %s' % (name, L.lemo_last_error().decode()))
    return r


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None, as c_void_p."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def cur_stream(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def header_symbols():
    """Every function name declared in include/lemo_b200.h (used by the CPU-side ABI test)."""
    import re
    txt = open(HEADER).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(lemo_[a-z0-9_]+)\s*\(', txt)))
