"""ProxFitter: Python handle of the fused PROX stage-2 driver (lemo_fit_prox_* in include/lemo_b200.h, csrc/fit_prox.cu).

One instance = one B-frame sliding window on one GPU.  It is what `FittingMonitor.run_fitting` dispatches to when the optimiser is
Adam and the loss is an `SMPLifyLoss` whose active terms the driver covers (fitting_temp_slide.py here); it can also be used
directly (bench.py, tests).  Nothing here computes: every call is one C-ABI call that enqueues kernels on the current stream.
"""
import ctypes as C
import numpy as np
import torch

from .. import _lib

WEIGHT_KEYS = ['data_weight', 'body_pose_weight', 'shape_weight', 'bending_prior_weight', 'hand_prior_weight', 'expr_prior_weight',
               'jaw_prior_weight', 'sdf_penetration_weight', 'contact_loss_weight', 'motion_prior_smooth_weight', 'friction_normal_weight',
               'friction_tangent_weight']
LOSS_KEYS = ['joint_loss', 'pprior_loss', 'shape_loss', 'angle_prior_loss', 'hand_prior_loss', 'expression_loss', 'jaw_prior_loss',
             'sdf_penetration_loss', 'loss_fric_tangent', 'loss_fric_normal', 'contact_loss', 'motion_prior_smooth_loss', '_r0', '_r1', '_r2',
             'total_loss']
PARAM_DIMS = dict(transl=3, global_orient=3, pose_embedding=32, left_hand_pose=12, right_hand_pose=12, jaw_pose=3, leye_pose=3,
                  reye_pose=3, expression=10)


def _f32(a, n=None):
    a = np.ascontiguousarray(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, np.float32))
    return a.reshape(n) if n is not None else a


class ProxFitter:
    def __init__(self, body_model, vposer, enc, n_frames, device, joint_map=None, camera=None, cam2world=None, sdf=None, grid_min=None,
                 grid_max=None, fric_ids=None, contact_ids=None, markers81=None, scene_v=None, smooth_stats=None, weights=None,
                 use_joints_conf=True, sdf_penetration=True, use_friction=True, contact=True, use_motion_smooth_prior=True,
                 use_cuda_graph=True):
        """camera = (R[3,3], t[3], fx, fy, (cx, cy)); cam2world = (R[3,3], t[3]); sdf: device tensor [D,D,D] (kept alive here);
        scene_v: device tensor [m,3]; smooth_stats = (Xmean[243], Xstd[243]); weights: dict of WEIGHT_KEYS."""
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('lemo_b200 runs on CUDA devices only (no CPU fallback)')
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device('cuda', idx)
        self.B = int(n_frames)
        B = self.B
        cfg = _lib.LemoProxConfigC()
        self._keep = []
        i32 = lambda a: np.ascontiguousarray(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a), np.int32)
        cfg.n_frames = B
        if joint_map is not None:
            jm = i32(joint_map)
            self._keep.append(jm)
            cfg.h_joint_map, cfg.n_joints_mapped = jm.ctypes.data, jm.shape[0]
            self.Jm = jm.shape[0]
        else:
            cfg.h_joint_map, cfg.n_joints_mapped = None, 127
            self.Jm = 127
        Rc, tc, fx, fy, cc = camera if camera is not None else (np.eye(3), np.zeros(3), 5000.0, 5000.0, (0.0, 0.0))
        cfg.cam_R[:] = list(_f32(Rc, 9)); cfg.cam_t[:] = list(_f32(tc, 3))
        cfg.fx, cfg.fy, cfg.cx, cfg.cy = float(fx), float(fy), float(cc[0]), float(cc[1])
        Rw, tw = cam2world if cam2world is not None else (np.eye(3), np.zeros(3))
        cfg.R[:] = list(_f32(Rw, 9)); cfg.t[:] = list(_f32(tw, 3))
        self._sdf = None
        if sdf is not None:
            from .camera import one_volume
            self._sdf = one_volume(sdf).to(self.device, torch.float32).contiguous()      # ONE volume (the reference repeats it B x)
            cfg.sdf, cfg.sdf_dim = self._sdf.data_ptr(), self._sdf.shape[-1]
            cfg.grid_min[:] = list(_f32(grid_min).reshape(-1)[:3]); cfg.grid_max[:] = list(_f32(grid_max).reshape(-1)[:3])
        cfg.sdf_penetration, cfg.use_friction = int(bool(sdf_penetration)), int(bool(use_friction))
        cfg.contact, cfg.use_motion_smooth_prior = int(bool(contact)), int(bool(use_motion_smooth_prior and enc is not None))
        if fric_ids is not None:
            a = i32(fric_ids); self._keep.append(a)
            cfg.h_fric_ids, cfg.n_fric = a.ctypes.data, a.shape[0]
        if contact_ids is not None:
            a = i32(contact_ids); self._keep.append(a)
            cfg.h_contact_ids, cfg.n_contact = a.ctypes.data, a.shape[0]
        if markers81 is None:
            from ..fit import load_tables
            markers81 = load_tables()['markers81']
        m81 = i32(markers81); self._keep.append(m81)
        cfg.h_markers81 = m81.ctypes.data
        self._scene = None
        if scene_v is not None:
            self._scene = scene_v.reshape(-1, 3).to(self.device, torch.float32).contiguous()
            cfg.scene_v, cfg.n_scene = self._scene.data_ptr(), self._scene.shape[0]
        if smooth_stats is None:
            from ..fit import load_tables
            t = load_tables()
            smooth_stats = (t['smooth_Xmean'], t['smooth_Xstd'])
        mean, std = _f32(smooth_stats[0], 243), _f32(smooth_stats[1], 243)
        self._keep += [mean, std]
        cfg.h_smooth_mean, cfg.h_smooth_std = mean.ctypes.data, std.ctypes.data
        self._weights = self._weights_struct(weights or {}, use_joints_conf)
        cfg.weights = self._weights
        cfg.use_cuda_graph = 1 if use_cuda_graph else 0
        with torch.cuda.device(idx):
            self._dmodel = body_model.device_model(self.device)
            self._vp = vposer.handle(self.device, B, private=True)
            self._enc = enc.net(self.device, 1, 245, B - 1 + 16, private=True) if cfg.use_motion_smooth_prior else None
            h = C.c_void_p()
            _lib.call('lemo_fit_prox_create', self._dmodel.handle, self._vp.handle, self._enc.handle if self._enc else None, C.byref(cfg), idx,
                      C.byref(h))
        self.handle = h
        self.iters_run = 0

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.lib().lemo_fit_prox_destroy(self.handle)
        except Exception:
            pass

    @staticmethod
    def _weights_struct(w, use_joints_conf=True):
        s = _lib.LemoProxWeightsC()
        for k in WEIGHT_KEYS:
            v = w.get(k, 0.0)
            setattr(s, k, float(v.item() if torch.is_tensor(v) else v))
        s.use_joints_conf = int(bool(use_joints_conf))
        return s

    def set_weights(self, weights, erase_n=0, use_joints_conf=True):
        """loss.reset_loss_weights(curr_weights) + the closure's `grad[0:erase_n] = 0` (fitting_temp_slide.py:281-288, :548-562)."""
        self._weights = self._weights_struct(weights, use_joints_conf)
        _lib.call('lemo_fit_prox_set_weights', self.handle, C.byref(self._weights), int(erase_n), _lib.cur_stream(self.device))

    def _dev(self, a, shape):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).detach().to(self.device, torch.float32).contiguous()
        assert tuple(t.shape) == tuple(shape), 'expected shape %s, got %s' % (tuple(shape), tuple(t.shape))
        return t

    def set_window(self, params, gt_joints, joints_conf, joint_weights):
        """params: dict with transl, global_orient, pose_embedding, left/right_hand_pose, jaw/leye/reye_pose, expression, betas ([B, .])."""
        B = self.B
        w = _lib.LemoProxWindowC()
        keep = []
        for k, n in list(PARAM_DIMS.items()) + [('betas', 10)]:
            if params.get(k) is not None:
                t = self._dev(params[k], (B, n)); keep.append(t)
                setattr(w, k, t.data_ptr())
        g = self._dev(gt_joints, (B, self.Jm, 2)); keep.append(g)
        w.gt_joints = g.data_ptr()
        if joints_conf is not None:
            c = self._dev(joints_conf, (B, self.Jm)); keep.append(c)
            w.joints_conf = c.data_ptr()
        jw = self._dev(joint_weights, (B, self.Jm)); keep.append(jw)
        w.joint_weights = jw.data_ptr()
        _lib.call('lemo_fit_prox_set_window', self.handle, C.byref(w), _lib.cur_stream(self.device))
        for t in keep:
            t.record_stream(torch.cuda.current_stream(self.device))

    def run(self, n_iters, lr=0.005, resume=False):
        """n_iters closure steps with a fresh Adam (optim_factory.py:77-80; S2.yaml lr 0.005); resume=True continues the previous call's
        moments and step count.  Asynchronous."""
        _lib.call('lemo_fit_prox_run', self.handle, int(n_iters), float(lr), int(bool(resume)), _lib.cur_stream(self.device))
        self.iters_run += n_iters

    def eval(self):
        """One closure evaluation (loss + gradients after the erase) without an optimiser step."""
        _lib.call('lemo_fit_prox_eval', self.handle, _lib.cur_stream(self.device))

    def _out(self):
        o = _lib.LemoProxParamsOutC()
        d = {k: torch.empty(self.B, n, device=self.device) for k, n in PARAM_DIMS.items()}
        for k, t in d.items():
            setattr(o, k, t.data_ptr())
        return o, d

    def params(self):
        o, d = self._out()
        _lib.call('lemo_fit_prox_get', self.handle, C.byref(o), None, None, _lib.cur_stream(self.device))
        return d

    def grads(self):
        o, d = self._out()
        _lib.call('lemo_fit_prox_get', self.handle, None, C.byref(o), None, _lib.cur_stream(self.device))
        return d

    def losses(self):
        """dict of the last closure's loss terms (device scalars, the reference's loss_dict names) incl. 'total_loss'."""
        l = torch.empty(16, device=self.device)
        _lib.call('lemo_fit_prox_get', self.handle, None, None, _lib.ptr(l), _lib.cur_stream(self.device))
        return {k: l[i] for i, k in enumerate(LOSS_KEYS) if not k.startswith('_')}

    def kernel_launches(self):
        return int(_lib.lib().lemo_fit_prox_kernel_launches(self.handle))
