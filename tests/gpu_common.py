"""Shared helpers for the -m gpu parity tests: synthetic model / VPoser / oracle context, built once per session."""
import functools
import numpy as np
import torch

from oracle import synth, ref_body as rb, ref_loops as rl

DEV = 'cuda:0'


@functools.lru_cache(maxsize=None)
def model_np(n_verts=synth.V, seed=0):
    return synth.make_smplx_model(seed, n_verts=n_verts)


@functools.lru_cache(maxsize=None)
def vposer_w():
    return synth.make_vposer_weights(1)


@functools.lru_cache(maxsize=None)
def smplx_module(n_verts=synth.V, batch=1):
    import lemo_b200.smplx as smplx
    return smplx.create(model_np(n_verts), model_type='smplx', gender='male', ext='npz', num_pca_comps=12,
                        batch_size=batch).to(DEV)


@functools.lru_cache(maxsize=None)
def vposer_module():
    from lemo_b200.vposer import VPoserDecoder
    return VPoserDecoder(vposer_w()).to(DEV)


@functools.lru_cache(maxsize=None)
def enc_module():
    from lemo_b200.fit import load_smooth_prior
    return load_smooth_prior().to(DEV)


@functools.lru_cache(maxsize=None)
def oracle_ctx(dtype=torch.float32, n_verts=synth.V):
    return rl.FitContext(model_np(n_verts), vposer_w(), synth.load_enc_weights(), synth.load_tables(), dtype=dtype)


def rand_pose(B, seed, scale=0.3):
    g = np.random.default_rng(seed)
    f = lambda *s: (scale * g.standard_normal(s)).astype(np.float32)
    return dict(transl=f(B, 3), global_orient=f(B, 3), body_pose=f(B, 63), jaw_pose=f(B, 3), leye_pose=f(B, 3),
                reye_pose=f(B, 3), left_hand_pose=f(B, 12), right_hand_pose=f(B, 12),
                betas=g.standard_normal((B, 10)).astype(np.float32), expression=f(B, 10))


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_q(a, b, q=0.99):
    """q-quantile of |a-b| / max|b| (robust to isolated LeakyReLU-kink sign flips, see test_gpu_priors.py)."""
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.quantile(np.abs(a - b), q) / max(np.abs(b).max(), 1e-30))
