"""Shared host logic of the two AMASS fitting scripts (mirrors of opt_amass_perframe.py / opt_amass_temp.py): command-line flags,
model / prior loading, the clip source, the infill pre-stage loop and the result files.

Everything numeric runs on the device through the C ABI (InfillStage, PerFrameFitter, TemporalFitter); this module only moves clips
in and results out.  Out of scope: the AMASS reader `loader.optimize_loader_amass_new.TrainLoader` -- any
iterable that yields the reference DataLoader's 6-tuples `[clip_img [1,4,208,T], smplx_beta [1,10], gender [1], rot_0_pivot [1], _, _]`
can be passed as `dataloader`; without one, `--synthetic_clips N` builds N AMASS-shaped clips on the device.
"""
import argparse
import glob
import os

import numpy as np
import torch

from . import synth
from . import smplx as smplx_mod
from .vposer import VPoserDecoder
from .infill import InfillStage, InfillPool, body_repr, load_infill_prior, load_infill_stats
from .models.AE import AE

_HERE = os.path.dirname(os.path.abspath(__file__))


def base_parser(temporal):
    """The reference's flags (opt_amass_perframe.py:18-44 / opt_amass_temp.py:18-51), same names, defaults and help strings."""
    p = argparse.ArgumentParser()
    p.add_argument('--amass_dir', type=str, default='/local/home/szhang/AMASS/amass', help='path to AMASS dataset')
    p.add_argument('--body_model_path', type=str, default='/mnt/hdd/PROX/body_models', help='path to smplx body models')
    p.add_argument('--clip_seconds', default=4, type=int, help='length (seconds) of each motion sequence')
    p.add_argument('--body_mode', type=str, default='local_markers_4chan', choices=['local_markers', 'local_markers_4chan'],
                   help='which body representation to use')
    p.add_argument('--infill_model_path', type=str, default='runs/59547/AE_last_model.pkl', help='path to pretrained infilling prior')
    p.add_argument('--conv_k', default=3, type=int, help='conv kernel size')
    if temporal:
        p.add_argument('--smooth_model_path', type=str, default='runs/15217/Enc_last_model.pkl', help='path to pretrained smoothness prior')
    p.add_argument('--start', default=0, type=int, help='from which sequence to start')
    p.add_argument('--end', default=100, type=int, help='until which sequence to end')
    p.add_argument('--step', default=20, type=int, help='optimize 1 sequence every [step] sequences')
    p.add_argument('--dataset_name', type=str, default='TotalCapture', help='which dataset in AMASS to optimize')
    if temporal:
        p.add_argument('--perframe_res_dir', type=str, default='res_opt_amass_perframe', help='path to body params optimized per frame')
        p.add_argument('--save_dir', type=str, default='res_opt_amass_temp', help='path to save optimized body params')
    else:
        p.add_argument('--save_dir', type=str, default='res_opt_amass_perframe', help='path to save optimized body params')
    p.add_argument('--weight_loss_rec_markers', type=float, default=1.0, help='weight for marker reconstruction loss (motion infilling prior)')
    if temporal:
        p.add_argument('--weight_loss_contact_vel', type=float, default=0.03, help='weight for foot contact friction loss')
        p.add_argument('--weight_loss_smooth', type=float, default=1e6, help='weight for smoothness loss (motion smoothness prior)')
    p.add_argument('--weight_loss_vposer', type=float, default=0.02, help='weight for vposer prior loss')
    p.add_argument('--weight_loss_shape', type=float, default=0.01, help='weight for body shape prior loss')
    p.add_argument('--weight_loss_hand', type=float, default=0.01, help='weight for hand pose prior loss')
    # ---- additions of this engine (not in the reference)
    p.add_argument('--synthetic_clips', type=int, default=0,
                   help='[lemo_b200] fit N synthetic AMASS-shaped clips instead of reading AMASS (no licensed data needed)')
    p.add_argument('--synthetic_model', action='store_true',
                   help='[lemo_b200] random SMPL-X-shaped body model and VPoser weights instead of the licensed files under --body_model_path')
    p.add_argument('--seqs_per_batch', type=int, default=8, help='[lemo_b200] clips fitted side by side on the GPU')
    p.add_argument('--device', type=str, default='cuda', help='[lemo_b200] CUDA device (there is no CPU path)')
    return p


def load_vposer(vposer_model_path, vp_model='snapshot', device='cuda'):
    """human_body_prior.tools.model_loader.load_vposer (:56-69) for the decoder this path uses: newest snapshots/*.pt state_dict."""
    pts = sorted(glob.glob(os.path.join(vposer_model_path, 'snapshots', '*.pt')), key=os.path.getmtime)
    if not pts:
        raise FileNotFoundError('no VPoser snapshot under %s/snapshots' % vposer_model_path)
    sd = torch.load(pts[-1], map_location='cpu')
    return VPoserDecoder(sd).to(device), None


def load_models(args, device, batch_size):
    """-> (smplx_male, smplx_female, vposer) as the scripts build them (opt_amass_temp.py:66-87)."""
    if args.synthetic_model:
        male = smplx_mod.create(synth.make_smplx_model(0), model_type='smplx', gender='male', ext='npz', num_pca_comps=12,
                                batch_size=batch_size).to(device)
        female = smplx_mod.create(synth.make_smplx_model(1), model_type='smplx', gender='female', ext='npz', num_pca_comps=12,
                                  batch_size=batch_size).to(device)
        return male, female, VPoserDecoder(synth.make_vposer_weights(1)).to(device)
    smplx_model_path = os.path.join(args.body_model_path, 'smplx_model')
    vposer_model_path = os.path.join(args.body_model_path, 'vposer_v1_0')
    vposer, _ = load_vposer(vposer_model_path, vp_model='snapshot', device=device)
    kw = dict(model_type='smplx', ext='npz', num_pca_comps=12, create_global_orient=True, create_body_pose=True, create_betas=True,
              create_left_hand_pose=True, create_right_hand_pose=True, create_expression=True, create_jaw_pose=True, create_leye_pose=True,
              create_reye_pose=True, create_transl=True, batch_size=batch_size)
    return (smplx_mod.create(smplx_model_path, gender='male', **kw).to(device),
            smplx_mod.create(smplx_model_path, gender='female', **kw).to(device), vposer)


def load_state(path, fallback_asset):
    """torch.load of a reference checkpoint (runs/*/..._last_model.pkl); the shipped weights are also packaged as assets/*.npz."""
    if os.path.exists(path):
        return torch.load(path, map_location=lambda storage, loc: storage)
    return dict(np.load(os.path.join(_HERE, 'assets', fallback_asset)))


def load_infill_model(args):
    if args.body_mode != 'local_markers_4chan':
        raise NotImplementedError("body_mode 'local_markers' (1-channel AE) is not on this path; the shipped prior is local_markers_4chan")
    ae = AE(downsample=True, in_channel=4, kernel=args.conv_k)
    ae.load_state_dict(load_state(args.infill_model_path, 'ae_infill_59547.npz'))
    return ae


def synthetic_dataloader(n_clips, T_frames, device):
    """AMASS-shaped synthetic clips in the reference DataLoader's tuple format (T = clip_seconds*30 marker frames -> T-1 columns)."""
    st64 = load_infill_stats()
    for s in range(n_clips):
        body68, con = synth.synth_marker_clip(100 + s, T=T_frames)
        clip, rot0 = body_repr(torch.from_numpy(body68).to(device), torch.from_numpy(con).to(device), stats=st64, device=device)
        g = np.random.default_rng(1000 + s)
        beta = torch.from_numpy((0.5 * g.standard_normal((1, 10))).astype(np.float32)).to(device)
        gender = torch.tensor([s % 2], device=device)
        yield [clip.unsqueeze(0), beta, gender, rot0.reshape(1), None, None]


def infill_all(args, dataloader, device):
    """The inference stage with self-supervised fine-tuning (opt_amass_temp.py:144-221) + the per-clip post-processing
    (:262-325), on the device.  -> list of dicts(markers_rec [T,67,3], contact [T,4], beta [10], gender int) and the gender array."""
    pool = InfillPool(load_infill_model(args), n_streams=max(1, min(8, args.seqs_per_batch)), device=device, stats=load_infill_stats())
    clips, genders, pending = [], [], []

    def flush():
        outs = pool.run_many([p[0] for p in pending], [p[1] for p in pending])          # a batch of clips fine-tuned concurrently
        for (clip_img, rot0, beta, gender), (m_rec, con, _) in zip(pending, outs):
            clips.append(dict(markers_rec=m_rec, contact=con, beta=beta, gender=gender))
            genders.append(gender)
        pending.clear()
    for step, data in enumerate(dataloader):
        if step == args.end:
            break
        clip_img, smplx_beta, gender, rot_0_pivot = data[0], data[1], data[2], data[3]
        pending.append((torch.as_tensor(clip_img).to(device)[0], torch.as_tensor(rot_0_pivot).reshape(-1)[0:1],
                        torch.as_tensor(smplx_beta).reshape(-1)[:10].float().to(device), int(torch.as_tensor(gender).reshape(-1)[0])))
        if len(pending) == len(pool.stages):
            flush()
    if pending:
        flush()
    return clips, np.asarray(genders).reshape(-1, 1)


def clip_ids(args, n_available):
    return [i for i in range(args.start, args.end, args.step) if i < n_available]


def batches(ids, clips, per_batch):
    """Group clip ids by gender (one body model per fitter) into batches of <= per_batch."""
    for g in (0, 1):
        sel = [i for i in ids if clips[i]['gender'] == g]
        for k in range(0, len(sel), per_batch):
            yield g, sel[k:k + per_batch]
