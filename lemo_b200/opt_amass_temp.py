"""Drop-in for the reference's opt_amass_temp.py: same flags (:18-51), same `optimize()` entry point (:62), same input / output files
(reads `{perframe_res_dir}/{dataset_name}/body_params_opt_clip_{i}.npy`, writes `{save_dir}/{dataset_name}/gender_list.npy`,
`contact_lbl_rec_clip_{i}.npy [T,4]`, `body_params_opt_clip_{i}.npy [T,72] f32`: :233,:270,:457-458).

What runs where: infill pre-stage (:144-325) = InfillStage on the device; the temporal Adam loop (:329-455: B = T frames, marker L1 +
Enc smoothness prior + foot-contact velocity + L2 priors, 100 steps, lr .01 -> .005 after step 60) = lemo_fit_run for
`--seqs_per_batch` clips side by side, one CUDA graph per iteration.

    python -m lemo_b200.opt_amass_temp --synthetic_model --synthetic_clips 8 --start 0 --end 8 --step 1

`pipeline()` chains infill -> per-frame -> temporal on the device without the intermediate .npy round trip (bench.py `pipeline`).
"""
import os

import numpy as np
import torch

from . import amass_common as ac
from .fit import PerFrameFitter, TemporalFitter
from .models.AE_sep import Enc

TOTAL_STEPS = 100          # opt_amass_temp.py:347


def _load_smooth_encoder(args, device):
    enc = Enc(downsample=False, z_channel=64)
    enc.load_state_dict(ac.load_state(args.smooth_model_path, 'enc_smooth_15217.npz'))
    return enc.to(device)


def optimize(args=None, dataloader=None, init_params=None):
    """init_params: optional {clip id: [T,72]} (else read from --perframe_res_dir like the reference)."""
    args = ac.base_parser(temporal=True).parse_args([]) if args is None else args
    device = torch.device(args.device)
    T = args.clip_seconds * 30 - 1
    smplx_male, smplx_female, vposer = ac.load_models(args, device, T)
    print('[INFO] vposer model loaded')
    enc = _load_smooth_encoder(args, device)
    if dataloader is None:
        if args.synthetic_clips <= 0:
            raise RuntimeError('no dataloader given: pass the reference TrainLoader DataLoader, or use --synthetic_clips N')
        dataloader = ac.synthetic_dataloader(args.synthetic_clips, T + 1, device)
    print('[INFO] inference stage (with self-supervised finetuning)')
    clips, gender_list = ac.infill_all(args, dataloader, device)
    save_folder = os.path.join(args.save_dir, args.dataset_name)
    os.makedirs(save_folder, exist_ok=True)
    np.save('{}/gender_list.npy'.format(save_folder), gender_list)
    weights = dict(w_rec=args.weight_loss_rec_markers, w_contact=args.weight_loss_contact_vel, w_smooth=args.weight_loss_smooth,
                   w_vposer=args.weight_loss_vposer, w_shape=args.weight_loss_shape, w_hand=args.weight_loss_hand)
    print('[INFO] temporal optimizing ...')
    ids = ac.clip_ids(args, len(clips))
    results, fitters = {}, {}
    for g, batch in ac.batches(ids, clips, args.seqs_per_batch):
        S = len(batch)
        Tc = clips[batch[0]]['markers_rec'].shape[0]
        key = (g, S, Tc)
        if key not in fitters:
            fitters[key] = TemporalFitter(smplx_female if g == 0 else smplx_male, vposer, S, Tc, enc=enc, device=device, weights=weights)
        fit = fitters[key]
        for s, i in enumerate(batch):
            print('current clip:', i)
            if init_params is not None:
                init = np.asarray(init_params[i], np.float32)
            else:
                init = np.load(os.path.join(args.perframe_res_dir, args.dataset_name, 'body_params_opt_clip_{}.npy'.format(i)))   # [T, 72]
            np.save('{}/contact_lbl_rec_clip_{}.npy'.format(save_folder, i), clips[i]['contact'].cpu().numpy())
            fit.set_sequence(s, init, clips[i]['markers_rec'], clips[i]['contact'])
        fit.run(n_iters=TOTAL_STEPS, lr0=0.01, lr1=0.005, lr_switch=60)
        p72, _ = fit.results()
        p72 = p72.cpu().numpy()
        for s, i in enumerate(batch):
            np.save('{}/body_params_opt_clip_{}.npy'.format(save_folder, i), p72[s])
            results[i] = p72[s]
    return results


class Pipeline:
    """infill -> per-frame -> temporal for S clips of one gender, everything resident on one GPU (no .npy round trip between the stages).
    Handles are created once and reused for every batch of clips."""

    def __init__(self, smplx_model, vposer, infill_ae, enc, S, T, device='cuda', perframe_steps=100, temporal_steps=100):
        self.S, self.T, self.device = S, T, torch.device(device)
        self.pool = ac.InfillPool(infill_ae, n_streams=min(S, 8), device=self.device, stats=ac.load_infill_stats())
        self.pf = PerFrameFitter(smplx_model, vposer, S, T, device=self.device)
        self.tf = TemporalFitter(smplx_model, vposer, S, T, enc=enc, device=self.device)
        self.perframe_steps, self.temporal_steps = perframe_steps, temporal_steps

    def run(self, clip_imgs, rot0s, betas):
        """clip_imgs [S,4,208,T], rot0s [S], betas [S,10] -> (params72 [S,T,72] temporal result, contact [S,T,4]); asynchronous."""
        S = self.S
        outs = self.pool.run_many([clip_imgs[s] for s in range(S)], [rot0s[s:s + 1] for s in range(S)])     # S clips fine-tuned concurrently
        recs, cons = [o[0] for o in outs], [o[1] for o in outs]
        for s in range(S):
            self.pf.set_sequence(s, betas[s].detach().cpu().numpy() if torch.is_tensor(betas) else betas[s], recs[s])
        self.pf.run(n_iters=self.perframe_steps)
        init72, _ = self.pf.results()
        self.tf.set_sequences(init72, torch.stack(recs), torch.stack(cons))
        self.tf.run(n_iters=self.temporal_steps, lr0=0.01, lr1=0.005, lr_switch=60)
        p72, _ = self.tf.results()
        return p72, torch.stack(cons)


if __name__ == '__main__':
    optimize(ac.base_parser(temporal=True).parse_args())
