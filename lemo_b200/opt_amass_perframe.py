"""Drop-in for the reference's opt_amass_perframe.py: same flags (:18-44), same `optimize()` entry point (:55), same output files
(`{save_dir}/{dataset_name}/gender_list.npy`, `contact_lbl_rec_clip_{i}.npy [T,4]`, `body_params_opt_clip_{i}.npy [T,72] f32`, :222,:240,:364).

What runs where: infill pre-stage (mask, 60-step AE fine-tune, inference, de-normalisation, global reconstruction: :117-288) =
InfillStage on the device; the per-frame Adam loop (:293-361: T warm-started B=1 problems x 100 steps, lr .1/.01 -> .01@>60 ->
.003@>80) = lemo_fit_run_perframe, `--seqs_per_batch` clips side by side.  The host only reads clips and writes .npy files.

    python -m lemo_b200.opt_amass_perframe --synthetic_model --synthetic_clips 8 --start 0 --end 8 --step 1
"""
import os

import numpy as np
import torch

from . import amass_common as ac
from .fit import PerFrameFitter

TOTAL_STEPS = 100          # opt_amass_perframe.py:321


def optimize(args=None, dataloader=None):
    args = ac.base_parser(temporal=False).parse_args([]) if args is None else args
    device = torch.device(args.device)
    T = args.clip_seconds * 30 - 1                                   # :78 / loader: velocity representation drops a frame
    smplx_male, smplx_female, vposer = ac.load_models(args, device, 1)
    print('[INFO] vposer / smplx models loaded')
    if dataloader is None:
        if args.synthetic_clips <= 0:
            raise RuntimeError('no dataloader given: pass the reference TrainLoader DataLoader, or use --synthetic_clips N')
        dataloader = ac.synthetic_dataloader(args.synthetic_clips, T + 1, device)
    print('[INFO] inference stage (with self-supervised finetuning)')
    clips, gender_list = ac.infill_all(args, dataloader, device)
    save_folder = os.path.join(args.save_dir, args.dataset_name)
    os.makedirs(save_folder, exist_ok=True)
    np.save('{}/gender_list.npy'.format(save_folder), gender_list)
    weights = dict(w_rec=args.weight_loss_rec_markers, w_vposer=args.weight_loss_vposer, w_shape=args.weight_loss_shape,
                   w_hand=args.weight_loss_hand)
    print('[INFO] optimizing per frame...')
    ids = ac.clip_ids(args, len(clips))
    results = {}
    fitters = {}
    for g, batch in ac.batches(ids, clips, args.seqs_per_batch):
        S = len(batch)
        Tc = clips[batch[0]]['markers_rec'].shape[0]
        key = (g, S, Tc)
        if key not in fitters:
            fitters[key] = PerFrameFitter(smplx_female if g == 0 else smplx_male, vposer, S, Tc, device=device, weights=weights)
        fit = fitters[key]
        for s, i in enumerate(batch):
            print('current clip:', i)
            np.save('{}/contact_lbl_rec_clip_{}.npy'.format(save_folder, i), clips[i]['contact'].cpu().numpy())
            fit.set_sequence(s, clips[i]['beta'].cpu().numpy(), clips[i]['markers_rec'])
        fit.run(n_iters=TOTAL_STEPS)
        p72, _ = fit.results()
        p72 = p72.cpu().numpy()
        for s, i in enumerate(batch):
            np.save('{}/body_params_opt_clip_{}.npy'.format(save_folder, i), p72[s])       # [T, 72]
            results[i] = p72[s]
    return results


if __name__ == '__main__':
    optimize(ac.base_parser(temporal=False).parse_args())
