"""Script-level call surface: `optimize()` of the mirrors of opt_amass_perframe.py / opt_amass_temp.py keeps the reference's
command-line flags (names and defaults transcribed from opt_amass_perframe.py:18-44 and opt_amass_temp.py:18-51) and output files."""
import os

import numpy as np
import pytest
import torch

REF_FLAGS_COMMON = {'amass_dir': '/local/home/szhang/AMASS/amass', 'body_model_path': '/mnt/hdd/PROX/body_models', 'clip_seconds': 4,
                    'body_mode': 'local_markers_4chan', 'infill_model_path': 'runs/59547/AE_last_model.pkl', 'conv_k': 3, 'start': 0, 'end': 100,
                    'step': 20, 'dataset_name': 'TotalCapture', 'weight_loss_rec_markers': 1.0, 'weight_loss_vposer': 0.02,
                    'weight_loss_shape': 0.01, 'weight_loss_hand': 0.01}
REF_FLAGS_TEMP = dict(REF_FLAGS_COMMON, smooth_model_path='runs/15217/Enc_last_model.pkl', perframe_res_dir='res_opt_amass_perframe',
                      save_dir='res_opt_amass_temp', weight_loss_contact_vel=0.03, weight_loss_smooth=1e6)
REF_FLAGS_PERFRAME = dict(REF_FLAGS_COMMON, save_dir='res_opt_amass_perframe')


def test_flags_match_the_reference_scripts():
    from lemo_b200 import amass_common as ac
    for temporal, ref in ((True, REF_FLAGS_TEMP), (False, REF_FLAGS_PERFRAME)):
        a = vars(ac.base_parser(temporal).parse_args([]))
        for k, v in ref.items():
            assert a[k] == v, (k, a[k], v)
        extra = set(a) - set(ref)
        assert extra == {'synthetic_clips', 'synthetic_model', 'seqs_per_batch', 'device'}, extra
    import lemo_b200.opt_amass_temp as t, lemo_b200.opt_amass_perframe as p
    assert callable(t.optimize) and callable(p.optimize) and t.TOTAL_STEPS == 100 and p.TOTAL_STEPS == 100


@pytest.mark.gpu
def test_perframe_then_temporal_scripts_write_the_reference_files(tmp_path):
    """python -m lemo_b200.opt_amass_perframe ... ; python -m lemo_b200.opt_amass_temp ... on 3 synthetic clips: the files the reference
    scripts write exist with the reference shapes / dtypes, the temporal stage reads the per-frame stage's files, and it ends closer to
    the infilled markers' smooth solution (lower temporal loss) than the per-frame initialisation it started from."""
    from lemo_b200 import amass_common as ac
    import lemo_b200.opt_amass_temp as t, lemo_b200.opt_amass_perframe as p
    pf_dir, tp_dir = str(tmp_path / 'pf'), str(tmp_path / 'tp')
    common = ['--synthetic_model', '--synthetic_clips', '3', '--start', '0', '--end', '3', '--step', '1', '--clip_seconds', '1',
              '--seqs_per_batch', '2', '--device', 'cuda:0']
    res_pf = p.optimize(ac.base_parser(False).parse_args(common + ['--save_dir', pf_dir]))
    res_tp = t.optimize(ac.base_parser(True).parse_args(common + ['--save_dir', tp_dir, '--perframe_res_dir', pf_dir]))
    T = 1 * 30 - 1
    for d, res in ((pf_dir, res_pf), (tp_dir, res_tp)):
        folder = os.path.join(d, 'TotalCapture')
        g = np.load(os.path.join(folder, 'gender_list.npy'))
        assert g.shape == (3, 1) and set(g.ravel().tolist()) <= {0, 1}
        for i in range(3):
            bp = np.load(os.path.join(folder, 'body_params_opt_clip_%d.npy' % i))
            cl = np.load(os.path.join(folder, 'contact_lbl_rec_clip_%d.npy' % i))
            assert bp.shape == (T, 72) and bp.dtype == np.float32 and np.isfinite(bp).all()
            assert cl.shape == (T, 4) and set(np.unique(cl).tolist()) <= {0.0, 1.0}
            assert np.array_equal(bp, res[i])
    # the temporal stage starts from the per-frame files and moves
    assert any(np.abs(res_tp[i] - res_pf[i]).max() > 1e-4 for i in range(3))
    # shape (betas) is never optimised: columns 6:16 are the clip's betas in both stages
    for i in range(3):
        assert np.allclose(res_tp[i][:, 6:16], res_pf[i][:, 6:16])
        assert np.allclose(res_pf[i][:, 6:16], res_pf[i][0:1, 6:16])
