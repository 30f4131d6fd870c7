"""Export the reference's DATA tables (index lists, normalisation stats, shipped prior weights,
sample parameter clips) into the small asset/fixture files this repo ships.

Runs ONLY in the build container (needs /root/reference).  No reference *source* is copied: the
outputs are integer index tables, float statistics and network weights.

  lemo_b200/assets/lemo_tables.npz   marker / foot-vertex index tables + smooth/infill stats
  lemo_b200/assets/enc_smooth_15217.npz   Enc weights of runs/15217/Enc_last_model.pkl
  lemo_b200/assets/ae_infill_59547.npz    AE (infill prior) weights of runs/59547/AE_last_model.pkl
  tests/golden/seed_clips.npz        the ten shipped [119,72] result clips + contact labels

Index provenance (all under /root/reference):
  loader/SSM2.json, loader/SSM2_withhand.json              (opt_amass_temp.py:237-241)
  body_segments/{L,R}_Leg.json o foot_verts_id/*.npy       (opt_amass_temp.py:97-113)
The foot tables are resolved with the reference's own expression
`np.asarray(list(set(verts_ind)))[bool_mask]` in THIS interpreter.
"""
import json, os, sys
import numpy as np
import torch

REF = '/root/reference'
OUT_A = os.path.join(os.path.dirname(__file__), '..', 'lemo_b200', 'assets')
OUT_G = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden')


def main():
    j = lambda p: json.load(open(os.path.join(REF, p)))
    ssm2 = list(j('loader/SSM2.json')['markersets'][0]['indices'].values())
    ssm2h = list(j('loader/SSM2_withhand.json')['markersets'][0]['indices'].values())
    assert ssm2h[:67] == ssm2
    tabs = dict(markers67=np.asarray(ssm2, np.int32), markers81=np.asarray(ssm2h, np.int32))
    for side, S in (('left', 'L'), ('right', 'R')):
        leg = np.asarray(list(set(j('body_segments/%s_Leg.json' % S)['verts_ind'])))
        for part in ('heel', 'toe'):
            m = np.load(os.path.join(REF, 'foot_verts_id/%s_%s_verts_id.npy' % (side, part)))
            tabs['%s_%s' % (side, part)] = leg[m].astype(np.int32)
    st = np.load(os.path.join(REF, 'preprocess_stats/preprocess_stats_smooth_withHand_global_markers.npz'))
    tabs['smooth_Xmean'] = st['Xmean'].reshape(243).astype(np.float32)
    tabs['smooth_Xstd'] = st['Xstd'].reshape(243).astype(np.float32)
    st = np.load(os.path.join(REF, 'preprocess_stats/preprocess_stats_infill_local_markers_4chan.npz'))
    for k in st.files:
        tabs['infill_' + k] = np.asarray(st[k], np.float64)     # float64 like the file: the de-normalisation runs in numpy float64
    np.savez(os.path.join(OUT_A, 'lemo_tables.npz'), **tabs)
    print({k: v.shape for k, v in tabs.items()})

    w = torch.load(os.path.join(REF, 'runs/15217/Enc_last_model.pkl'), map_location='cpu')
    np.savez(os.path.join(OUT_A, 'enc_smooth_15217.npz'), **{k: v.numpy() for k, v in w.items()})
    w = torch.load(os.path.join(REF, 'runs/59547/AE_last_model.pkl'), map_location='cpu')
    np.savez_compressed(os.path.join(OUT_A, 'ae_infill_59547.npz'), **{k: v.numpy() for k, v in w.items()})

    clips = {}
    for stage in ('perframe', 'temp'):
        for c in (0, 20, 40, 60, 80):
            d = os.path.join(REF, 'res_opt_amass_%s/TotalCapture' % stage)
            clips['%s_params_%d' % (stage, c)] = np.load(os.path.join(d, 'body_params_opt_clip_%d.npy' % c))
            clips['%s_contact_%d' % (stage, c)] = np.load(os.path.join(d, 'contact_lbl_rec_clip_%d.npy' % c))
    np.savez_compressed(os.path.join(OUT_G, 'seed_clips.npz'), **clips)


def export_prox_tables():
    """lemo_b200/assets/prox_tables.npz: OpenPose joint maps of the reference's own smpl_to_openpose (temp_prox/misc_utils.py:87-197) for
    model_type='smplx', every (use_hands, use_face, use_face_contour) combination and both OpenPose formats; the friction (307) and contact
    (1121) vertex-id lists built with the reference's expressions (fit_temp_loadprox_slide.py:349-362, `list(set(...))` in THIS
    interpreter); and small known-answer vectors of the reference's prior / robustifier modules (prior.py:53-98, misc_utils.py:61-85)."""
    sys.path.insert(0, os.path.join(REF, 'temp_prox'))
    import misc_utils as mu
    import prior as pr
    j = lambda p: json.load(open(os.path.join(REF, p)))
    tabs = {}
    for fmt in ('coco25', 'coco19'):
        for h in (0, 1):
            for f in (0, 1):
                for c in (0, 1):
                    m = mu.smpl_to_openpose('smplx', use_hands=bool(h), use_face=bool(f), use_face_contour=bool(c), openpose_format=fmt)
                    tabs['smplx_%s_h%d_f%d_c%d' % (fmt, h, f, c)] = np.asarray(m, np.int64)
    seg = lambda part: list(set(j('body_segments/%s.json' % part)['verts_ind']))
    tabs['friction_ids'] = np.concatenate([seg(p) for p in ('L_Leg', 'R_Leg', 'gluteus')]).astype(np.int64)
    tabs['contact_ids'] = np.concatenate([seg(p) for p in ('L_Leg', 'R_Leg', 'L_Hand', 'R_Hand', 'gluteus', 'back', 'thighs')]).astype(np.int64)
    np.savez(os.path.join(OUT_A, 'prox_tables.npz'), **tabs)
    print({k: v.shape for k, v in tabs.items()})
    # known-answer vectors of the reference modules themselves
    g = torch.Generator().manual_seed(7)
    pose = 0.5 * torch.randn(5, 63, generator=g)
    pose_g = 0.5 * torch.randn(5, 66, generator=g)
    res = 3.0 * torch.randn(4, 118, 2, generator=g)
    gold = dict(pose=pose.numpy(), pose_g=pose_g.numpy(), res=res.numpy(),
                angle=pr.SMPLifyAnglePrior()(pose).numpy(), angle_g=pr.SMPLifyAnglePrior()(pose_g, with_global_pose=True).numpy(),
                l2=pr.L2Prior()(pose).numpy(), gmof100=mu.GMoF(rho=100)(res).numpy(), gmof_unscaled=mu.GMoF_unscaled(rho=0.5)(res).numpy(),
                mapped=mu.JointMapper(tabs['smplx_coco25_h1_f1_c0'])(torch.arange(127.).view(1, 127, 1).expand(2, 127, 3)).numpy())
    np.savez_compressed(os.path.join(OUT_G, 'reference_golden_prox.npz'), **gold)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'prox':
        export_prox_tables()
    else:
        main()
