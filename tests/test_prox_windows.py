"""CPU: PROX sliding-window orchestration and result format (lemo_b200/temp_prox/windows.py) against golden tables produced by executing
the reference's own window-building lines + torch DataLoader (tools/make_window_golden.py -> tests/golden/reference_golden_windows.npz)."""
import os
import pickle

import numpy as np
import pytest

from lemo_b200.temp_prox import windows as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden_windows.npz')


def test_sliding_windows_match_reference():
    gold = np.load(GOLD)
    assert len(gold.files) >= 10
    for key in gold.files:
        n, B = (int(x[1:]) for x in key.split('_'))
        got = W.sliding_windows(n, B)
        want = gold[key]
        assert len(got) == want.shape[0], key
        for g, w in zip(got, want):
            assert np.array_equal(g, w), key                      # integer tables: bit-exact


def test_window_properties():
    ws = W.sliding_windows(1000, 100)
    assert [int(w[0]) for w in ws[:4]] == [0, 70, 140, 210]       # 70 % stride (data_parser_slide.py:200)
    assert all(len(w) == 100 for w in ws)                          # drop_last
    assert len(np.intersect1d(ws[0], ws[1])) == 30                 # 30 % overlap: the frames a later window starts from and freezes half of
    assert W.sliding_windows(99, 100) == []                        # shorter than one window: nothing to fit
    q = W.sliding_windows(305, 100)[-1]                            # the reference's quirk: leftovers of two truncated slices form one batch
    assert np.array_equal(q, np.concatenate([np.arange(210, 305), np.arange(280, 285)]))
    assert not W.WindowChain.erase_first(0) and W.WindowChain.erase_first(1)


def _frame(g, shift=0.0):
    d = {'transl': 3, 'global_orient': 3, 'betas': 10, 'body_pose': 63, 'pose_embedding': 32, 'left_hand_pose': 12, 'right_hand_pose': 12,
         'jaw_pose': 3, 'leye_pose': 3, 'reye_pose': 3, 'expression': 10}
    return {k: (g.standard_normal((1, n)) + shift).astype(np.float32) for k, n in d.items()}


def test_window_chain_init_and_store(tmp_path):
    g = np.random.default_rng(0)
    prox, cur = str(tmp_path / 'prox'), str(tmp_path / 'cur')
    names = ['f%03d' % i for i in range(6)]
    frames = {}
    for n in names:                                                # PROX results for every frame, in the reference's pickle layout
        frames[n] = _frame(g)
        os.makedirs(os.path.join(prox, 'results', n))
        with open(os.path.join(prox, 'results', n, '000.pkl'), 'wb') as f:
            pickle.dump(frames[n], f, protocol=2)
    chain = W.WindowChain(cur, prox)
    init = chain.init_for(names[:4])
    assert init['transl'].shape == (4, 3) and init['pose_embedding'].shape == (4, 32)
    assert np.array_equal(init['transl'][2], frames['f002']['transl'][0])
    mean_betas = np.mean(np.stack([frames[n]['betas'][0] for n in names[:4]]), 0)
    assert np.array_equal(init['betas'], np.repeat(mean_betas[None], 4, 0))                 # fit_temp_loadprox_slide.py:494-497
    # store this window, then the next window starts from THIS run's values on the overlap and from PROX elsewhere
    body = {k: v + 10.0 for k, v in init.items() if k not in ('pose_embedding', 'body_pose')}
    cam = {'rotation': np.tile(np.eye(3, dtype=np.float32)[None], (4, 1, 1)), 'translation': np.zeros((4, 3), np.float32)}
    chain.store(names[:4], body, cam, pose_embedding=init['pose_embedding'] + 10.0, body_pose=init['body_pose'] + 10.0)
    with open(os.path.join(cur, 'results', 'f001', '000.pkl'), 'rb') as f:
        res = pickle.load(f)
    assert res['transl'].shape == (1, 3) and res['camera_rotation'].shape == (1, 3, 3) and res['pose_embedding'].shape == (1, 32)
    assert set(res) == set(body) | {'camera_rotation', 'camera_translation', 'pose_embedding', 'body_pose'}
    nxt = chain.init_for(names[2:6])
    assert np.array_equal(nxt['transl'][0], body['transl'][2]) and np.array_equal(nxt['transl'][3], frames['f005']['transl'][0])
    assert np.array_equal(W.read_prox_pkl(os.path.join(cur, 'results', 'f003', '000.pkl'))['global_orient'], body['global_orient'][3])
