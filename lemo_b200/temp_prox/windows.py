"""Sliding-window orchestration of a PROX recording and the per-frame result format -- the host logic either side of the fitting path
(SURVEY section 8f.3): temp_prox/data_parser_slide.py:199-212 (70 %-stride windows, consumed by a DataLoader with batch_size = window,
drop_last), :329-335 (initialise every frame from this run's results if they exist, else from the PROX results), :106-126 (result
reader); temp_prox/main_slide.py:327-330 (first window: nothing frozen, later windows: first 15 % frozen);
temp_prox/fit_temp_loadprox_slide.py:494-498 (window-mean betas) and :577-594 (one protocol-2 pickle per frame, every value [1, d]).

Pure host code: no kernels, no device work.  The file I/O of the reference (image, depth, OpenPose json) needs PROX data and stays out
of scope; what is kept is everything that decides WHICH frames a window fits, WHAT they start from and HOW results are stored, so a
recording can be streamed through `SMPLifyLoss` / `FittingMonitor` window by window.
"""
import os
import pickle

import numpy as np

PARAM_KEYS = ('transl', 'global_orient', 'betas', 'body_pose', 'pose_embedding', 'left_hand_pose', 'right_hand_pose', 'jaw_pose',
              'leye_pose', 'reye_pose', 'expression')


def sliding_windows(n_frames, batch_size):
    """Frame indices (0-based positions in the recording's frame list) of every window the reference fits, in order.

    The reference concatenates frames[0:B] and frames[s(i+1) : min(s(i+1)+B, n)] for i = 0 .. (n - B) - s with s = int(0.7 B)
    (data_parser_slide.py:200-210) and lets DataLoader(batch_size=B, drop_last=True, shuffle=False) cut the concatenation every B frames
    (main_slide.py:142-149).  Windows therefore overlap by 30 %, and whatever does not fill a last full batch is dropped."""
    B = int(batch_size)
    s = int(B * 0.7)
    ids = list(range(0, min(B, n_frames)))
    seq_n = (n_frames - B) - s
    for i in range(int(seq_n) + 1):
        start = s * (i + 1)
        end = min(start + B, n_frames)
        ids += list(range(start, end))
    return [np.asarray(ids[k * B:(k + 1) * B], np.int64) for k in range(len(ids) // B)]


def read_prox_pkl(pkl_path):
    """data_parser_slide.py:106-126: one frame's parameters, first (only) row of every array."""
    with open(pkl_path, 'rb') as f:
        data = pickle.load(f)
    return {k: data[k][0] for k in PARAM_KEYS}


def frame_result(body_params, camera_params, i, pose_embedding=None, body_pose=None):
    """fit_temp_loadprox_slide.py:577-589: the dict pickled for frame i of a window -- 'camera_<name>' and every body-model parameter as
    [1, d] arrays, plus pose_embedding [1, 32] and the decoded body_pose [1, 63] when VPoser is used."""
    res = {'camera_' + str(k): np.asarray(v)[i][None] for k, v in camera_params.items()}
    res.update({k: np.asarray(v)[i][None] for k, v in body_params.items()})
    if pose_embedding is not None:
        res['pose_embedding'] = np.asarray(pose_embedding)[i][None]
        res['body_pose'] = np.asarray(body_pose)[i][None]
    return res


class WindowChain:
    """Results of a recording keyed by frame name, on disk in the reference's layout `<dir>/results/<frame name>/000.pkl`.

    init_for(names) gives the [B, d] start values of a window: this run's result of a frame if it exists (the 30 % overlap with the
    previous window), else the PROX result (data_parser_slide.py:329-335), with betas replaced by their window mean
    (fit_temp_loadprox_slide.py:494-497).  store(names, ...) writes one protocol-2 pickle per frame (:591-594)."""

    def __init__(self, current_params_dir, prox_params_dir):
        self.current_params_dir, self.prox_params_dir = current_params_dir, prox_params_dir

    def _path(self, root, name):
        return os.path.join(root, 'results', name, '000.pkl')

    def init_for(self, names):
        rows = []
        for name in names:
            p = self._path(self.current_params_dir, name)
            if not os.path.exists(p):
                p = self._path(self.prox_params_dir, name)
            rows.append(read_prox_pkl(p))
        out = {k: np.stack([np.asarray(r[k]) for r in rows], 0) for k in PARAM_KEYS}
        mean_betas = np.mean(out['betas'], axis=0)
        out['betas'] = np.repeat(np.expand_dims(mean_betas, axis=0), len(names), axis=0)
        return out

    @staticmethod
    def erase_first(window_index):
        """main_slide.py:327-330 + fitting_temp_slide.py:281-288: every window but the first keeps its first int(0.15 B) frames fixed."""
        return window_index != 0

    def store(self, names, body_params, camera_params, pose_embedding=None, body_pose=None):
        for i, name in enumerate(names):
            p = self._path(self.current_params_dir, name)
            os.makedirs(os.path.dirname(p), exist_ok=True)
            with open(p, 'wb') as f:
                pickle.dump(frame_result(body_params, camera_params, i, pose_embedding, body_pose), f, protocol=2)
