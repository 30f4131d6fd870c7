"""CPU restatement of the infill pre-stage (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows, in numpy float64 exactly where the reference is float64:
  * body representation  `get_local_markers_4chan`          /root/reference/utils/utils.py:209-265
  * inverse               `reconstruct_global_body`          /root/reference/utils/utils.py:180-203
  * quaternion algebra    Quaternions.{__mul__, between, from_angle_axis}, Pivots.from_quaternions
                          /root/reference/utils/Quaternions.py:71-118,396-406, utils/Pivots.py:79-88
  * mask / pad / fine-tune / crop / contact labels / de-normalise   /root/reference/opt_amass_temp.py:152-325
Pinned against the reference's own functions by tests/golden/reference_golden_infill.npz (oracle/make_golden.py imports
utils.utils from /root/reference with a stub for the absent torchgeometry and records inputs + outputs).
"""
import numpy as np

# marker ids whose rows are blanked before the AE sees the clip (opt_amass_temp.py:167-168)
MASK_MARKER_ID = np.array([14, 15, 18, 19, 29, 2, 20, 21, 30, 25, 16, 45, 46, 48, 49, 59, 32, 50, 51, 55, 60, 47])
PAD = (8, 8, 1, 1)          # left, right, top, bottom (opt_amass_temp.py:183)


# ------------------------------------------------------------------------------------------------ quaternions (w, x, y, z)
def q_mul(q, r):
    q0, q1, q2, q3 = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    r0, r1, r2, r3 = r[..., 0], r[..., 1], r[..., 2], r[..., 3]
    return np.stack([r0 * q0 - r1 * q1 - r2 * q2 - r3 * q3,
                     r0 * q1 + r1 * q0 - r2 * q3 + r3 * q2,
                     r0 * q2 + r1 * q3 + r2 * q0 - r3 * q1,
                     r0 * q3 - r1 * q2 + r2 * q1 + r3 * q0], axis=-1)


def q_conj(q):
    return q * np.array([1.0, -1.0, -1.0, -1.0])


def q_rot(q, v):
    """rotate vectors v [...,3] by q [...,4] (broadcast): imaginary part of q (0,v) q*"""
    vs = np.concatenate([np.zeros(v.shape[:-1] + (1,)), v], axis=-1)
    shape = np.broadcast_shapes(q.shape[:-1], v.shape[:-1]) + (4,)
    qb, vb = np.broadcast_to(q, shape), np.broadcast_to(vs, shape)
    return q_mul(qb, q_mul(vb, q_conj(qb)))[..., 1:4]


def q_between(v0, v1):
    a = np.cross(v0, v1)
    w = np.sqrt((v0 ** 2).sum(-1) * (v1 ** 2).sum(-1)) + (v0 * v1).sum(-1)
    q = np.concatenate([w[..., None], a], axis=-1)
    return q / np.sqrt((q ** 2).sum(-1))[..., None]


def q_angle_axis(angle, axis):
    axis = axis / (np.sqrt(np.sum(axis ** 2, axis=-1)) + 1e-10)[..., None]
    angle = np.asarray(angle, np.float64)
    return np.concatenate([np.cos(angle / 2.0)[..., None], axis * np.sin(angle / 2.0)[..., None]], axis=-1)


def pivot(q):
    """Pivots.from_quaternions(q).ps: heading of q's rotated +z axis in the xz plane"""
    d = q_rot(q, np.broadcast_to(np.array([0.0, 0.0, 1.0]), q.shape[:-1] + (3,)))
    return np.arctan2(d[..., 0], d[..., 2])


def gaussian_filter1d_nearest(x, sigma, axis=0, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d(x, sigma, axis, mode='nearest') restated (order 0)."""
    r = int(truncate * float(sigma) + 0.5)
    k = np.exp(-0.5 / (sigma * sigma) * np.arange(-r, r + 1) ** 2)
    k = k / k.sum()
    x = np.moveaxis(np.asarray(x, np.float64), axis, 0)
    n = x.shape[0]
    out = np.zeros_like(x)
    for j, w in zip(range(-r, r + 1), k):
        idx = np.clip(np.arange(n) + j, 0, n - 1)
        out += w * x[idx]
    return np.moveaxis(out, 0, axis)


# ------------------------------------------------------------------------------------------------ body representation
def get_local_markers_4chan(cur_body, contact_lbls):
    """cur_body [T, 1+67, 3] (pelvis + SSM2 markers, world z up), contact [T,4] -> ([4, T-1, 208] float64, rot_0_pivot [1]).
    The input dtype matters like in the reference: the floor shift runs in the input's dtype, everything after in float64."""
    cur_body = np.array(cur_body, copy=True)
    cur_body[:, :, [1, 2]] = cur_body[:, :, [2, 1]]
    cur_body[:, :, 1] = cur_body[:, :, 1] - cur_body[:, :, 1].min()
    reference = cur_body[:, 0] * np.array([1, 0, 1])
    cur_body = np.concatenate([reference[:, None], cur_body], axis=1)           # float64 from here on
    velocity = (cur_body[1:, 0:1] - cur_body[0:-1, 0:1]).copy()
    cur_body[:, :, 0] = cur_body[:, :, 0] - cur_body[:, 0:1, 0]
    cur_body[:, :, 2] = cur_body[:, :, 2] - cur_body[:, 0:1, 2]
    sdr_l, sdr_r, hip_l, hip_r = 28, 58, 29, 59
    across = (cur_body[:, sdr_r] - cur_body[:, sdr_l]) + (cur_body[:, hip_r] - cur_body[:, hip_l])
    across = across / np.sqrt((across ** 2).sum(-1))[..., None]
    forward = np.cross(across, np.array([[0, 1, 0]]))
    forward = gaussian_filter1d_nearest(forward, 20, axis=0)
    forward = forward / np.sqrt((forward ** 2).sum(-1))[..., None]
    target = np.array([[0, 0, 1]]).repeat(len(forward), axis=0)
    rotation = q_between(forward, target)[:, None]                                # [T,1,4]
    cur_body = q_rot(rotation, cur_body)
    velocity = q_rot(rotation[1:], velocity)
    rvelocity = pivot(q_mul(rotation[1:], q_conj(rotation[:-1])))                 # [T-1,1]
    rot_0_pivot = pivot(rotation[0])                                              # [1]
    cur_body[:, :, [1, 2]] = cur_body[:, :, [2, 1]]
    cur_body = cur_body[0:-1, 1:, :].reshape(len(cur_body) - 1, -1)
    local = np.concatenate([cur_body, contact_lbls[0:-1]], axis=-1)[None]
    T, d = local.shape[1], local.shape[2]
    chan = lambda v: np.repeat(v, d).reshape(1, T, d)
    return np.concatenate([local, chan(velocity[:, :, 0]), chan(velocity[:, :, 2]), chan(rvelocity)], axis=0), rot_0_pivot


def reconstruct_global_body(body_joints_input, rot_0_pivot):
    """[T, 1+68+1, 3] = zero reference + local (pelvis + markers) + global trajectory (vx, vy, r) -> [T, 68, 3] world positions."""
    x = np.array(body_joints_input, np.float64, copy=True)
    root = x[:, -1]
    root_r, root_x, root_z = root[:, 2], root[:, 0], root[:, 1]
    x = x[:, 0:-1]
    x[:, :, [1, 2]] = x[:, :, [2, 1]]
    rotation = np.array([[1.0, 0.0, 0.0, 0.0]])
    translation = np.array([[0.0, 0.0, 0.0]])
    yaxis = np.array([0.0, 1.0, 0.0])
    for i in range(len(x)):
        if i == 0:
            rotation = q_mul(q_angle_axis(-np.asarray(rot_0_pivot, np.float64).reshape(1), yaxis), rotation)
        x[i] = q_rot(rotation, x[i])
        x[i, :, 0] += translation[0, 0]
        x[i, :, 2] += translation[0, 2]
        rotation = q_mul(q_angle_axis(-root_r[i:i + 1], yaxis), rotation)
        translation = translation + q_rot(rotation, np.array([[root_x[i], 0.0, root_z[i]]]))
    x[:, :, [1, 2]] = x[:, :, [2, 1]]
    return x[:, 1:, :]


# ------------------------------------------------------------------------------------------------ the clip-level pipeline
def mask_rows(d_rows=208):
    """rows of channel 0 zeroed before the AE (markers' xyz rows, +3 for the pelvis) and the rows the fine-tune loss uses
    (indices into the PADDED image, which has one reflected row on top) -- opt_amass_temp.py:167-203"""
    r1 = MASK_MARKER_ID * 3 + 3
    masked = np.concatenate([r1, r1 + 1, r1 + 2])
    loss_rows = sorted(set(range(d_rows + 2)) - set((masked + 1).tolist()))[0:-5]
    return masked, np.asarray(loss_rows)


def prepare_input(clip_img):
    """clip_img [4, 208, T] (normalised) -> masked + reflect-padded [4, 210, T+16] float32"""
    x = np.array(clip_img, np.float32, copy=True)
    masked, _ = mask_rows(x.shape[1])
    x[0, masked, :] = 0.0
    x[0, -4:, :] = 0.0
    return np.pad(x, ((0, 0), (PAD[2], PAD[3]), (PAD[0], PAD[1])), mode='reflect')


def finalize(rec_pad, clip_img, stats, rot_0_pivot):
    """AE output on the padded clip [210, T+16] + the unmasked clip [4,208,T] -> (markers_rec [T,67,3] f32, contact [T,4] f32,
    markers_input [T,67,3] f32): crop, contact labels, de-normalise (float64 stats), global reconstruction
    (opt_amass_temp.py:205-325)."""
    rec = np.asarray(rec_pad, np.float32)[1:-1, 8:-8]                                  # [208, T]
    clip = np.asarray(clip_img, np.float32)
    T = rec.shape[-1]
    import torch
    sig = torch.sigmoid(torch.from_numpy(np.ascontiguousarray(rec[-4:, :].T))).numpy()
    contact = np.where(sig > 0.5, 1.0, 0.0).astype(np.float32)
    traj = np.stack([clip[1, 0], clip[2, 0], clip[3, 0]], axis=0)                       # [3, T]
    outs = []
    for body in (rec[0:-4, :], clip[0, 0:-4, :]):
        bj = np.concatenate([traj, body], axis=0).T.reshape(T, -1).astype(np.float32)   # [T, 3 + 204]
        # float64 products assigned into the float32 array, exactly like the reference's in-place slice assignments
        bj[:, 3:] = bj[:, 3:] * stats['Xstd_local'][0:-4] + stats['Xmean_local'][0:-4]
        bj[:, 0:2] = bj[:, 0:2] * stats['Xstd_global_xy'] + stats['Xmean_global_xy']
        bj[:, 2] = bj[:, 2] * stats['Xstd_global_r'] + stats['Xmean_global_r']
        bj = bj.reshape(T, -1, 3)
        bj64 = np.concatenate([np.zeros([T, 1, 3]), bj[:, 1:], bj[:, 0:1]], axis=1)      # float64 from here on
        glob = reconstruct_global_body(bj64, rot_0_pivot)[:, 1:, :]
        outs.append(glob.astype(np.float32))
    return outs[0], contact, outs[1]


def load_stats():
    import os
    t = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'lemo_b200', 'assets', 'lemo_tables.npz'))
    return {k[len('infill_'):]: t[k] for k in t.files if k.startswith('infill_')}
