"""CPU (PyTorch, fp32/fp64, autograd) restatement of the reference's body-model path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Followed sources (all under /root/reference unless marked third-party):
  lbs core            human_body_prior/body_model/lbs.py:34-263        (pinned by tests/golden/lbs_*.npz)
  SMPL-X wrapper      smplx==0.1.26 SMPLX.forward [third-party, parity unpinned]; concat order as
                      human_body_prior/body_model/body_model.py:230
  6D <-> aa           utils/utils.py:50-137
  tgm conversions     torchgeometry==0.1.2 [third-party, parity unpinned]
  VPoser decode       human_body_prior/train/vposer_smpl.py:49-62,107-121,152-161
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# rotation conversions
# ----------------------------------------------------------------------------------------------
def rodrigues(aa):
    """lbs.py:166-193.  aa [N,3] -> R [N,3,3].  Note the 1e-8 added to every component (lbs.py:178)."""
    n = aa.shape[0]
    theta = torch.norm(aa + 1e-8, dim=1, keepdim=True)
    u = aa / theta
    c, s = torch.cos(theta)[:, :, None], torch.sin(theta)[:, :, None]
    ux, uy, uz = u[:, 0:1], u[:, 1:2], u[:, 2:3]
    o = torch.zeros_like(ux)
    K = torch.cat([o, -uz, uy, uz, o, -ux, -uy, ux, o], 1).view(n, 3, 3)
    eye = torch.eye(3, dtype=aa.dtype).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def gram_schmidt_6d(x6):
    """utils/utils.py:64-70 (ContinousRotReprDecoder.decode).  [N,6] viewed (3,2) -> R [N,3,3]."""
    m = x6.reshape(-1, 3, 2)
    b1 = F.normalize(m[:, :, 0], dim=1)
    d = (b1 * m[:, :, 1]).sum(1, keepdim=True)
    b2 = F.normalize(m[:, :, 1] - d * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], -1)


def tgm_rotmat_to_quat(R34, eps=1e-6):
    """torchgeometry 0.1.2 rotation_matrix_to_quaternion (third-party, recalled; `~mask` patch applied).
    Input [N,3,4]; works on the transposed matrix; output (w,x,y,z)."""
    t = R34.transpose(1, 2)
    m_d2 = t[:, 2, 2] < eps
    m_d0_d1 = t[:, 0, 0] > t[:, 1, 1]
    m_d0_nd1 = t[:, 0, 0] < -t[:, 1, 1]
    t0 = 1 + t[:, 0, 0] - t[:, 1, 1] - t[:, 2, 2]
    q0 = torch.stack([t[:, 1, 2] - t[:, 2, 1], t0, t[:, 0, 1] + t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2]], -1)
    t1 = 1 - t[:, 0, 0] + t[:, 1, 1] - t[:, 2, 2]
    q1 = torch.stack([t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] + t[:, 1, 0], t1, t[:, 1, 2] + t[:, 2, 1]], -1)
    t2 = 1 - t[:, 0, 0] - t[:, 1, 1] + t[:, 2, 2]
    q2 = torch.stack([t[:, 0, 1] - t[:, 1, 0], t[:, 2, 0] + t[:, 0, 2], t[:, 1, 2] + t[:, 2, 1], t2], -1)
    t3 = 1 + t[:, 0, 0] + t[:, 1, 1] + t[:, 2, 2]
    q3 = torch.stack([t3, t[:, 1, 2] - t[:, 2, 1], t[:, 2, 0] - t[:, 0, 2], t[:, 0, 1] - t[:, 1, 0]], -1)
    c0 = (m_d2 & m_d0_d1).to(t.dtype)[:, None]
    c1 = (m_d2 & ~m_d0_d1).to(t.dtype)[:, None]
    c2 = (~m_d2 & m_d0_nd1).to(t.dtype)[:, None]
    c3 = (~m_d2 & ~m_d0_nd1).to(t.dtype)[:, None]
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0[:, None] * c0 + t1[:, None] * c1 + t2[:, None] * c2 + t3[:, None] * c3)
    return q * 0.5


def tgm_quat_to_aa(q):
    """torchgeometry 0.1.2 quaternion_to_angle_axis (third-party, recalled)."""
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = torch.sqrt(s2)
    c = q[..., 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    return torch.stack([q1 * k, q2 * k, q3 * k], -1)


def rotmat_to_aa(R):
    """utils/utils.py:74-81 / vposer_smpl.py:152-161: pad to 3x4, tgm rotation_matrix_to_angle_axis."""
    return tgm_quat_to_aa(tgm_rotmat_to_quat(F.pad(R.reshape(-1, 3, 3), [0, 1])))


def tgm_aa_to_rotmat(aa, eps=1e-6):
    """torchgeometry 0.1.2 angle_axis_to_rotation_matrix [:, :3, :3] (utils/utils.py:84-90; init only)."""
    th2 = (aa * aa).sum(1, keepdim=True)
    th = torch.sqrt(th2)
    w = aa / (th + eps)
    wx, wy, wz = w[:, 0:1], w[:, 1:2], w[:, 2:3]
    c, s = torch.cos(th), torch.sin(th)
    Rn = torch.cat([c + wx * wx * (1 - c), wx * wy * (1 - c) - wz * s, wy * s + wx * wz * (1 - c),
                    wz * s + wx * wy * (1 - c), c + wy * wy * (1 - c), -wx * s + wy * wz * (1 - c),
                    -wy * s + wx * wz * (1 - c), wx * s + wy * wz * (1 - c), c + wz * wz * (1 - c)], 1)
    rx, ry, rz = aa[:, 0:1], aa[:, 1:2], aa[:, 2:3]
    one = torch.ones_like(rx)
    Rt = torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], 1)
    m = (th2 > eps).to(aa.dtype)
    return (m * Rn + (1 - m) * Rt).view(-1, 3, 3)


def convert_to_6D_all(aa):
    """utils/utils.py:127-130: aa [N,3] -> first two columns of R, row-major [r00,r01,r10,r11,r20,r21]."""
    return tgm_aa_to_rotmat(aa)[:, :, :2].reshape(-1, 6)


def convert_to_3D_rot(x75):
    """utils/utils.py:111-123: [B,75]=(transl3, rot6d6, rest66) -> [B,72] with aa global orient."""
    R = gram_schmidt_6d(x75[:, 3:9])
    return torch.cat([x75[:, :3], rotmat_to_aa(R), x75[:, 9:]], -1)


# ----------------------------------------------------------------------------------------------
# LBS core  (lbs.py:34-263)
# ----------------------------------------------------------------------------------------------
def rigid_chain(R, Jrest, parents):
    """lbs.py:196-263.  R [B,J,3,3], Jrest [B,J,3] -> posed joints [B,J,3], rel transforms A [B,J,4,4]."""
    B, NJ = R.shape[:2]
    rel = Jrest.clone()
    rel[:, 1:] = Jrest[:, 1:] - Jrest[:, parents[1:]]
    M = torch.zeros(B, NJ, 4, 4, dtype=R.dtype)
    M[:, :, :3, :3] = R
    M[:, :, :3, 3] = rel
    M[:, :, 3, 3] = 1
    chain = [M[:, 0]]
    for i in range(1, NJ):
        chain.append(chain[int(parents[i])] @ M[:, i])
    G = torch.stack(chain, 1)
    posed = G[:, :, :3, 3]
    corr = torch.zeros_like(G)
    corr[:, :, :3, 3] = torch.einsum('bjik,bjk->bji', G[:, :, :3, :3], Jrest)
    return posed, G - corr


def lbs(betas, pose_aa, m):
    """lbs.py:34-119.  betas [B,NB], pose_aa [B,J*3]; m = dict of torch tensors.
    Returns verts [B,V,3], posed joints [B,J,3], plus v_posed for diagnostics."""
    B = betas.shape[0]
    v_shaped = m['v_template'][None] + torch.einsum('bl,mkl->bmk', betas, m['shapedirs'])
    Jrest = torch.einsum('bik,ji->bjk', v_shaped, m['J_regressor']).contiguous()
    R = rodrigues(pose_aa.reshape(-1, 3)).view(B, -1, 3, 3)
    feat = (R[:, 1:] - torch.eye(3, dtype=R.dtype)).reshape(B, -1)
    v_posed = v_shaped + (feat @ m['posedirs']).view(B, -1, 3)
    posed, A = rigid_chain(R, Jrest, m['parents'])
    # lbs.py:106-111 repeats W B times and bmm's; same contraction without the B*V*J temporary
    T = torch.einsum('vj,bjk->bvk', m['lbs_weights'], A.reshape(B, -1, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=R.dtype)], 2)
    verts = (T @ vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, posed, v_posed


def model_to_torch(m, dtype=torch.float32):
    out = {}
    for k, v in m.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        out[k] = t.to(dtype) if t.is_floating_point() else t.long()
    return out


class SMPLXRef:
    """smplx==0.1.26 SMPLX.forward semantics (third-party, parity unpinned).
    use_pca=True, num_pca_comps=12, flat_hand_mean=False, no joint_mapper, no face contour."""

    def __init__(self, model_np, num_pca_comps=12, dtype=torch.float32):
        self.m = model_to_torch(model_np, dtype)
        self.npc = num_pca_comps
        self.dtype = dtype
        m = self.m
        self.lh_comp = m['hands_componentsl'][:num_pca_comps]
        self.rh_comp = m['hands_componentsr'][:num_pca_comps]
        pm = torch.zeros(165, dtype=dtype)
        pm[75:120] = m['hands_meanl']
        pm[120:165] = m['hands_meanr']
        self.pose_mean = pm

    def full_pose(self, global_orient, body_pose, left_hand_pose, right_hand_pose,
                  jaw_pose=None, leye_pose=None, reye_pose=None):
        B = global_orient.shape[0]
        z3 = torch.zeros(B, 3, dtype=self.dtype)
        jaw = z3 if jaw_pose is None else jaw_pose
        le = z3 if leye_pose is None else leye_pose
        re = z3 if reye_pose is None else reye_pose
        lh = left_hand_pose @ self.lh_comp
        rh = right_hand_pose @ self.rh_comp
        return torch.cat([global_orient, body_pose, jaw, le, re, lh, rh], 1) + self.pose_mean

    def __call__(self, transl, global_orient, betas, body_pose, left_hand_pose, right_hand_pose,
                 expression=None, jaw_pose=None, leye_pose=None, reye_pose=None):
        B = global_orient.shape[0]
        m = self.m
        if expression is None:
            expression = torch.zeros(B, 10, dtype=self.dtype)
        fp = self.full_pose(global_orient, body_pose, left_hand_pose, right_hand_pose, jaw_pose, leye_pose, reye_pose)
        verts, joints, _ = lbs(torch.cat([betas, expression], 1), fp, m)
        extra = verts[:, m['extra_joint_vids']]
        tri = m['faces'][m['lmk_faces_idx']]                       # [51,3]
        lmk = torch.einsum('blkd,lk->bld', verts[:, tri], m['lmk_bary_coords'])
        joints = torch.cat([joints, extra, lmk], 1)
        return verts + transl[:, None], joints + transl[:, None], fp


# ----------------------------------------------------------------------------------------------
# VPoser decode (vposer_smpl.py:107-121), eval mode (dropout = identity)
# ----------------------------------------------------------------------------------------------
class VPoserRef:
    def __init__(self, w, dtype=torch.float32):
        self.w = {k: torch.from_numpy(v).to(dtype) for k, v in w.items()}

    def decode_matrot(self, z):
        w = self.w
        x = F.leaky_relu(F.linear(z, w['dec_fc1_w'], w['dec_fc1_b']), 0.2)
        x = F.leaky_relu(F.linear(x, w['dec_fc2_w'], w['dec_fc2_b']), 0.2)
        x = F.linear(x, w['dec_out_w'], w['dec_out_b'])
        return gram_schmidt_6d(x)                                   # [B*21,3,3]

    def decode_aa(self, z):
        """decode(Z, 'aa') -> [B,1,21,3]."""
        return rotmat_to_aa(self.decode_matrot(z)).view(z.shape[0], 1, 21, 3)


def gen_body_mesh(params72, smplx_ref, vposer_ref):
    """utils/utils.py:141-154 (and :156-169 for joints): params72 -> (verts, joints)."""
    B = params72.shape[0]
    body_pose = vposer_ref.decode_aa(params72[:, 16:48]).view(B, -1)
    v, j, _ = smplx_ref(transl=params72[:, 0:3], global_orient=params72[:, 3:6], betas=params72[:, 6:16],
                        body_pose=body_pose, left_hand_pose=params72[:, 48:60], right_hand_pose=params72[:, 60:72])
    return v, j
