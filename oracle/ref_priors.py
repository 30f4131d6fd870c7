"""CPU restatement of the motion priors and Chamfer distance.  TEST INFRASTRUCTURE ONLY.

  Enc (smoothness prior)  /root/reference/models/AE_sep.py:11-30,77-99  (downsample=False, z_channel=64)
  AE  (infilling prior)   /root/reference/models/AE.py:11-108           (downsample=True, in_channel=4, kernel=3)
  chamferDist             /root/reference/temp_prox/dist_chamfer.py:10-53 around the external `chamfer`
                          CUDA extension (ChamferDistancePytorch@719b0f1, third-party, parity unpinned):
                          exact brute-force NN, squared L2, first minimum wins (strict <), int32 indices,
                          grad 2*g*(x1-x2) scattered to both clouds.
Networks are functional over a state_dict with the reference's key names, so the shipped .pkl
weights (exported to .npz) load unchanged.
"""
import torch
import torch.nn.functional as F


def enc_forward(x, sd):
    """Enc(downsample=False).forward -> z.  10x (conv3x3 p1 + LeakyReLU 0.2), no pooling."""
    h = x
    for blk in range(1, 6):
        for li in (0, 2):
            h = F.leaky_relu(F.conv2d(h, sd['enc_blc%d.main.%d.weight' % (blk, li)],
                                      sd['enc_blc%d.main.%d.bias' % (blk, li)], padding=1), 0.2)
    return h


def ae_forward(x, sd):
    """AE(downsample=True, in_channel=C, kernel=3).forward -> (rec, z)."""
    sizes = [x.shape]
    h = x
    for blk in range(1, 6):
        for li in (0, 2):
            h = F.leaky_relu(F.conv2d(h, sd['enc_blc%d.main.%d.weight' % (blk, li)],
                                      sd['enc_blc%d.main.%d.bias' % (blk, li)], padding=1), 0.2)
        h = F.max_pool2d(h, 3, 2, 1)
        sizes.append(h.shape)
    z = h
    for blk in range(1, 6):
        tgt = sizes[5 - blk]                                   # x_down4 .. x_down1, input
        w1, b1 = sd['dec_blc%d.deconv1.weight' % blk], sd['dec_blc%d.deconv1.bias' % blk]
        w2, b2 = sd['dec_blc%d.deconv2.weight' % blk], sd['dec_blc%d.deconv2.bias' % blk]
        # ConvTranspose2d(k3,s2,p1)(x, output_size=tgt): output_padding = tgt - ((in-1)*2 - 2 + 3)
        oph = tgt[2] - ((h.shape[2] - 1) * 2 + 1)
        opw = tgt[3] - ((h.shape[3] - 1) * 2 + 1)
        h = F.leaky_relu(F.conv_transpose2d(h, w1, b1, stride=2, padding=1, output_padding=(oph, opw)), 0.2)
        h = F.conv_transpose2d(h, w2, b2, stride=1, padding=1)
        if blk < 5:
            h = F.leaky_relu(h, 0.2)
    return h, z


def chamfer(xyz1, xyz2):
    """-> dist1 [B,n], dist2 [B,m] (squared), idx1, idx2 (int32).  Differentiable via gather."""
    d = ((xyz1[:, :, None, :] - xyz2[:, None, :, :]) ** 2).sum(-1)      # [B,n,m]
    i1 = d.argmin(2)
    i2 = d.argmin(1)
    n1 = torch.gather(xyz2, 1, i1[:, :, None].expand(-1, -1, 3))
    n2 = torch.gather(xyz1, 1, i2[:, :, None].expand(-1, -1, 3))
    return ((xyz1 - n1) ** 2).sum(-1), ((xyz2 - n2) ** 2).sum(-1), i1.int(), i2.int()
