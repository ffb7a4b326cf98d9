"""CPU restatement of the PointNet++ SSG / MSG semantic-segmentation networks
(Partsize-identical/models/pointnet2_sem_seg.py, pointnet2_sem_seg_msg.py, pointnet_util.py:179-348)
on top of the C oracle's primitives.

TEST INFRASTRUCTURE ONLY (see oracle/pcb_oracle.c): used by tests/ as the checker and by
bench.py as the timed CPU baseline / --impl reference arm on the GPU box, where the reference
tree itself does not exist.  Index-producing steps (FPS, ball query, three-NN) run in the C
oracle; gathers, 1x1 convolutions, batch norm and pooling are plain PyTorch fp32 CPU ops in the
reference's tensor layouts, so autograd provides the backward pass.  Pinned against the
reference's own outputs in tests/test_oracle_models.py (tests/golden/models.npz).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import oracle as orc


def _take(points, idx):
    """points [B,N,C], idx [B,...] -> [B,...,C]   (index_points, pointnet_util.py:46-63)."""
    B = points.shape[0]
    b = torch.arange(B).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[b, idx]


def _fps(xyz, npoint):
    start = torch.randint(0, xyz.shape[1], (xyz.shape[0],), dtype=torch.long)     # pointnet_util.py:79
    return torch.from_numpy(orc.farthest_point_sample(xyz.detach().numpy(), npoint, start.numpy()))


def _ball(radius, k, xyz, new_xyz):
    return torch.from_numpy(orc.query_ball_point(radius, k, xyz.detach().numpy(), new_xyz.detach().numpy()))


def _shared_mlp(x, convs, bns):
    for conv, bn in zip(convs, bns):
        x = F.relu(bn(conv(x)))
    return x


def _stack(cin, widths, conv, norm):
    convs, bns = nn.ModuleList(), nn.ModuleList()
    for w in widths:
        convs.append(conv(cin, w, 1))
        bns.append(norm(w))
        cin = w
    return convs, bns


class SetAbstraction(nn.Module):                       # pointnet_util.py:179-219
    def __init__(self, npoint, radius, nsample, cin, mlp):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs, self.mlp_bns = _stack(cin, mlp, nn.Conv2d, nn.BatchNorm2d)

    def forward(self, xyz, points):
        xyz_t, pts_t = xyz.permute(0, 2, 1), points.permute(0, 2, 1)
        new_xyz = _take(xyz_t, _fps(xyz_t, self.npoint))
        idx = _ball(self.radius, self.nsample, xyz_t, new_xyz)
        g = torch.cat([_take(xyz_t, idx) - new_xyz.unsqueeze(2), _take(pts_t, idx)], dim=-1)   # [dxyz | feat]
        g = _shared_mlp(g.permute(0, 3, 2, 1), self.mlp_convs, self.mlp_bns)                  # [B,C,K,S]
        return new_xyz.permute(0, 2, 1), g.max(dim=2)[0]


class SetAbstractionMsg(nn.Module):                    # pointnet_util.py:222-284
    def __init__(self, npoint, radii, nsamples, cin, mlps):
        super().__init__()
        self.npoint, self.radii, self.nsamples = npoint, radii, nsamples
        self.conv_blocks, self.bn_blocks = nn.ModuleList(), nn.ModuleList()
        for mlp in mlps:
            c, b = _stack(cin + 3, mlp, nn.Conv2d, nn.BatchNorm2d)
            self.conv_blocks.append(c)
            self.bn_blocks.append(b)

    def forward(self, xyz, points):
        xyz_t, pts_t = xyz.permute(0, 2, 1), points.permute(0, 2, 1)
        new_xyz = _take(xyz_t, _fps(xyz_t, self.npoint))
        outs = []
        for i, (r, k) in enumerate(zip(self.radii, self.nsamples)):
            idx = _ball(r, k, xyz_t, new_xyz)
            g = torch.cat([_take(pts_t, idx), _take(xyz_t, idx) - new_xyz.unsqueeze(2)], dim=-1)  # [feat | dxyz]
            g = _shared_mlp(g.permute(0, 3, 2, 1), self.conv_blocks[i], self.bn_blocks[i])
            outs.append(g.max(dim=2)[0])
        return new_xyz.permute(0, 2, 1), torch.cat(outs, dim=1)


class FeaturePropagation(nn.Module):                   # pointnet_util.py:287-348
    def __init__(self, cin, mlp):
        super().__init__()
        self.mlp_convs, self.mlp_bns = _stack(cin, mlp, nn.Conv1d, nn.BatchNorm1d)

    def forward(self, xyz1, xyz2, points1, points2):
        x1, x2, p2 = xyz1.permute(0, 2, 1), xyz2.permute(0, 2, 1), points2.permute(0, 2, 1)
        N, S = x1.shape[1], x2.shape[1]
        if S == 1:
            interp = p2.repeat(1, N, 1)
        else:
            dist, idx = orc.three_nn(x1.detach().numpy(), x2.detach().numpy(), 3)
            dist, idx = torch.from_numpy(dist), torch.from_numpy(idx)
            rec = 1.0 / (dist + 1e-8)
            w = rec / rec.sum(dim=2, keepdim=True)
            interp = (_take(p2, idx) * w.unsqueeze(-1)).sum(dim=2)
        if points1 is not None:
            interp = torch.cat([points1.permute(0, 2, 1), interp], dim=-1)
        return _shared_mlp(interp.permute(0, 2, 1), self.mlp_convs, self.mlp_bns)


class _SemSeg(nn.Module):
    def _head(self, num_classes):
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)

    def forward(self, x):
        l0_xyz = x[:, :3, :]
        l1_xyz, l1 = self.sa1(l0_xyz, x)
        l2_xyz, l2 = self.sa2(l1_xyz, l1)
        l3_xyz, l3 = self.sa3(l2_xyz, l2)
        l4_xyz, l4 = self.sa4(l3_xyz, l3)
        l3 = self.fp4(l3_xyz, l4_xyz, l3, l4)
        l2 = self.fp3(l2_xyz, l3_xyz, l2, l3)
        l1 = self.fp2(l1_xyz, l2_xyz, l1, l2)
        l0 = self.fp1(l0_xyz, l1_xyz, None, l1)
        y = self.conv2(self.drop1(F.relu(self.bn1(self.conv1(l0)))))
        return F.log_softmax(y, dim=1).permute(0, 2, 1), l4


class PointNet2SSG(_SemSeg):                           # pointnet2_sem_seg.py:7-48
    def __init__(self, num_classes):
        super().__init__()
        self.sa1 = SetAbstraction(1024, 0.1, 32, 12, [32, 32, 64])
        self.sa2 = SetAbstraction(256, 0.2, 32, 67, [64, 64, 128])
        self.sa3 = SetAbstraction(64, 0.4, 32, 131, [128, 128, 256])
        self.sa4 = SetAbstraction(16, 0.8, 32, 259, [256, 256, 512])
        self.fp4 = FeaturePropagation(768, [256, 256])
        self.fp3 = FeaturePropagation(384, [256, 256])
        self.fp2 = FeaturePropagation(320, [256, 128])
        self.fp1 = FeaturePropagation(128, [128, 128, 128])
        self._head(num_classes)


class PointNet2MSG(_SemSeg):                           # pointnet2_sem_seg_msg.py:7-42
    def __init__(self, num_classes):
        super().__init__()
        self.sa1 = SetAbstractionMsg(1024, [0.05, 0.1], [16, 32], 9, [[16, 16, 32], [32, 32, 64]])
        self.sa2 = SetAbstractionMsg(256, [0.1, 0.2], [16, 32], 96, [[64, 64, 128], [64, 96, 128]])
        self.sa3 = SetAbstractionMsg(64, [0.2, 0.4], [16, 32], 256, [[128, 196, 256], [128, 196, 256]])
        self.sa4 = SetAbstractionMsg(16, [0.4, 0.8], [16, 32], 512, [[256, 256, 512], [256, 384, 512]])
        self.fp4 = FeaturePropagation(1536, [256, 256])
        self.fp3 = FeaturePropagation(512, [256, 256])
        self.fp2 = FeaturePropagation(352, [256, 128])
        self.fp1 = FeaturePropagation(128, [128, 128, 128])
        self._head(num_classes)


def msg_train_step(net, opt, x, labels):
    """One training step of BASELINE config 2 on CPU: forward, NLL loss, backward, Adam."""
    opt.zero_grad(set_to_none=True)
    logp, _ = net(x)
    loss = F.nll_loss(logp.reshape(-1, logp.shape[-1]), labels.reshape(-1))
    loss.backward()
    opt.step()
    return float(loss.detach())
