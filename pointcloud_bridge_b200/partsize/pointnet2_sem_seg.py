"""PointNet++ single-scale-grouping semantic segmentation network, drop-in for
Partsize-identical/models/pointnet2_sem_seg.py (same `get_model(num_classes)` / `get_loss()`,
same parameter names, same [B,9,N] -> ([B,N,num_classes] log-probabilities, l4_points) contract).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .pointnet_util import (PointNetFeaturePropagation, PointNetSetAbstraction, _rows, conv_bn_relu_rows,
                            farthest_point_sample, index_points, query_ball_point)


class get_model(nn.Module):
    # layer table: pointnet2_sem_seg.py:11-22
    SA = [(1024, 0.1, 32, 9 + 3, [32, 32, 64]), (256, 0.2, 32, 64 + 3, [64, 64, 128]),
          (64, 0.4, 32, 128 + 3, [128, 128, 256]), (16, 0.8, 32, 256 + 3, [256, 256, 512])]
    FP = [(768, [256, 256]), (384, [256, 256]), (320, [256, 128]), (128, [128, 128, 128])]

    def __init__(self, num_classes):
        super().__init__()
        for i, (npoint, radius, nsample, cin, mlp) in enumerate(self.SA, 1):
            setattr(self, f"sa{i}", PointNetSetAbstraction(npoint, radius, nsample, cin, mlp, False))
        for i, (cin, mlp) in zip((4, 3, 2, 1), self.FP):
            setattr(self, f"fp{i}", PointNetFeaturePropagation(cin, mlp))
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.conv2 = nn.Conv1d(128, num_classes, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)

    def forward(self, xyz, pre=None):
        # pre: the dictionary of `index_chain(xyz)` (indices computed ahead of the step) or None
        sa = pre["sa"] if pre is not None else (None,) * 4
        fp = pre["fp"] if pre is not None else (None,) * 4
        l0_points, l0_xyz = xyz, xyz[:, :3, :]
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points, sa[0])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, sa[1])
        ops.mark_mid_step()                                  # the small layers begin (engine.Trainer.prefetch)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points, sa[2])
        l4_xyz, l4_points = self.sa4(l3_xyz, l3_points, sa[3])
        l3_points = self.fp4(l3_xyz, l4_xyz, l3_points, l4_points, fp[0])
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points, fp[1])
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points, fp[2])
        l0_points = self.fp1(l0_xyz, l1_xyz, None, l1_points, fp[3])
        return _seg_head(self, l0_points), l4_points

    @staticmethod
    def _ball_indices(m, xyz_r, new_xyz):
        return [query_ball_point(m.radius, m.nsample, xyz_r, new_xyz)]

    @torch.no_grad()
    def index_chain(self, xyz):
        """Every sampling / grouping / interpolation index of one forward pass.  They are functions of the input
        coordinates only (each level's centroids are exact gathers of the level below), so a step runner can compute
        them for the NEXT batch on a side stream while the current batch trains (engine.Trainer.prefetch) -- farthest
        point sampling is a serial chain that keeps 16 of 148 SMs busy for a tenth of the step.  Same calls, same order,
        same CPU-generator draws for the FPS start indices as the forward pass itself (pointnet_util.py:66-112,
        316-328).  xyz: the network input [B, 9, N].  Returns {"sa": [(new_xyz, [idx...])] * 4, "fp": [(idx, weight)] * 4}
        (fp in the order fp4, fp3, fp2, fp1) for `forward(xyz, pre=...)`."""
        cur = _rows(xyz[:, :3, :])
        levels, sa, fp = [cur], [], []
        for i in (1, 2, 3, 4):
            m = getattr(self, f"sa{i}")
            new_xyz = index_points(cur, farthest_point_sample(cur, m.npoint))
            sa.append((new_xyz, self._ball_indices(m, cur, new_xyz)))
            levels.append(new_xyz)
            cur = new_xyz
        for lvl in (3, 2, 1, 0):
            _, idx, weight = ops.three_nn(levels[lvl], levels[lvl + 1], 3)
            fp.append((idx, weight))
        return {"sa": sa, "fp": fp}


def _seg_head(net, feats_bcn):
    """conv1 -> bn1 -> ReLU -> dropout -> conv2 -> log_softmax over classes -> [B,N,classes]
    (pointnet2_sem_seg.py:43-47), evaluated on point-major rows."""
    x = _rows(feats_bcn)
    B, N, C = x.shape
    if x.is_cuda and ops.fused_inference_enabled() and not net.bn1.training and not net.drop1.training:
        # evaluation under bf16 autocast: conv1 + bn1 folded, ReLU in the epilogue; conv2 + bias on the same kernel
        h = ops.mlp_rows_infer(x.reshape(B * N, C), ops.folded_mlp(net.conv1, 0, [net.conv1], [net.bn1]))
        nc = net.conv2.weight.shape[0]
        lg = ops.mlp_rows_infer(h, ops.folded_mlp(net.conv2, 0, [net.conv2], [None]), last_act=False)[:, :nc]
        return F.log_softmax(lg.float(), dim=-1).view(B, N, -1)
    x = net.drop1(conv_bn_relu_rows(x.reshape(B * N, C), net.conv1, net.bn1))
    w, b = net.conv2.weight.flatten(1), net.conv2.bias
    nc = w.shape[0]
    if ops.head_logits_enabled() and x.is_cuda and nc <= 32:
        # training runner: bias-free classifier rows (zero-padded to 8 columns when a bf16 shadow exists) go
        # straight into the loss kernel, which adds the bias (ops.nll_logit_rows)
        return ops.LogitRows(ops.linear_rows(x, w, pad_n=True), b, nc, B, N)
    if x.is_cuda and x.dtype == torch.bfloat16 and ops.own_gemm_supported(x):
        # bf16 rows: the classifier on the tcgen05 row GEMM as well (class columns zero-padded to 8), bias added on the rows
        x = ops.linear_rows(x, w, pad_n=True)[:, :nc]
        x = x if b is None else x + b.to(x.dtype)
        return F.log_softmax(x.float(), dim=-1).view(B, N, -1)
    pad = (-nc) % 8 if x.is_cuda else 0
    if pad:                        # [M,5] bf16 rows are not 16-byte aligned: cuBLAS falls back to legacy kernels
        x = F.linear(x, F.pad(w, (0, 0, 0, pad)), F.pad(b, (0, pad)) if b is not None else None)[:, :nc]
    else:
        x = F.linear(x, w, b)
    return F.log_softmax(x.float(), dim=-1).view(B, N, -1)


class get_loss(nn.Module):
    """pointnet2_sem_seg.py:51-57."""

    def forward(self, pred, target, trans_feat, weight):
        return F.nll_loss(pred, target, weight=weight)
