"""Drop-in for Highway_bridge/models/pointnet2_utils.py on the B200 kernels.

This flavour keeps coordinates point-major (`xyz [B,N,3]`) and features channels-first
(`[B,C,N]`), clamps gather indices into range instead of raising (pointnet2_utils.py:34-36) and
orders grouped channels [dxyz | features].  Same function names/signatures and the same module
parameter names (`mlp_convs`, `mlp_bns`, `conv_blocks`, `bn_blocks`, `attention`,
`boundary_aware`) as the reference; the unused AVSNet sampler (pointnet2_utils.py:363-485, its
hook is commented out at :83-95) is not part of the path and is not provided.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..partsize.pointnet_util import _bn_rows, _cf_view, _rows, conv_bn_relu_rows, mlp_rows

__all__ = [
    "square_distance", "index_points", "sample_and_group", "farthest_point_sample", "query_ball_point",
    "three_nn", "three_interpolate", "SetAbstraction", "FeaturePropagation", "EnhancedFeaturePropagation",
    "MultiScaleSetAbstraction", "seq_rows",
]


def square_distance(src, dst):
    """pointnet2_utils.py:7-14."""
    return ops.square_distance(src, dst)


def index_points(points, idx):
    """points [B,N,C], idx [B,S(,K)] -> [B,S(,K),C]; indices clamped to [0, N-1]
    (pointnet2_utils.py:17-39)."""
    return ops.gather(points, idx, clamp=True)


def farthest_point_sample(xyz, npoint):
    """pointnet2_utils.py:63-80."""
    return ops.furthest_point_sample(xyz, npoint)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """pointnet2_utils.py:97-112."""
    return ops.ball_query(radius, nsample, xyz, new_xyz)


def sample_and_group(npoint, radius, nsample, xyz, points, pad_to=1):
    """xyz [B,N,3], points [B,N,D] | None -> new_xyz [B,S,3], new_points [B,S,nsample,3+D]
    (pointnet2_utils.py:42-60).  pad_to > 1: channels rounded up with zero columns (ops.group_points)."""
    new_xyz = index_points(xyz, farthest_point_sample(xyz, npoint))
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    return new_xyz, ops.group_points(xyz, points, new_xyz, idx, xyz_first=True, clamp=True, pad_to=pad_to)


def three_nn(xyz1, xyz2, k=3):
    """k nearest xyz2 points per xyz1 point: what `square_distance().sort()[..., :k]` selects
    (pointnet2_utils.py:183-186; k=4 at :253-256)."""
    dist, idx, _ = ops.three_nn(xyz1, xyz2, k)
    return dist, idx


def three_interpolate(points2, idx, dist):
    """points2 [B,D,S] channels-first, (idx, dist) [B,N,k] -> [B,D,N]  (pointnet2_utils.py:189-196)."""
    rec = 1.0 / (dist + 1e-8)
    weight = rec / torch.sum(rec, dim=2, keepdim=True)
    return ops.three_interpolate(points2, idx, weight, channels_first=True)


def _seq_rows_plain(seq: nn.Sequential, x):
    """Layer-by-layer evaluation with library ops (the pre-fusion form; kept as the comparison path of the tests)."""
    for layer in seq:
        if isinstance(layer, (nn.Conv1d, nn.Conv2d)):
            x = F.linear(x, layer.weight.flatten(1), layer.bias)
        elif isinstance(layer, (nn.BatchNorm1d, nn.BatchNorm2d)):
            x = _bn_rows(layer, x)
        else:
            x = layer(x)
    return x


def seq_rows(seq: nn.Sequential, x):
    """Evaluate a Sequential of Conv1d(1x1)/BatchNorm1d/activations on point-major rows [M,C].
    In training mode every Conv -> BatchNorm -> ReLU triple runs through the fused row kernels
    (`conv_bn_relu_rows`: bias-free GEMM, one BN+ReLU kernel, row weight gradient), and inside a step runner
    the remaining convolutions use its bf16 weight shadows with output rows padded to 8 channels -- the 3-,
    6- and 9-channel layers of the BriStruNet encoders otherwise send M = 2 M-row GEMMs to cuBLAS's
    unaligned legacy kernels.  Zero pad columns are carried between layers and dropped at the end."""
    layers = list(seq)
    true_c = x.shape[1]
    i = 0
    while i < len(layers):
        layer = layers[i]
        if isinstance(layer, (nn.Conv1d, nn.Conv2d)):
            w = layer.weight.flatten(1)
            nxt = layers[i + 1] if i + 1 < len(layers) else None
            nxt2 = layers[i + 2] if i + 2 < len(layers) else None
            if isinstance(nxt, (nn.BatchNorm1d, nn.BatchNorm2d)) and type(nxt2) is nn.ReLU and nxt.training and x.is_cuda:
                x = conv_bn_relu_rows(x, layer, nxt)
                true_c = w.shape[0]
                i += 3
                continue
            if isinstance(nxt, (nn.BatchNorm1d, nn.BatchNorm2d)) and type(nxt2) is nn.ReLU and not nxt.training \
                    and x.is_cuda and x.dim() == 2 and ops.fused_inference_enabled():
                # evaluation under bf16 autocast: BatchNorm folded in, bias + ReLU in the GEMM epilogue
                x = ops.mlp_rows_infer(x, ops.folded_mlp(layer, id(nxt), [layer], [nxt]))
                true_c = w.shape[0]
                i += 3
                continue
            if ops._step_ctx is not None and x.is_cuda and ops._step_ctx.shadow(w) is not None:
                if x.shape[1] % 8:
                    x = F.pad(x, (0, -x.shape[1] % 8))                   # aligned rows for forward / dgrad / wgrad
                y = ops.linear_rows(x, w, pad_n=True)                    # [M, n8], pad columns zero
                if layer.bias is not None:
                    y = y + F.pad(layer.bias, (0, y.shape[1] - w.shape[0])).to(y.dtype)
                x = y
            else:
                if x.shape[1] != w.shape[1]:                             # zero pad columns from the previous layer
                    w = F.pad(w, (0, x.shape[1] - w.shape[1]))
                x = F.linear(x, w, layer.bias)
            true_c = layer.weight.shape[0]
        elif isinstance(layer, (nn.BatchNorm1d, nn.BatchNorm2d)):
            if x.shape[1] != true_c:
                x = x[:, :true_c].contiguous()
            x = _bn_rows(layer, x)
        elif isinstance(layer, nn.Linear):
            if x.shape[1] != true_c:
                x = x[:, :true_c]
            x = layer(x)
            true_c = x.shape[1]
        else:                                   # ReLU / LeakyReLU / Sigmoid / Dropout: pad columns may become non-zero,
            if x.shape[1] != true_c and not isinstance(layer, (nn.ReLU, nn.LeakyReLU, nn.Dropout)):
                x = x[:, :true_c].contiguous()  # drop them first unless the activation maps 0 -> 0
            x = layer(x)
        i += 1
    if x.shape[1] != true_c:
        x = x[:, :true_c].contiguous()
    return x


def _interpolate_rows(xyz1, xyz2, points2, k):
    """Inverse-distance interpolation of points2 [B,D,S] at xyz1, returned as rows [B,N,D]."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    p2 = _rows(points2)
    if S == 1:
        return p2.repeat(1, N, 1)
    _, idx, weight = ops.three_nn(xyz1, xyz2, k)
    return ops.three_interpolate(p2, idx, weight, channels_first=False)


class SetAbstraction(nn.Module):
    """pointnet2_utils.py:115-156.  forward(xyz [B,N,3], points [B,C,N] | None)
    -> new_xyz [B,S,3], new_points [B,mlp[-1],S]."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out))
            last = out

    def forward(self, xyz, points):
        xyz = xyz.contiguous()
        pts = _rows(points) if points is not None else None
        if ops.fused_inference_enabled() and not self.mlp_bns[0].training and self.nsample <= 128:
            pk = ops.packed_mlp(self, 0, self.mlp_convs, self.mlp_bns, 3 + (pts.shape[2] if pts is not None else 0))
            if pk.ok:
                new_xyz = index_points(xyz, farthest_point_sample(xyz, self.npoint))
                idx = query_ball_point(self.radius, self.nsample, xyz, new_xyz)
                y = ops.sa_fused(xyz, pts, new_xyz, idx, pk, xyz_first=True)
                return new_xyz, _cf_view(y, xyz.shape[0], self.npoint)
        new_xyz, grouped = sample_and_group(self.npoint, self.radius, self.nsample, xyz, pts, pad_to=8)
        B, S, K, C = grouped.shape
        y = mlp_rows(grouped.view(B * S * K, C), self.mlp_convs, self.mlp_bns, pool_k=K)
        return new_xyz, _cf_view(y, B, S)


class MultiScaleSetAbstraction(nn.Module):
    """pointnet2_utils.py:302-360: one FPS, per-radius grouping through its own MLP, scales
    concatenated on channels."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for _ in radius_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel
            for out in mlp:
                convs.append(nn.Conv2d(last, out, 1))
                bns.append(nn.BatchNorm2d(out))
                last = out
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points):
        xyz = xyz.contiguous()
        pts = _rows(points) if points is not None else None
        B = xyz.shape[0]
        S = self.npoint
        new_xyz = index_points(xyz, farthest_point_sample(xyz, S))
        outs, forked = [], []
        # every radius in one scan of the cloud (the reference calls query_ball_point once per radius)
        idxs = ops.ball_query_multi(self.radius_list, self.nsample_list, xyz, new_xyz)
        for i, (radius, K) in enumerate(zip(self.radius_list, self.nsample_list)):
            idx = idxs[i]
            if ops.fused_inference_enabled() and not self.bn_blocks[i][0].training and K <= 128:
                pk = ops.packed_mlp(self, i, self.conv_blocks[i], self.bn_blocks[i],
                                    3 + (pts.shape[2] if pts is not None else 0))
                if pk.ok:
                    outs.append(ops.sa_fused(xyz, pts, new_xyz, idx, pk, xyz_first=True))
                    continue
            # inside a step runner the scales run on their own streams (fork / join of the captured graph), grouping included
            side = ops.scale_stream(xyz.device, i) if (i > 0 and self.training and xyz.is_cuda) else None
            if side is None:
                grouped = ops.group_points(xyz, pts, new_xyz, idx, xyz_first=True, clamp=True, pad_to=8)
                outs.append(mlp_rows(grouped.view(B * S * K, -1), self.conv_blocks[i], self.bn_blocks[i], pool_k=K))
                continue
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grouped = ops.group_points(xyz, pts, new_xyz, idx, xyz_first=True, clamp=True, pad_to=8)
                o = mlp_rows(grouped.view(B * S * K, -1), self.conv_blocks[i], self.bn_blocks[i], pool_k=K)
            for t in (xyz, pts, new_xyz, idx):
                if t is not None:
                    t.record_stream(side)
            o.record_stream(main)
            forked.append(side)
            outs.append(o)
        for side in forked:
            torch.cuda.current_stream().wait_stream(side)
        return new_xyz, _cf_view(torch.cat(outs, dim=1), B, S)


class FeaturePropagation(nn.Module):
    """pointnet2_utils.py:159-211.  forward(xyz1 [B,N,3], xyz2 [B,S,3], points1 [B,D1,N] | None,
    points2 [B,D2,S]) -> [B,mlp[-1],N]."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out))
            last = out

    def forward(self, xyz1, xyz2, points1, points2):
        B, N, _ = xyz1.shape
        x = _interpolate_rows(xyz1.contiguous(), xyz2.contiguous(), points2, 3)
        if points1 is not None:
            x = torch.cat([_rows(points1).to(x.dtype), x], dim=-1)
        return _cf_view(mlp_rows(x.view(B * N, -1), self.mlp_convs, self.mlp_bns), B, N)


class EnhancedFeaturePropagation(nn.Module):
    """pointnet2_utils.py:214-299: 4-NN interpolation, squeeze-excite style channel gate,
    MLP (+ identity when widths match) and a coordinate-driven "boundary" branch."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        self.attention = nn.Sequential(
            nn.Conv1d(in_channel, in_channel // 4, 1), nn.BatchNorm1d(in_channel // 4), nn.ReLU(),
            nn.Conv1d(in_channel // 4, in_channel, 1), nn.Sigmoid())
        self.skip_connection = in_channel == mlp[-1]
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out))
            last = out
        self.boundary_aware = nn.Sequential(
            nn.Conv1d(3, 16, 1), nn.BatchNorm1d(16), nn.ReLU(), nn.Conv1d(16, mlp[-1], 1))

    def forward(self, xyz1, xyz2, points1, points2):
        B, N, _ = xyz1.shape
        xyz1 = xyz1.contiguous()
        x = _interpolate_rows(xyz1, xyz2.contiguous(), points2, 4)
        if points1 is not None:
            x = torch.cat([_rows(points1).to(x.dtype), x], dim=-1)
        x = x.view(B * N, -1)
        x = x * seq_rows(self.attention, x)
        edge = seq_rows(self.boundary_aware, xyz1.view(B * N, 3))
        y = mlp_rows(x, self.mlp_convs, self.mlp_bns)
        if self.skip_connection:
            y = y + x
        return _cf_view(y + edge, B, N)
