"""Drop-in for Partsize-identical/models/pointnet_util.py on the B200 kernels.

Same public names, argument order, tensor layouts, int64 index outputs, error behaviour and
nn.Module parameter names (`mlp_convs.i`, `mlp_bns.i`, `conv_blocks.i.j`, `bn_blocks.i.j`) as
the reference, so its checkpoints load and its networks can import this module instead.
Behind the names: farthest_point_sample / query_ball_point / index_points / three-NN are the
CUDA kernels of libpcbridge (no distance matrices, no sorts), and the shared MLPs run on
point-major rows ([B*S*K, C] @ W^T) instead of permuted NCHW tensors.

Two additions that the reference only has inline (pointnet_util.py:325-334):
`three_nn` and `three_interpolate`.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

__all__ = [
    "square_distance", "index_points", "farthest_point_sample", "query_ball_point", "sample_and_group",
    "sample_and_group_all", "three_nn", "three_interpolate", "PointNetSetAbstraction",
    "PointNetSetAbstractionMsg", "PointNetFeaturePropagation",
]


# --------------------------------------------------------------------------------------------
# L1 primitives (reference: pointnet_util.py:22-174)
# --------------------------------------------------------------------------------------------
def square_distance(src, dst):
    """src [B,N,C], dst [B,M,C] -> [B,N,M]; -2*src.dst^T + |src|^2 + |dst|^2 in the reference's
    fp32 operation order (pointnet_util.py:22-43)."""
    return ops.square_distance(src, dst)


def index_points(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,C] / [B,S,K,C] (pointnet_util.py:46-63).
    An index outside [-N, N) is an IndexError in the reference; here it is recorded on the
    device and raised by ops.check_index_errors() (or at once with PCB_STRICT_INDEX=1)."""
    return ops.gather(points, idx, clamp=False)


def farthest_point_sample(xyz, npoint):
    """xyz [B,N,3] -> LongTensor [B,npoint] (pointnet_util.py:66-88).  The start index is drawn
    from the CPU default generator exactly as the reference does."""
    return ops.furthest_point_sample(xyz, npoint)


def query_ball_point(radius, nsample, xyz, new_xyz):
    """-> LongTensor [B,S,nsample] (pointnet_util.py:91-112)."""
    return ops.ball_query(radius, nsample, xyz, new_xyz)


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, pad_to=1):
    """FPS -> centroids -> ball query -> grouped [dxyz | features]  (pointnet_util.py:116-152).
    xyz [B,N,3], points [B,N,D] or None -> new_xyz [B,S,3], new_points [B,S,nsample,3+D]
    (pad_to > 1: channel count rounded up with zero columns, see ops.group_points)."""
    fps_idx = farthest_point_sample(xyz, npoint)
    new_xyz = index_points(xyz, fps_idx)
    idx = query_ball_point(radius, nsample, xyz, new_xyz)
    new_points = ops.group_points(xyz, points, new_xyz, idx, xyz_first=True, pad_to=pad_to)
    if returnfps:
        return new_xyz, new_points, index_points(xyz, idx), fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """One group holding every point, centroid at the origin (pointnet_util.py:157-174)."""
    B, N, d = xyz.shape
    new_xyz = torch.zeros(B, 1, d, device=xyz.device, dtype=xyz.dtype)
    grouped = xyz.view(B, 1, N, d)
    if points is not None:
        grouped = torch.cat([grouped, points.view(B, 1, N, -1)], dim=-1)
    return new_xyz, grouped


def three_nn(xyz1, xyz2, k=3):
    """(dist, idx) of the k nearest xyz2 [B,S,3] points for each xyz1 [B,N,3] point, ascending,
    ties by index -- what `square_distance(...).sort()[:, :, :k]` yields (pointnet_util.py:325-328)."""
    dist, idx, _ = ops.three_nn(xyz1, xyz2, k)
    return dist, idx


def three_interpolate(points2, idx, dist):
    """Inverse-distance interpolation of points2 [B,S,D] at the neighbours (idx, dist) [B,N,k]
    -> [B,N,D]  (pointnet_util.py:330-334)."""
    rec = 1.0 / (dist + 1e-8)
    weight = rec / torch.sum(rec, dim=2, keepdim=True)
    return ops.three_interpolate(points2, idx, weight, channels_first=False)


# --------------------------------------------------------------------------------------------
# shared MLP on rows
# --------------------------------------------------------------------------------------------
def _rows(t_bcn):
    """[B,C,N] (any strides) -> point-major [B,N,C] contiguous; free when the tensor already is a
    permuted view of point-major memory, which is what these modules hand to each other."""
    t = t_bcn.permute(0, 2, 1)
    return t if t.is_contiguous() else t.contiguous()


def _bn_rows(bn, x):
    """BatchNorm{1,2}d over every row of x [M,C]: the same statistics as the reference's
    BatchNorm2d over (B, K, S) / BatchNorm1d over (B, N)."""
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    use_batch = bn.training or not bn.track_running_stats
    mom = 0.0 if bn.momentum is None else bn.momentum
    return F.batch_norm(x, bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, bn.weight, bn.bias,
                        use_batch, mom, bn.eps)


def conv_bn_relu_rows(x, conv, bn, pool_k=1, out=None):
    """One shared-MLP layer on rows: 1x1 conv -> BN -> ReLU, optionally followed by the max over
    groups of `pool_k` consecutive rows (the neighbour axis).  In training mode the elementwise
    half runs in libpcbridge's fused row kernels (csrc/bn_rows.cu): the conv bias is folded into
    the normalisation and ReLU / max-pool happen in the same pass."""
    if ops.mlp_rows_fused_supported(x, [conv], [bn], pool_k):
        # bf16 training: tcgen05 GEMM with the BatchNorm statistics in its epilogue + one elementwise kernel
        return ops.mlp_rows_fused(x, [conv], [bn], pool_k, out)
    w = conv.weight.flatten(1)
    if ops._step_ctx is not None and x.is_cuda and x.shape[1] % 8 and bn.training:
        # inside a step runner: rows that are not a multiple of 8 channels wide (3 xyz / 9 / 259 concatenated
        # features in the BriStruNet modules) get zero pad columns once, so that the forward, dgrad and weight
        # gradient GEMMs all see 16-byte aligned rows (otherwise cuBLAS's unaligned legacy kernels)
        x = F.pad(x, (0, -x.shape[1] % 8))
    padded = x.shape[1] != w.shape[1]                     # zero pad columns from group_points(pad_to=8)
    if bn.training and x.is_cuda and bn.momentum is not None and bn.affine \
            and x.shape[0] % pool_k == 0 and 1 < pool_k + 1 <= 256 and x.shape[0] > 1:
        # bias-free: BN(xW + b) == BN(xW) + running-mean shift.  Output channels that are not a multiple
        # of 8 (196 in the MSG network) are carried as zero pad columns through BN into the next layer:
        # 392-byte rows would send the neighbouring GEMMs to cuBLAS's unaligned legacy kernels
        y = ops.linear_rows(x, w, pad_n=True)
        if ops.bn_rows_supported(y, bn, pool_k):
            return ops.bn_relu_rows(y, conv.bias, bn, relu=True, pool_k=pool_k, out=out)   # out: see mlp_rows
        y = y[:, :w.shape[0]]
        x = y if conv.bias is None else y + conv.bias
    else:
        x = F.linear(x, F.pad(w, (0, x.shape[1] - w.shape[1])) if padded else w, conv.bias)
    x = F.relu(_bn_rows(bn, x), inplace=True)
    if pool_k > 1:
        x = x.view(-1, pool_k, x.shape[-1]).max(dim=1)[0]
    return x


def mlp_rows(x, convs, bns, pool_k=1, out=None):
    """x [M,Cin] -> [M/pool_k,Cout]: (1x1 conv -> BN -> ReLU) per layer (pointnet_util.py:213-215), the
    last layer followed by the max over `pool_k` neighbours (:217).  `out`: optional [M/pool_k, Cout] column
    slice of a wider buffer that the last layer's fused kernel writes into when it can (training mode; the
    result then IS `out`, which the caller checks by data pointer)."""
    n = len(convs)
    if ops.mlp_rows_fused_supported(x, convs, bns, pool_k):
        x = ops.mlp_rows_fused(x, convs, bns, pool_k, out)        # the whole stack as one autograd node
    elif x.is_cuda and ops.fused_inference_enabled() and not any(bn.training for bn in bns):
        # evaluation under bf16 autocast: BatchNorm folded into the weights, one tcgen05 GEMM per layer with the
        # bias + ReLU (+ max over the neighbours) in its epilogue -- every layer width
        x = ops.mlp_rows_infer(x, ops.folded_mlp(convs[0], id(convs[-1]), convs, bns), pool_k)
    else:
        for i, (conv, bn) in enumerate(zip(convs, bns)):
            x = conv_bn_relu_rows(x, conv, bn, pool_k if i == n - 1 else 1, out if i == n - 1 else None)
    if x.shape[1] != convs[-1].weight.shape[0]:          # zero pad columns of the last layer are dropped
        x = x[:, :convs[-1].weight.shape[0]].contiguous()
    return x


def _cf_view(rows, B, S):
    """[B*S, C] rows -> the reference's channels-first [B,C,S] shape, as a view."""
    return rows.view(B, S, -1).permute(0, 2, 1)


# --------------------------------------------------------------------------------------------
# L2 modules (reference: pointnet_util.py:179-348)
# --------------------------------------------------------------------------------------------
class PointNetSetAbstraction(nn.Module):
    """pointnet_util.py:179-219.  forward(xyz [B,3,N], points [B,D,N] | None)
    -> new_xyz [B,3,S], new_points [B,mlp[-1],S]."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv2d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm2d(out))
            last = out

    def forward(self, xyz, points, pre=None):
        # pre: (new_xyz [B,S,3], [idx [B,S,K]]) computed ahead of the step (get_model.index_chain) or None
        xyz_r = _rows(xyz)
        pts_r = _rows(points) if points is not None else None
        B = xyz_r.shape[0]
        if pre is not None and not self.group_all:
            new_xyz, (idx,) = pre
            grouped = ops.group_points(xyz_r, pts_r, new_xyz, idx, xyz_first=True, pad_to=8)
            S, K, C = grouped.shape[1:]
            y = mlp_rows(grouped.reshape(B * S * K, C), self.mlp_convs, self.mlp_bns, pool_k=K)
            return new_xyz.permute(0, 2, 1), _cf_view(y, B, S)
        if not self.group_all and ops.fused_inference_enabled() and not self.mlp_bns[0].training:
            pk = ops.packed_mlp(self, 0, self.mlp_convs, self.mlp_bns, 3 + (pts_r.shape[2] if pts_r is not None else 0))
            if pk.ok and self.nsample <= 128:
                new_xyz = index_points(xyz_r, farthest_point_sample(xyz_r, self.npoint))
                idx = query_ball_point(self.radius, self.nsample, xyz_r, new_xyz)
                y = ops.sa_fused(xyz_r, pts_r, new_xyz, idx, pk, xyz_first=True)
                return new_xyz.permute(0, 2, 1), _cf_view(y, B, self.npoint)
        if self.group_all:
            new_xyz, grouped = sample_and_group_all(xyz_r, pts_r)
        else:
            new_xyz, grouped = sample_and_group(self.npoint, self.radius, self.nsample, xyz_r, pts_r, pad_to=8)
        S, K, C = grouped.shape[1:]
        y = mlp_rows(grouped.reshape(B * S * K, C), self.mlp_convs, self.mlp_bns, pool_k=K)   # max over neighbours
        return new_xyz.permute(0, 2, 1), _cf_view(y, B, S)


class PointNetSetAbstractionMsg(nn.Module):
    """pointnet_util.py:222-284: one FPS, then per radius ball query -> [features | dxyz] ->
    shared MLP -> max; scales concatenated on channels."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint, self.radius_list, self.nsample_list = npoint, radius_list, nsample_list
        self.conv_blocks = nn.ModuleList()
        self.bn_blocks = nn.ModuleList()
        for mlp in mlp_list:
            convs, bns = nn.ModuleList(), nn.ModuleList()
            last = in_channel + 3
            for out in mlp:
                convs.append(nn.Conv2d(last, out, 1))
                bns.append(nn.BatchNorm2d(out))
                last = out
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)

    def forward(self, xyz, points, pre=None):
        # pre: (new_xyz [B,S,3], [idx per radius]) computed ahead of the step (get_model.index_chain) or None
        xyz_r = _rows(xyz)
        pts_r = _rows(points) if points is not None else None
        B = xyz_r.shape[0]
        S = self.npoint
        outs, joined, forked = [], None, []
        if pre is not None:
            new_xyz, idxs = pre
        else:
            new_xyz = index_points(xyz_r, farthest_point_sample(xyz_r, S))
            # every radius in one scan of the cloud (the reference calls query_ball_point once per radius, :250)
            idxs = ops.ball_query_multi(self.radius_list, self.nsample_list, xyz_r, new_xyz)
        groupeds = None
        if self.training and xyz_r.is_cuda and pts_r is not None and pts_r.requires_grad:
            # all scales grouped by one autograd node: their scatter-add backward shares one gradient buffer
            groupeds = ops.group_points_multi(xyz_r, pts_r, new_xyz, idxs, xyz_first=False, pad_to=8)
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            idx = idxs[i]
            if ops.fused_inference_enabled() and not self.bn_blocks[i][0].training and K <= 128:
                pk = ops.packed_mlp(self, i, self.conv_blocks[i], self.bn_blocks[i],
                                    3 + (pts_r.shape[2] if pts_r is not None else 0))
                if pk.ok:
                    outs.append(ops.sa_fused(xyz_r, pts_r, new_xyz, idx, pk, xyz_first=False))
                    continue
            grouped = groupeds[i] if groupeds is not None else \
                ops.group_points(xyz_r, pts_r, new_xyz, idx, xyz_first=False, pad_to=8)          # [feat | dxyz | 0]
            if joined is None and self.training and grouped.is_cuda:
                # the scales write their pooled rows straight into the concatenated output (no torch.cat copy)
                widths = [blk[-1].weight.shape[0] for blk in self.conv_blocks]
                joined = torch.empty(B * S, sum(widths), dtype=grouped.dtype, device=grouped.device)
            off = sum(o.shape[1] for o in outs)
            slot = joined[:, off:off + self.conv_blocks[i][-1].weight.shape[0]] if joined is not None else None
            # inside a step runner the scales of a level run on their own streams (a fork / join of the captured graph):
            # they share nothing but their inputs, and the kernels of the small levels leave most SMs idle
            side = ops.scale_stream(grouped.device, i) if (i > 0 and self.training and grouped.is_cuda) else None
            if side is None:
                outs.append(mlp_rows(grouped.view(B * S * K, -1), self.conv_blocks[i], self.bn_blocks[i], pool_k=K, out=slot))
            else:
                main = torch.cuda.current_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    o = mlp_rows(grouped.view(B * S * K, -1), self.conv_blocks[i], self.bn_blocks[i], pool_k=K, out=slot)
                grouped.record_stream(side)
                o.record_stream(main)
                forked.append(side)
                outs.append(o)
        for side in forked:
            torch.cuda.current_stream().wait_stream(side)
        if joined is not None and all(o.data_ptr() == joined.data_ptr() + joined.element_size() * sum(p.shape[1] for p in outs[:j])
                                      and o.stride(0) == joined.stride(0) for j, o in enumerate(outs)):
            y = ops.join_columns(joined, outs)
        else:
            y = torch.cat(outs, dim=1)
        return new_xyz.permute(0, 2, 1), _cf_view(y, B, S)


class PointNetFeaturePropagation(nn.Module):
    """pointnet_util.py:287-348: three-NN inverse-distance interpolation of points2 onto xyz1,
    concat with points1, Conv1d/BN/ReLU stack.  All tensors channels-first as in the reference."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        last = in_channel
        for out in mlp:
            self.mlp_convs.append(nn.Conv1d(last, out, 1))
            self.mlp_bns.append(nn.BatchNorm1d(out))
            last = out

    def forward(self, xyz1, xyz2, points1, points2, pre=None):
        # pre: (idx [B,N,3], weight [B,N,3]) of the three nearest xyz2 points, computed ahead of the step, or None
        x1, x2, p2 = _rows(xyz1), _rows(xyz2), _rows(points2)
        B, N, _ = x1.shape
        S = x2.shape[1]
        if S == 1:
            interp = p2.repeat(1, N, 1)
        else:
            if pre is not None:
                idx, weight = pre
            else:
                _, idx, weight = ops.three_nn(x1, x2, 3)
            p1 = _rows(points1) if points1 is not None else None
            if self.mlp_bns[0].training and ops.fp_concat_supported(p1, p2):
                # bf16 training: interpolate + concat + cast in one pass, rows padded for the GEMM
                rows = ops.fp_concat(p1, p2, idx, weight, pad_to=8)
                y = mlp_rows(rows.view(B * N, -1), self.mlp_convs, self.mlp_bns)
                return _cf_view(y, B, N)
            interp = ops.three_interpolate(p2, idx, weight, channels_first=False)
        if points1 is not None:
            interp = torch.cat([_rows(points1).to(interp.dtype), interp], dim=-1)
        y = mlp_rows(interp.view(B * N, -1), self.mlp_convs, self.mlp_bns)
        return _cf_view(y, B, N)
