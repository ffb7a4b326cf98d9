"""Networks of Highway_bridge/models/model.py on the B200 kernels: `PointNet2` (:12-56),
`EnhancedPointNet2` = BriStruNet / "BridgeSeg" (:58-147), `MultiScaleFeatureFusion` (:149-166)
and `BridgeStructureLoss` (:169-263).  Constructor arguments, forward signatures
(xyz [B,N,3], colours [B,N,3]) -> logits [B,num_classes,N], and parameter names equal the
reference's, so `load_state_dict(checkpoint['model_state_dict'])` works
(Highway_bridge/inference.py:108-110).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..partsize.pointnet_util import _cf_view, _rows
from .attention_modules import (BridgeStructureEncoding, ColorFeatureExtraction, CompositeFeatureFusion,
                                GeometricFeatureExtraction)
from .pointnet2_utils import (EnhancedFeaturePropagation, FeaturePropagation, MultiScaleSetAbstraction,
                              SetAbstraction, seq_rows)


class PointNet2(nn.Module):
    def __init__(self, num_classes=8):
        super().__init__()
        self.sa1 = SetAbstraction(1024, 0.1, 32, 6, [64, 64, 128])
        self.sa2 = SetAbstraction(256, 0.2, 32, 131, [128, 128, 256])
        self.sa3 = SetAbstraction(64, 0.4, 32, 259, [256, 256, 512])
        self.fp3 = FeaturePropagation(768, [256, 256])
        self.fp2 = FeaturePropagation(384, [256, 128])
        self.fp1 = FeaturePropagation(128, [128, 128, 128])
        self.conv1 = nn.Conv1d(128, 128, 1)
        self.bn1 = nn.BatchNorm1d(128)
        self.drop1 = nn.Dropout(0.5)
        self.conv2 = nn.Conv1d(128, num_classes, 1)

    def forward(self, xyz, points):
        points = points.transpose(1, 2)
        l1_xyz, l1 = self.sa1(xyz, points)
        l2_xyz, l2 = self.sa2(l1_xyz, l1)
        l3_xyz, l3 = self.sa3(l2_xyz, l2)
        l2 = self.fp3(l2_xyz, l3_xyz, l2, l3)
        l1 = self.fp2(l1_xyz, l2_xyz, l1, l2)
        l0 = self.fp1(xyz, l1_xyz, None, l1)
        x = _rows(l0)
        B, N, C = x.shape
        head = nn.Sequential(self.conv1, self.bn1, nn.ReLU(), self.drop1, self.conv2)
        return _cf_view(seq_rows(head, x.reshape(B * N, C)), B, N)


class MultiScaleFeatureFusion(nn.Module):
    def __init__(self, in_channels_list, out_channels):
        super().__init__()
        self.convs = nn.ModuleList(
            nn.Sequential(nn.Conv1d(c, out_channels, 1), nn.BatchNorm1d(out_channels), nn.ReLU())
            for c in in_channels_list)

    def forward(self, features_list):
        n = features_list[2].shape[2]
        outs = []
        for feat, conv in zip(features_list, self.convs):
            rows = _rows(feat)                                         # [B,S,C]
            B, S, C = rows.shape
            if S != n:                                                 # F.interpolate(..., size=n), nearest
                src = torch.floor(torch.arange(n, device=rows.device, dtype=torch.float32) * (S / n)).long()
                rows = rows[:, src.clamp_(max=S - 1)]
            outs.append(seq_rows(conv, rows.reshape(B * n, C)))
        return _cf_view(torch.cat(outs, dim=1), features_list[2].shape[0], n)


class EnhancedPointNet2(nn.Module):
    def __init__(self, num_classes=5):
        super().__init__()
        input_ch = 3
        self.bri_enc = BridgeStructureEncoding(input_ch, 32, 4)
        self.color_encoder = ColorFeatureExtraction(3, 6)
        self.feature_fusion = CompositeFeatureFusion(input_ch, 6)
        self.sa1 = MultiScaleSetAbstraction(1024, [0.1, 0.2], [16, 32], input_ch + 3, [64, 64, 128])
        self.sa2 = MultiScaleSetAbstraction(512, [0.2, 0.4], [16, 32], 259, [128, 128, 256])
        self.sa3 = MultiScaleSetAbstraction(128, [0.4, 0.8], [16, 32], 515, [256, 256, 512])
        self.geometric1 = GeometricFeatureExtraction(128 * 2)          # constructed but unused (model.py:129)
        self.geometric2 = GeometricFeatureExtraction(256 * 2)
        self.geometric3 = GeometricFeatureExtraction(512 * 2)
        self.fp3 = EnhancedFeaturePropagation(1536, [1024, 256])
        self.fp2 = EnhancedFeaturePropagation(512, [256, 256])
        self.fp1 = EnhancedFeaturePropagation(256 + input_ch, [256, 128])
        self.fusion = MultiScaleFeatureFusion(in_channels_list=[256, 256, 128], out_channels=128)
        self.final_fusion = nn.Sequential(nn.Conv1d(384, 128, 1), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.5),
                                          nn.Conv1d(128, num_classes, 1))
        self.num_classes = num_classes
        self.cls_head = nn.Sequential(                                   # never called (model.py:101-112)
            nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.ReLU(inplace=True), nn.Dropout(0.5),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.ReLU(inplace=True), nn.Dropout(0.5),
            nn.Linear(256, num_classes))

    def forward(self, xyz, features=None):
        pos = self.bri_enc(xyz)
        col = self.color_encoder(features.transpose(1, 2), xyz)
        fused = self.feature_fusion(pos, col)
        l1_xyz, l1 = self.sa1(xyz, fused)
        l2_xyz, l2 = self.sa2(l1_xyz, l1)
        l2 = self.geometric2(l2, l2_xyz)
        l3_xyz, l3 = self.sa3(l2_xyz, l2)
        l3 = self.geometric3(l3, l3_xyz)
        l2 = self.fp3(l2_xyz, l3_xyz, l2, l3)
        l1 = self.fp2(l1_xyz, l2_xyz, l1, l2)
        l0 = self.fp1(xyz, l1_xyz, fused, l1)
        mix = _rows(self.fusion([l2, l1, l0]))
        B, N, C = mix.shape
        return _cf_view(seq_rows(self.final_fusion, mix.reshape(B * N, C)), B, N)


def _weighted_smoothed_ce(x, target, weight, label_smoothing):
    """F.cross_entropy(x, target, weight=weight, label_smoothing=ls) with mean reduction, written out with
    fixed-shape ops: ATen's version divides by `weight.gather(0, target.masked_select(~ignore_mask)).sum()`,
    and masked_select synchronises with the host (not CUDA-graph capturable).  Same arithmetic:
    (1 - ls) * sum_i w[y_i] * (-logp[i, y_i]) / W  +  ls / C * sum_i sum_c w[c] * (-logp[i, c]) / W,  W = sum_i w[y_i]."""
    logp = F.log_softmax(x, dim=1)
    wy = weight[target]
    den = wy.sum()
    nll = -(wy * logp.gather(1, target.view(-1, 1)).squeeze(1)).sum() / den
    smooth = -(logp * weight.view(1, -1)).sum() / den
    return (1 - label_smoothing) * nll + smooth * (label_smoothing / x.shape[1])


class BridgeStructureLoss(nn.Module):
    """Label-smoothed cross entropy whose class weights grow when the predicted classes violate
    the vertical ordering abutment < girder < deck < parapet (model.py:169-263).  Plain PyTorch,
    not on the hot path; kept so that the BriStruNet training step is the reference's."""

    ORDER = {1: dict(below=[2, 3, 4]), 2: dict(above=[1], below=[3, 4]), 3: dict(above=[1, 2], below=[4]),
             4: dict(above=[1, 2, 3])}

    def __init__(self, num_classes=5, alpha=20.0, rel_margin=0.2, class_weights=None):
        super().__init__()
        self.alpha, self.rel_margin = alpha, rel_margin
        base = torch.tensor([1.5, 1.0, 1.2, 1.5, 1.0]) if class_weights is None else class_weights
        self.base_weights = base
        self.register_buffer("base_weights_buffer", base)

    @staticmethod
    def _mean_height(points, mask):
        m = mask.unsqueeze(-1)
        p = points * m
        lo, hi = p.amin(dim=1, keepdim=True), p.amax(dim=1, keepdim=True)
        rel = (p - lo) / (hi - lo + 1e-7)
        return (rel[..., 2] * mask).sum(dim=1) / mask.sum(dim=1).clamp(min=1)

    def forward(self, outputs, labels, points):
        logits = outputs.transpose(1, 2)
        B = labels.shape[0]
        preds = logits.argmax(dim=-1)
        w = self.base_weights_buffer.repeat(B, 1).to(logits.device)
        # The reference branches in Python on `present` / `m.any()` (model.py:204-230), i.e. one host sync per
        # class; the same arithmetic with 0/1 factors instead: a class that is absent contributes +0, and
        # _mean_height of an empty mask is exactly 0 (p = 0, rel = 0 / 1e-7, divided by clamp(0, 1) = 1).
        present = {c: (labels == c).any().to(w.dtype) for c in (1, 2, 3, 4)}
        height = {c: self._mean_height(points, preds == c) for c in (1, 2, 3, 4)}
        for c, rel in self.ORDER.items():
            for lower in rel.get("above", []):
                v = F.relu(-(height[c] - height[lower]) + self.rel_margin) * present[lower]
                w[:, c] += self.alpha * v
                w[:, lower] += self.alpha * v * 0.5
            for upper in rel.get("below", []):
                v = F.relu(-(height[upper] - height[c]) + self.rel_margin) * present[upper]
                w[:, c] += self.alpha * v
                w[:, upper] += self.alpha * v * 0.3
        w[:, 0] += self.alpha * (1 - (preds == 0).float().mean(dim=1))
        freq = (labels.view(-1, 1) == torch.arange(5, device=labels.device)).sum(dim=0).float().clamp(min=1)   # bincount
        cw = (1 / freq.sqrt()).to(logits.device)
        cw[1] *= 2.0
        cw[4] *= 2.0
        return _weighted_smoothed_ce(logits.reshape(-1, 5), labels.reshape(-1), w.mean(dim=0) * cw, 0.2)
