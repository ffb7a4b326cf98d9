"""DGCNN semantic segmentation network, drop-in for Highway_bridge/models/DGCNN.py
(`DGCNN(num_classes=5, k=20)`; methods `knn`, `get_graph_feature`; forward(xyz [B,N,3],
features [B,N,C] | None) -> logits [B,N,num_classes]); same parameter names
(`conv1.0.weight`, `bn1.*`, ..., `local_bn`, `point_conv.*`).

`knn` never builds the [B,N,N] distance matrix (1 GiB at B=16, N=4096) and `get_graph_feature`
writes the [B,2D,N,k] edge tensor in one pass; both are libpcbridge kernels.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class DGCNN(nn.Module):
    def __init__(self, num_classes=5, k=20):
        super().__init__()
        self.k = k
        self.bn1 = nn.BatchNorm2d(64)
        self.bn2 = nn.BatchNorm2d(64)
        self.bn3 = nn.BatchNorm2d(64)
        self.bn4 = nn.BatchNorm2d(128)
        self.bn5 = nn.BatchNorm1d(1024)

        def edge_block(cin, cout, bn):
            return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), bn, nn.LeakyReLU(negative_slope=0.2))

        self.conv1 = edge_block(6, 64, self.bn1)
        self.conv2 = edge_block(64 * 2, 64, self.bn2)
        self.conv3 = edge_block(64 * 2, 64, self.bn3)
        self.conv4 = edge_block(64 * 2, 128, self.bn4)
        self.conv5 = nn.Sequential(nn.Conv1d(320, 1024, kernel_size=1, bias=False), self.bn5,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.local_bn = nn.BatchNorm1d(320)
        self.point_conv = nn.Sequential(
            nn.Conv1d(1344, 512, 1), nn.BatchNorm1d(512), nn.LeakyReLU(negative_slope=0.2),
            nn.Conv1d(512, 256, 1), nn.BatchNorm1d(256), nn.LeakyReLU(negative_slope=0.2),
            nn.Conv1d(256, num_classes, 1))

    def knn(self, x, k):
        """x [B,D,N] -> LongTensor [B,N,k]: k nearest (self included) by the pairwise distance of
        DGCNN.py:63-65, ordered by (distance, index)."""
        return ops.knn(x, k, channels_first=True)

    def get_graph_feature(self, x, k=20, idx=None):
        """x [B,D,N] -> [B,2D,N,k] = cat(x[idx] - x, x) on channels (DGCNN.py:72-109)."""
        if idx is None:
            idx = self.knn(x, k)
        return ops.graph_feature(x, idx)

    def forward(self, xyz, features=None):
        B, N, _ = xyz.shape
        # DGCNN.py:121-128: features are concatenated and immediately sliced away again
        x = xyz.transpose(2, 1)[:, :3, :].contiguous()
        k = min(self.k, N - 1)
        feats = []
        for li, conv in enumerate((self.conv1, self.conv2, self.conv3, self.conv4)):
            x = x.float()
            if ops.fused_inference_enabled() and not conv[1].training and k <= 128:
                # EdgeConv fused on the tensor cores: the [B,2D,N,k] edge tensor never exists
                pk = ops.packed_mlp(self, li, [conv[0]], [conv[1]], 2 * x.shape[1])
                if pk.ok:
                    idx = self.knn(x, k)
                    y = ops.sa_fused(None, x.transpose(1, 2).contiguous(), None, idx, pk, mode=1, slope=0.2)
                    x = y.view(B, N, -1).permute(0, 2, 1)
                    feats.append(x)
                    continue
            x = conv(self.get_graph_feature(x, k=k)).max(dim=-1)[0]
            feats.append(x)
        local = torch.cat(feats, dim=1)                                  # [B,320,N]
        local_n = F.leaky_relu(self.local_bn(local), negative_slope=0.2)
        glob = F.adaptive_max_pool1d(self.conv5(local), 1).expand(-1, -1, N)
        logits = self.point_conv(torch.cat([local_n, glob], dim=1))
        return logits.transpose(1, 2)
