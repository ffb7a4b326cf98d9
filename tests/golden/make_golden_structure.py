"""Generates tests/golden/structure.npz: outputs of the UNMODIFIED reference `BridgeStructureEncoding`
(Highway_bridge/models/attention_modules.py:523-687, imported from /root/reference) on a seeded synthetic cloud, CPU fp32:
its kNN indices, relative positions, `get_structure_features`, `compute_absolute_position_encoding` and the encoder output
with name-keyed seeded weights (tests/parity.py:seeded_fill_).  Pins csrc/structure.cu to the reference itself.

    python tests/golden/make_golden_structure.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import _reference  # noqa: E402
import parity  # noqa: E402
from pointcloud_bridge_b200 import synthetic  # noqa: E402


def main():
    assert _reference.available()
    torch.set_num_threads(8)
    _, _, am, _ = _reference.highway()
    out = {}
    for tag, (seed, B, N, k, scale) in {"a": (21, 2, 1024, 16, 3.0), "b": (22, 1, 512, 32, 1.0)}.items():
        xyz = synthetic.bridge_batch(seed, B, N)[0] * scale + 0.37
        t = torch.from_numpy(np.ascontiguousarray(xyz))
        enc = parity.seeded_fill_(am.BridgeStructureEncoding(channels=32, k_neighbors=k), 3).eval()
        with torch.no_grad():
            dist = torch.cdist(t, t)
            _, idx = dist.topk(k, dim=-1, largest=False)                       # attention_modules.py:584-586
            nbr = t.view(B * N, -1)[(idx + torch.arange(B).view(-1, 1, 1) * N).view(-1)].view(B, N, k, -1)
            rel = nbr - t.unsqueeze(2)
            feat = enc.get_structure_features(rel)
            absenc = enc.compute_absolute_position_encoding(t)
            y = enc(t)
        out[f"{tag}_xyz"], out[f"{tag}_idx"], out[f"{tag}_rel"] = xyz.astype(np.float32), idx.numpy(), rel.numpy()
        out[f"{tag}_feat"], out[f"{tag}_abs"], out[f"{tag}_out"] = feat.numpy(), absenc.numpy(), y.numpy()
        out[f"{tag}_k"] = np.int64(k)
    np.savez_compressed(os.path.join(HERE, "structure.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
