"""BridgeStructureEncoding input rows in one kernel (csrc/structure.cu) against the reference-shaped PyTorch
composition (attention_modules.py:552-613, 622-687 of the reference; `get_structure_features` of the drop-in module)."""
import numpy as np
import pytest
import torch

from pointcloud_bridge_b200 import ops, synthetic
from pointcloud_bridge_b200.highway import attention_modules as am

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cloud(seed, B, N):
    xyz, _, _ = synthetic.bridge_batch(seed, B, N)
    return torch.from_numpy(np.ascontiguousarray(xyz)).to(DEV)


@pytest.mark.parametrize("B,N,k,F", [(2, 1024, 16, 4), (1, 300, 16, 4), (2, 257, 7, 2), (1, 64, 32, 8)])
def test_structure_rows_match_the_pytorch_composition(B, N, k, F):
    xyz = _cloud(11 + N, B, N) * 3.0 + 0.37                                  # several grid cells
    enc = am.BridgeStructureEncoding(channels=16, k_neighbors=k, freq_bands=F).to(DEV).eval()
    idx = ops.knn_cdist(xyz, k)
    rows, feat = ops.structure_rows(xyz, idx, enc.freqs, enc.grid_size, bf16=False, feat=True)
    rel = ops.group_points(xyz, None, xyz, idx, xyz_first=True)              # [B,N,k,3], exact subtraction
    ref_feat = enc.get_structure_features(rel)                               # PyTorch ops (cuSOLVER eigvalsh in eval)
    ref_abs = enc.compute_absolute_position_encoding(xyz)
    a = 6 * F
    r = rows.view(B, N, k, -1)
    assert rows.shape[1] == (a + 16 + 7) // 8 * 8
    assert torch.equal(r[..., a:a + 3], rel)                                 # neighbour - centre: bit for bit
    assert torch.allclose(r[..., :a], ref_abs[:, :, None, :].expand(-1, -1, k, -1), rtol=0, atol=2e-6)
    assert torch.equal(r[..., a + 3:a + 16], feat[:, :, None, :].expand(-1, -1, k, -1))
    if rows.shape[1] > a + 16:
        assert float(r[..., a + 16:].abs().max()) == 0.0
    # statistics 3..12: plain fp32 reductions, summation order differs
    scale = float(rel.abs().max())
    assert torch.allclose(feat[..., 3:], ref_feat[..., 3:], rtol=2e-4, atol=2e-5 * scale)
    # shape features divide by the smallest eigenvalue: compare where the neighbourhood is not degenerate (the fp32
    # solver of the reference leaves that eigenvalue with an absolute error of ~1e-7 * the largest one)
    cov = torch.einsum("bnki,bnkj->bnij", rel.double(), rel.double()) / (k - 1)
    ev = torch.linalg.eigvalsh(cov)
    ok = ev[..., 0] > 1e-3 * ev[..., 2]
    assert int(ok.sum()) > 0.3 * ok.numel()
    den = ev[..., 0] + 1e-8
    exact = torch.stack([(ev[..., 0] - ev[..., 1]) / den, (ev[..., 1] - ev[..., 2]) / den, ev[..., 2] / den], -1).float()
    err_ours = ((feat[..., :3] - exact).abs() / (exact.abs() + 1))[ok]
    err_ref = ((ref_feat[..., :3] - exact).abs() / (exact.abs() + 1))[ok]
    print(f"shape features vs float64: ours max {float(err_ours.max()):.2e}, PyTorch fp32 max {float(err_ref.max()):.2e}")
    assert float(err_ours.max()) < 2e-3 and float(err_ours.max()) <= max(2.0 * float(err_ref.max()), 1e-4)
    # bf16 rows: the fp32 rows rounded once
    rows16, _ = ops.structure_rows(xyz, idx, enc.freqs, enc.grid_size, bf16=True)
    assert torch.equal(rows16, rows.to(torch.bfloat16))


@pytest.mark.parametrize("train", [False, True])
def test_encoder_fused_rows_equal_composed_rows(train, monkeypatch):
    """The whole encoder through the fused rows against the composed rows: same network output up to the
    conditioning of the eigenvalue features."""
    torch.manual_seed(3)
    xyz = _cloud(5, 2, 2048) * 2.0
    enc = am.BridgeStructureEncoding(channels=32, k_neighbors=16).to(DEV)
    enc.train(train)
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(am, "_FUSED_STRUCTURE", fused)
        torch.manual_seed(0)
        with torch.no_grad():
            outs.append(enc(xyz).float())
    rel = float((outs[0] - outs[1]).norm() / outs[1].norm())
    print("encoder output, fused vs composed rows: rel L2 %.2e" % rel)
    assert rel < 2e-3
