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
    ref_feat = enc.get_structure_features(rel)                               # PyTorch ops on an fp32 bmm covariance
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


@pytest.mark.parametrize("tag", ["a", "b"])
def test_structure_kernel_against_the_reference_golden(tag):
    """csrc/structure.cu against outputs of the UNMODIFIED reference class on CPU (tests/golden/structure.npz) and the
    numpy oracle: neighbour sets, offsets, the 13 statistics, the position encoding, and the encoder's output with the
    reference's seeded weights."""
    import parity
    from oracle import structure_oracle as so
    g = parity.load("structure.npz")
    xyz_np, idx_np, rel_np, feat_np, abs_np, out_np = (g[f"{tag}_{n}"] for n in ("xyz", "idx", "rel", "feat", "abs", "out"))
    k = int(g[f"{tag}_k"])
    xyz = torch.from_numpy(xyz_np).to(DEV)
    B, N, _ = xyz.shape
    # the kNN kernel finds the reference's neighbour SETS (order inside exact distance ties is unspecified in topk)
    idx = ops.knn_cdist(xyz, k)
    same = (np.sort(idx.cpu().numpy(), -1) == np.sort(idx_np, -1)).all(-1).mean()
    assert same > 0.999, same
    # on the reference's own indices: rows and statistics
    ridx = torch.from_numpy(idx_np).to(DEV)
    freqs = [1.0, 2.0, 4.0, 8.0]
    rows, feat = ops.structure_rows(xyz, ridx, freqs, 1.0, bf16=False, feat=True)
    r = rows.view(B, N, k, -1).cpu().numpy()
    f = feat.cpu().numpy()
    assert np.array_equal(r[..., 24:27], rel_np)
    assert np.allclose(r[..., :24], np.broadcast_to(abs_np[:, :, None, :], r[..., :24].shape), rtol=0, atol=2e-6)
    scale = float(np.abs(rel_np).max())
    assert np.allclose(f[..., 3:], feat_np[..., 3:], rtol=2e-4, atol=2e-5 * scale)
    ok = so.well_conditioned(rel_np)
    err = np.abs(f[..., :3] - feat_np[..., :3]) / (np.abs(feat_np[..., :3]) + 1)
    print("shape features vs the reference (well-conditioned neighbourhoods): max rel %.2e" % err[ok].max())
    assert err[ok].max() < 2e-3
    orows = so.rows(xyz_np, idx_np, np.asarray(freqs)).reshape(B, N, k, 40)              # the numpy oracle's rows
    assert np.allclose(np.delete(r, [27, 28, 29], -1), np.delete(orows, [27, 28, 29], -1), rtol=2e-4, atol=2e-5 * scale)
    # the whole encoder with the reference's weights, evaluation mode, fp32
    enc = parity.seeded_fill_(am.BridgeStructureEncoding(channels=32, k_neighbors=k), 3).to(DEV).eval()
    with torch.no_grad():
        y = enc(xyz).float().cpu().numpy()
    e = np.abs(y - out_np) / (np.abs(out_np).max() + 1e-12)
    print("encoder output vs the reference: p50 %.2e p99 %.2e max %.2e" % (np.median(e), np.quantile(e, 0.99), e.max()))
    assert np.median(e) < 1e-5 and np.quantile(e, 0.99) < 1e-3
