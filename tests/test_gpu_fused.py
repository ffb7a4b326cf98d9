"""Fused tcgen05 set-abstraction / EdgeConv block (csrc/sa_fused.cu) against (a) a PyTorch emulation
with the same bf16 operand rounding (tolerance: 2 bf16 ulps of the tile maximum -- accumulation
is fp32 on both sides) and (b) the unfused bf16-autocast modules (north_star bar for bf16: 1e-2
relative)."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import parity
from pointcloud_bridge_b200 import ops, synthetic
from pointcloud_bridge_b200.highway import DGCNN as dgcnn_mod
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg as ssg
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_stack(widths, seed):
    torch.manual_seed(seed)
    convs, bns = nn.ModuleList(), nn.ModuleList()
    for cin, cout in zip(widths[:-1], widths[1:]):
        convs.append(nn.Conv2d(cin, cout, 1))
        bn = nn.BatchNorm2d(cout)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.2)
            bn.running_mean.normal_(0, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
        bns.append(bn)
    return convs.to(DEV).eval(), bns.to(DEV).eval()


def emulate(rows, convs, bns, K, slope):
    """fp32 math on bf16-rounded operands, activations re-rounded to bf16 between layers."""
    x = rows.to(torch.bfloat16).float()
    for conv, bn in zip(convs, bns):
        w = conv.weight.flatten(1).float()
        b = conv.bias.float() if conv.bias is not None else torch.zeros(w.shape[0], device=DEV)
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        w = (w * scale[:, None]).to(torch.bfloat16).float()
        b = (b - bn.running_mean) * scale + bn.bias
        y = x @ w.t() + b
        x = torch.maximum(y, slope * y).to(torch.bfloat16).float()
    return x.view(-1, K, x.shape[-1]).max(dim=1)[0]


@pytest.mark.parametrize("D,K,S,widths,xyz_first", [
    (9, 32, 256, [12, 32, 32, 64], True),          # SSG sa1
    (9, 16, 200, [12, 16, 16, 32], False),         # MSG sa1, scale 0 (feat | dxyz), ragged tile count
    (96, 32, 64, [99, 64, 96, 128], False),        # MSG sa2, scale 1
    (0, 20, 128, [3, 16], True),                   # coordinates only, K does not divide 128
    (64, 16, 96, [67, 64, 64, 128], True),         # SSG sa2
])
def test_sa_fused_matches_bf16_emulation(D, K, S, widths, xyz_first):
    torch.manual_seed(D + K)
    B, N = 3, 512
    xyz = torch.rand(B, N, 3, device=DEV)
    pts = torch.randn(B, N, D, device=DEV) if D else None
    new_xyz = xyz[:, :S].contiguous()
    idx = torch.randint(0, N, (B, S, K), device=DEV)
    convs, bns = make_stack(widths, 1)
    pk = ops.PackedMLP(convs, bns, widths[0])
    assert pk.ok
    out = ops.sa_fused(xyz, pts, new_xyz, idx, pk, xyz_first=xyz_first, out_bf16=False)
    rows = ops.group_points(xyz, pts, new_xyz, idx, xyz_first=xyz_first).view(B * S * K, -1)
    ref = emulate(rows, convs, bns, K, 0.0)
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    tol = 2 * 2 ** -8 * ref.abs().max().item() + 1e-6
    print("sa_fused max err", err, "tol", tol)
    assert err <= tol
    assert (out - ref).abs().mean().item() <= 1e-3 * ref.abs().max().item()


def test_edgeconv_fused_matches_emulation():
    torch.manual_seed(3)
    B, N, D, K = 2, 1024, 64, 20
    x = torch.randn(B, D, N, device=DEV)
    idx = ops.knn(x, K)
    conv = nn.Conv2d(2 * D, 64, 1, bias=False)
    convs, bns = make_stack([2 * D, 64], 2)
    convs[0] = conv.to(DEV)
    pk = ops.PackedMLP(convs, bns, 2 * D)
    out = ops.sa_fused(None, x.transpose(1, 2).contiguous(), None, idx, pk, mode=1, slope=0.2, out_bf16=False)
    gf = ops.graph_feature(x, idx)                                # [B,2D,N,K]
    rows = gf.permute(0, 2, 3, 1).reshape(B * N * K, 2 * D)
    ref = emulate(rows, convs, bns, K, 0.2)
    err = (out - ref).abs().max().item()
    assert err <= 2 * 2 ** -8 * ref.abs().max().item() + 1e-6


def test_networks_fused_vs_unfused_bf16(monkeypatch):
    """Whole eval forwards under bf16 autocast: fused blocks on vs off (PCB_NO_FUSED=1)."""
    g = parity.load("models.npz")
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"])).to(DEV)
    xyz = torch.from_numpy(g["xyz"]).to(DEV)
    rgb = torch.from_numpy(g["rgb"]).to(DEV)
    cases = [(parity.seeded_fill_(ssg.get_model(13), 1), (x9[:1],), g["ssg_logp"]),
             (parity.seeded_fill_(msg.get_model(5), 2), (x9,), g["msg_logp"]),
             (parity.seeded_fill_(dgcnn_mod.DGCNN(5, 20), 3), (xyz[:1], rgb[:1]), g["dgcnn_logits"])]
    for net, args, ref in cases:
        net = net.to(DEV).eval()
        outs = []
        for fused in ("0", "1"):
            monkeypatch.setenv("PCB_NO_FUSED", fused)
            torch.manual_seed(4242)
            n0 = ops._lib.launches()
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(*args)
            y = (y[0] if isinstance(y, tuple) else y).float().cpu().numpy()
            outs.append(y)
        scale = np.abs(ref).max()
        e_fused = np.abs(outs[0] - ref).mean() / scale
        e_plain = np.abs(outs[1] - ref).mean() / scale
        print(type(net).__name__, "mean rel err vs fp32 reference: fused %.2e unfused %.2e" % (e_fused, e_plain))
        assert e_fused < 1e-2
        assert (outs[0].argmax(-1) == ref.argmax(-1)).mean() > 0.97


@pytest.mark.parametrize("widths,M,K", [
    ([131, 128, 128, 256], 64 * 32 * 2, 32),       # SSG sa3: wider than the one-kernel block takes
    ([259, 256, 256, 512], 16 * 32 * 3, 32),       # SSG sa4 / MSG sa4 scale 0
    ([515, 256, 384, 512], 16 * 16 * 2, 16),       # MSG sa4 scale 1
    ([259, 128, 196, 256], 64 * 16 * 2, 16),       # MSG sa3: 196 channels carried as 200
    ([1536, 256, 256], 300, 1),                    # feature propagation, ragged row count
    ([12, 32, 32, 64], 20 * 50, 20),               # pool_k does not divide 128: max taken outside the epilogue
])
def test_mlp_rows_infer_every_layer_width(widths, M, K):
    """BatchNorm-folded inference MLP on the tcgen05 GEMM with bias + ReLU (+ max over neighbours) epilogues
    (ops.mlp_rows_infer) against the bf16-rounding emulation: the layers that the one-kernel block rejects."""
    torch.manual_seed(M + K)
    convs, bns = make_stack(widths, 7)
    rows = torch.randn(M, widths[0], device=DEV)
    out = ops.mlp_rows_infer(rows, ops.FoldedMLP(convs, bns), pool_k=K)[:, :widths[-1]].float()
    ref = emulate(rows, convs, bns, K, 0.0)
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    tol = 3 * 2 ** -8 * ref.abs().max().item() + 1e-6
    print("mlp_rows_infer max err", err, "tol", tol)
    assert err <= tol
    assert (out - ref).abs().mean().item() <= 1e-3 * ref.abs().max().item()


def test_eval_bf16_networks_launch_no_library_gemm_for_the_sa_layers(monkeypatch):
    """SSG / MSG evaluation under bf16 autocast: every set-abstraction and feature-propagation MLP runs on kernels of
    libpcbridge (one-kernel block or BatchNorm-folded tcgen05 GEMMs) -- checked through the C-ABI entry counter."""
    g = parity.load("models.npz")
    x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"])).to(DEV)
    for net in (parity.seeded_fill_(ssg.get_model(13), 1), parity.seeded_fill_(msg.get_model(5), 2)):
        net = net.to(DEV).eval()
        seen = []
        orig = ops._call

        def spy(name, *a, **kw):
            seen.append(name)
            return orig(name, *a, **kw)
        monkeypatch.setattr(ops, "_call", spy)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            net(x9[:1])
        monkeypatch.setattr(ops, "_call", orig)
        n_fused = seen.count("pcb_sa_fused_bf16")
        n_gemm = seen.count("pcb_linear_bias_act_rows_bf16")
        print(type(net).__module__, "one-kernel blocks:", n_fused, "folded GEMM layers:", n_gemm)
        assert n_fused >= 2 and n_gemm >= 6 + 9            # wide SA layers + the feature-propagation stacks
