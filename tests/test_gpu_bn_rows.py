"""Fused BatchNorm(train)+ReLU(+max over neighbours) row kernels (csrc/bn_rows.cu) against the
plain PyTorch composition the reference uses (Conv bias -> BatchNorm -> ReLU -> torch.max,
pointnet_util.py:213-217), forward, backward and running statistics.  fp32: 2e-5 relative;
bf16 activations: 2e-2 (one bf16 rounding of the output)."""
import pytest
import torch
import torch.nn.functional as F

from pointcloud_bridge_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def reference(y, bias, bn, pool_k):
    z = F.relu(F.batch_norm(y.float() + bias, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps))
    if pool_k > 1:
        z = z.view(-1, pool_k, z.shape[-1]).max(dim=1)[0]
    return z


@pytest.mark.parametrize("M,C,pool_k", [(4096, 16, 1), (2048 * 32, 64, 32), (1024 * 16, 196, 16), (999 * 3, 512, 3), (64, 1024, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_relu_rows_matches_torch(M, C, pool_k, dtype):
    torch.manual_seed(M + C)
    y0 = (torch.randn(M, C, device=DEV) * 2 + 3 * torch.randn(C, device=DEV)).to(dtype)
    bias0 = torch.randn(C, device=DEV)
    bn_a, bn_b = torch.nn.BatchNorm1d(C).to(DEV).train(), torch.nn.BatchNorm1d(C).to(DEV).train()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.normal_()
        bn_b.load_state_dict(bn_a.state_dict())
    ya, ba = y0.clone().requires_grad_(True), bias0.clone().requires_grad_(True)
    yb, bb = y0.clone().requires_grad_(True), bias0.clone().requires_grad_(True)
    assert ops.bn_rows_supported(ya, bn_a, pool_k)
    za = ops.bn_relu_rows(ya, ba, bn_a, relu=True, pool_k=pool_k)
    zb = reference(yb, bb, bn_b, pool_k)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    scale = zb.abs().max().item() + 1e-6
    assert (za.float() - zb).abs().max().item() <= tol * scale
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-4, atol=1e-5)
    assert int(bn_a.num_batches_tracked) == 1
    g = torch.randn_like(zb)
    za.backward(g.to(za.dtype))
    zb.backward(g)
    gscale = yb.grad.abs().max().item() + 1e-9
    gtol = 5e-4 if dtype == torch.float32 else 3e-2
    assert (ya.grad.float() - yb.grad.float()).abs().max().item() <= gtol * gscale
    for pa, pb in ((bn_a.weight, bn_b.weight), (bn_a.bias, bn_b.bias)):
        assert (pa.grad - pb.grad).abs().max().item() <= gtol * (pb.grad.abs().max().item() + 1e-6)
    # the conv bias feeds a training-mode BN: its gradient is zero up to rounding on both paths
    assert ba.grad.abs().max().item() <= 1e-2 * (bn_b.weight.grad.abs().max().item() + 1.0)


@pytest.mark.parametrize("M,N,K,Kp", [(4096, 16, 12, 16), (100000, 32, 32, 32), (5000, 96, 99, 104), (777, 256, 520, 520),
                                      (65, 8, 8, 8), (20000, 128, 67, 72)])
def test_wgrad_rows_kernel(M, N, K, Kp):
    """gw = gy^T x over the long row dimension (bf16 operands, fp32 accumulation) against an fp64 matmul."""
    from pointcloud_bridge_b200 import ops
    torch.manual_seed(M + N)
    gy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    x = torch.zeros(M, Kp, device="cuda", dtype=torch.bfloat16)
    x[:, :K] = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    assert ops.wgrad_rows_supported(gy, x)
    gw = ops.wgrad_rows(gy, x, K)
    ref = (gy.double().t() @ x.double())[:, :K]
    assert gw.shape == (N, K) and gw.dtype == torch.float32
    scale = ref.abs().max().item()
    assert (gw.double() - ref).abs().max().item() <= 2e-5 * scale + 1e-4 * (M ** 0.5) * 1e-2
    # accumulation into an existing buffer
    out = torch.ones(N, K, device="cuda")
    ops.wgrad_rows(gy, x, K, out=out)
    assert (out.double() - 1.0 - ref).abs().max().item() <= 2e-5 * scale + 1e-4 * (M ** 0.5) * 1e-2
