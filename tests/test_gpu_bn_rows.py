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


@pytest.mark.parametrize("M,Cv,pool_k", [(1024 * 16, 196, 1), (512 * 16, 196, 16), (300, 12, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_relu_rows_with_zero_pad_columns(M, Cv, pool_k, dtype):
    """Rows padded with zero columns up to a multiple of 8 (196 -> 200 channels): the real channels match the
    unpadded computation, the pad columns of the output and of the input gradient are exactly zero."""
    torch.manual_seed(M + Cv)
    C = -(-Cv // 8) * 8
    y0 = (torch.randn(M, Cv, device=DEV) * 2 + 3 * torch.randn(Cv, device=DEV)).to(dtype)
    yp0 = torch.zeros(M, C, device=DEV, dtype=dtype)
    yp0[:, :Cv] = y0
    bias0 = torch.randn(Cv, device=DEV)
    bn_a, bn_b = torch.nn.BatchNorm1d(Cv).to(DEV).train(), torch.nn.BatchNorm1d(Cv).to(DEV).train()
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5)
        bn_a.bias.normal_()
        bn_b.load_state_dict(bn_a.state_dict())
    ya, ba = yp0.clone().requires_grad_(True), bias0.clone().requires_grad_(True)
    yb, bb = y0.clone().requires_grad_(True), bias0.clone().requires_grad_(True)
    za = ops.bn_relu_rows(ya, ba, bn_a, relu=True, pool_k=pool_k)
    zb = ops.bn_relu_rows(yb, bb, bn_b, relu=True, pool_k=pool_k)
    assert za.shape == (M // pool_k, C) and not za[:, Cv:].any()
    ftol = 1e-5 if dtype == torch.float32 else 1e-2        # the partial-sum grouping differs with the row pitch
    assert (za[:, :Cv].float() - zb.float()).abs().max().item() <= ftol * (zb.float().abs().max().item() + 1e-6)
    assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-5, atol=1e-6)
    g = torch.randn_like(zb)
    gp = torch.zeros_like(za)
    gp[:, :Cv] = g
    za.backward(gp)
    zb.backward(g)
    assert not ya.grad[:, Cv:].any()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (ya.grad[:, :Cv].float() - yb.grad.float()).abs().max().item() <= tol * (yb.grad.abs().max().item() + 1e-9)
    for pa, pb in ((bn_a.weight, bn_b.weight), (bn_a.bias, bn_b.bias)):
        assert pa.grad.shape == pb.grad.shape
        assert (pa.grad - pb.grad).abs().max().item() <= 1e-4 * (pb.grad.abs().max().item() + 1e-6)


def test_linear_rows_padded_output_channels():
    """linear_rows(pad_n=True) with a StepContext shadow: 196 output channels come out as 200 columns, the
    extra ones zero; values, input gradient and weight gradient equal the unpadded layer's."""
    torch.manual_seed(0)
    M, K, N = 4096, 24, 196
    conv = torch.nn.Conv2d(K, N, 1).to(DEV)
    nxt = torch.nn.Conv2d(N, 64, 1).to(DEV)
    mod = torch.nn.ModuleList([conv, nxt])
    x = torch.randn(M, K, device=DEV).to(torch.bfloat16).requires_grad_(True)
    ctx = ops.StepContext(mod, bf16=True)
    with ctx, torch.autocast("cuda", dtype=torch.bfloat16):
        yp = ops.linear_rows(x, conv.weight.flatten(1), pad_n=True)
        yr = ops.linear_rows(x, conv.weight.flatten(1), pad_n=False)
        assert yp.shape == (M, 200) and yr.shape == (M, N) and not yp[:, N:].any()
        assert (yp[:, :N].float() - yr.float()).abs().max().item() <= 1e-2 * yr.float().abs().max().item()
        # the next layer consumes the padded rows as they are
        zp = ops.linear_rows(yp, nxt.weight.flatten(1))
        zr = ops.linear_rows(yr, nxt.weight.flatten(1))
        assert (zp.float() - zr.float()).abs().max().item() <= 2e-2 * zr.float().abs().max().item()
        g = torch.randn_like(zr)
        gp = torch.autograd.grad(zp, [x, conv.weight, nxt.weight], g)
        gr = torch.autograd.grad(zr, [x, conv.weight, nxt.weight], g)
    for a, b in zip(gp, gr):
        assert a.shape == b.shape
        assert ((a.float() - b.float()).norm() / b.float().norm()).item() <= 2e-2


def test_flat_adam_matches_torch_adam_and_refreshes_shadows():
    """engine.FlatAdam (csrc/adam.cu) against torch.optim.Adam over 5 steps with weight decay, and the bf16
    shadow of a [12, 10] weight inside the flat buffer (zero-padded to [16, 16]) after every step."""
    from pointcloud_bridge_b200.engine import FlatAdam
    torch.manual_seed(0)
    n = 100_003
    p0 = torch.randn(n, device=DEV)
    ref = p0.clone().requires_grad_(True)
    opt_ref = torch.optim.Adam([ref], lr=1e-2, weight_decay=1e-2)
    p, g = p0.clone(), torch.zeros(n, device=DEV)
    off, rows, k, k8 = 777, 12, 10, 16
    index = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    j = torch.arange(rows * k, device=DEV)
    index[off:off + rows * k] = ((j // k) * k8 + (j % k)).int()
    shadow = torch.zeros(16 * k8, dtype=torch.bfloat16, device=DEV)
    opt = FlatAdam(p, g, lr=1e-2, weight_decay=1e-2, shadow_index=index, shadow_flat=shadow)
    for it in range(5):
        grad = torch.randn(n, device=DEV) * (1 + it)
        g.copy_(grad)
        ref.grad = grad.clone()
        if it == 3:
            opt.set_lr(3e-3)
            opt_ref.param_groups[0]["lr"] = 3e-3
        opt.step()
        opt_ref.step()
        torch.testing.assert_close(p, ref.detach(), rtol=2e-5, atol=2e-6)
        sh = shadow.view(16, k8)
        assert torch.equal(sh[:rows, :k], p[off:off + rows * k].view(rows, k).to(torch.bfloat16))
        assert not sh[rows:].any() and not sh[:, k:].any()
    assert int(opt.step_t) == 5


@pytest.mark.parametrize("dtype,pitch", [(torch.bfloat16, 8), (torch.float32, 5), (torch.float32, 16)])
def test_nll_logit_rows_matches_log_softmax_nll(dtype, pitch):
    """Loss kernel of the segmentation head (csrc/loss.cu) against F.nll_loss(F.log_softmax(logits + bias))."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    M, nc = 70_001, 5
    rows = torch.zeros(M, pitch, device=DEV, dtype=dtype)
    rows[:, :nc] = (torch.randn(M, nc, device=DEV) * 3).to(dtype)
    bias = torch.randn(nc, device=DEV)
    labels = torch.randint(0, nc, (M,), device=DEV)
    ra, ba = rows.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    rb, bb = rows.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    la = ops.nll_logit_rows(ops.LogitRows(ra, ba, nc, 1, M), labels)
    lb = F.nll_loss(F.log_softmax(rb[:, :nc].float() + bb, dim=-1), labels)
    assert abs(la.item() - lb.item()) <= 2e-6 * abs(lb.item()) + 1e-6
    (la * 2.5).backward()
    (lb * 2.5).backward()
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert (ra.grad[:, :nc].float() - rb.grad[:, :nc].float()).abs().max().item() <= tol * rb.grad.abs().max().item() + 1e-12
    assert not ra.grad[:, nc:].any()
    torch.testing.assert_close(ba.grad, bb.grad, rtol=1e-4, atol=1e-7)
    assert torch.equal(ops.LogitRows(rows, bias, nc, 1, M).log_probs().view(M, nc),
                       torch.log_softmax(rows[:, :nc].float() + bias, -1))


def test_eigvalsh3_matches_torch():
    """Closed-form float64 eigenvalues of symmetric 3x3 matrices against torch.linalg.eigvalsh (float64 reference):
    covariance-like, near-planar (one tiny eigenvalue), diagonal and repeated-eigenvalue cases."""
    torch.manual_seed(0)
    M = 20_000
    x = torch.randn(M, 32, 3, device=DEV) * torch.tensor([1.0, 0.3, 0.002], device=DEV)     # near-planar patches
    cov = torch.bmm(x.transpose(1, 2), x) / 31
    cov[:100] = torch.diag_embed(torch.rand(100, 3, device=DEV))                             # diagonal
    cov[100:200] = torch.eye(3, device=DEV) * torch.rand(100, 1, 1, device=DEV)              # triple eigenvalue
    ev = ops.eigvalsh3(cov)
    ref = torch.linalg.eigvalsh(cov.double())
    assert ev.shape == (M, 3) and (ev[:, 1:] >= ev[:, :-1]).all()
    scale = ref.abs().amax(dim=1, keepdim=True)
    assert ((ev.double() - ref).abs() / scale).max().item() <= 2e-7        # fp32 rounding of the result only
    assert ops.eigvalsh3(cov.view(100, 200, 3, 3)).shape == (100, 200, 3)
