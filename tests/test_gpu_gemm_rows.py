"""tcgen05 training GEMMs (csrc/gemm_rows.cu) and the BatchNorm halves around them against their PyTorch
composition (fp32 reference of the same op on the same bf16 operands).

Tolerances: the kernels accumulate in fp32 and round once to bf16, so a product must be within one bf16 ulp
(2^-8 relative) + fp32 summation-order noise of the fp32 reference; statistics are compared with the statistics of
the kernel's own bf16 output (they are defined on the stored values)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from pointcloud_bridge_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

SHAPES = [  # (M, K, N): K, N multiples of 8
    (129, 8, 8), (1000, 16, 16), (65536, 32, 64), (4173, 104, 200), (1024, 1536, 256), (300, 520, 512),
    (8192, 264, 128), (128, 64, 1024), (40000, 200, 256), (77, 72, 24),
]


def bf16_close(a, ref, what=""):
    a, ref = a.float(), ref.float()
    tol = ref.abs() * 2.0 ** -7 + 2e-3 * ref.abs().max().clamp_min(1e-6) * 2.0 ** -7 + 1e-6
    bad = (a - ref).abs() > tol
    assert not bad.any(), (what, int(bad.sum()), float((a - ref).abs().max()), float(ref.abs().max()))


@pytest.mark.parametrize("M,K,N", SHAPES)
def test_gemm_rows_matches_fp32_matmul(M, K, N):
    g = torch.Generator(device=DEV).manual_seed(M + K + N)
    x = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=DEV, generator=g) / K ** 0.5).to(torch.bfloat16)
    y = ops.gemm_rows(x, w, N)
    ref = x.float() @ w.float().t()
    bf16_close(y, ref, "plain")
    # fewer weight rows than output columns: the missing rows count as zero; pitched operands
    xw = torch.zeros(M, K + 16, device=DEV, dtype=torch.bfloat16)
    xw[:, :K] = x
    y2 = ops.gemm_rows(xw[:, :K], w[:max(N - 8, 8)], N)
    ref2 = ref.clone()
    ref2[:, max(N - 8, 8):] = 0
    bf16_close(y2, ref2, "pitched / short weight")


def _bn(n, seed):
    bn = nn.BatchNorm1d(n).to(DEV).train()
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(n, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(n, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(n, generator=g) + 0.5)
    return bn


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("M,K,N", [(1000, 16, 16), (65536, 32, 64), (4173, 104, 200), (1024, 1536, 256), (300, 520, 512),
                                   (524288, 16, 32)])
def test_linear_bn_stats_epilogue(M, K, N, deferred):
    """deferred: the GEMM leaves per-group (count, mean, M2) triples, the elementwise kernel merges them."""
    import ctypes
    lib = _lib.lib()
    g = torch.Generator(device=DEV).manual_seed(7 * M + K)
    Cv = N - 4 if N == 200 else N                        # 196 real channels carried as 200
    x = (torch.randn(M, K, device=DEV, generator=g) + 0.3).to(torch.bfloat16)
    w = torch.zeros(N, K, device=DEV)
    w[:Cv] = torch.randn(Cv, K, device=DEV, generator=g) / K ** 0.5
    w = w.to(torch.bfloat16)
    bias = torch.randn(Cv, device=DEV, generator=g)
    rm0 = torch.randn(Cv, device=DEV, generator=g) * 0.1
    rv0 = torch.rand(Cv, device=DEV, generator=g) + 0.5
    outs = []
    for rep in range(2):
        rm, rv = rm0.clone(), rv0.clone()
        y = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        stats = torch.full((3, N), float("nan"), device=DEV)
        work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=DEV)
        tick = ops._tickets(torch.device(DEV))
        gparts = torch.full((lib.pcb_gemm_max_groups(), 3, N), float("nan"), device=DEV) if deferred else None
        groups = ctypes.c_int(0)
        ops._call("pcb_linear_bn_stats_rows_bf16", x.device, x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0), M, N, N, K,
                  y.data_ptr(), y.stride(0), Cv, 1e-5,
                  None if deferred else stats[0].data_ptr(), None if deferred else stats[1].data_ptr(),
                  None if deferred else stats[2].data_ptr(), work.data_ptr(), tick.data_ptr(),
                  gparts.data_ptr() if deferred else None, ctypes.byref(groups))
        assert 1 <= groups.value <= lib.pcb_gemm_max_groups()
        # the elementwise kernel that follows owns the running statistics (and, deferred, the final merge)
        z = torch.empty_like(y)
        ones, zeros = torch.ones(Cv, device=DEV), torch.zeros(Cv, device=DEV)
        ops._call("pcb_bn_apply_rows", y.device, y.data_ptr(), 1, M, N, Cv, 1, stats[0].data_ptr(), stats[1].data_ptr(),
                  ones.data_ptr(), zeros.data_ptr(), 1, z.data_ptr(), z.stride(0), None, stats[2].data_ptr(),
                  bias.data_ptr(), 0.1, rm.data_ptr(), rv.data_ptr(), gparts.data_ptr() if deferred else None,
                  groups.value, 1e-5, None)
        torch.cuda.synchronize()
        zr = torch.relu((y.float()[:, :Cv] - stats[0, :Cv]) * stats[1, :Cv])
        assert torch.allclose(z.float()[:, :Cv], zr, rtol=2.0 ** -7, atol=1e-5)
        assert int(tick.abs().sum()) == 0                 # tickets are left zeroed
        outs.append((y, stats, rm, rv))
    y, stats, rm, rv = outs[0]
    bf16_close(y, x.float() @ w.float().t(), "y")
    yf = y.float()[:, :Cv]
    mean, var = yf.mean(0), yf.var(0, unbiased=False)
    assert torch.allclose(stats[0, :Cv], mean, rtol=1e-4, atol=1e-5 * float(yf.abs().max()))
    assert torch.allclose(stats[1, :Cv], torch.rsqrt(var + 1e-5), rtol=2e-4)
    assert torch.allclose(stats[2, :Cv], var, rtol=2e-4, atol=1e-7)
    assert torch.allclose(rm, 0.9 * rm0 + 0.1 * (mean + bias), rtol=1e-4, atol=1e-5)
    assert torch.allclose(rv, 0.9 * rv0 + 0.1 * yf.var(0, unbiased=True), rtol=2e-4, atol=1e-6)
    if Cv < N:
        assert float(stats[:, Cv:].abs().max()) == 0.0
    # deterministic: same bits on the second run
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("deferred", [False, True])
@pytest.mark.parametrize("M,K,N,relu", [(1000, 16, 16, 1), (65536, 64, 32, 1), (4173, 256, 200, 1), (300, 512, 520, 1),
                                         (2048, 128, 128, 0)])
def test_dgrad_bn_epilogue_and_apply(M, K, N, relu, deferred):
    """gz = gy . wt^T through the previous layer's BN + ReLU: dy, the two column sums and the final gy."""
    import ctypes
    lib = _lib.lib()
    g = torch.Generator(device=DEV).manual_seed(3 * M + N)
    Cv = N - 4 if N == 200 else N
    gy = torch.randn(M, K, device=DEV, generator=g).to(torch.bfloat16)
    wt = torch.zeros(N, K, device=DEV)
    wt[:Cv] = torch.randn(Cv, K, device=DEV, generator=g) / K ** 0.5
    wt = wt.to(torch.bfloat16)
    yprev = torch.zeros(M, N, device=DEV)
    yprev[:, :Cv] = torch.randn(M, Cv, device=DEV, generator=g) * 1.3 + 0.2
    yprev = yprev.to(torch.bfloat16)
    yf = yprev.float()[:, :Cv]
    mean, invstd = yf.mean(0).contiguous(), torch.rsqrt(yf.var(0, unbiased=False) + 1e-5).contiguous()
    gamma = (torch.rand(Cv, device=DEV, generator=g) + 0.5).contiguous()
    beta = (torch.randn(Cv, device=DEV, generator=g) * 0.3).contiguous()
    dy = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    sums = torch.full((3, N), float("nan"), device=DEV)
    work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=DEV)
    tick = ops._tickets(torch.device(DEV))
    gparts = torch.full((lib.pcb_gemm_max_groups(), 3, N), float("nan"), device=DEV) if deferred else None
    groups = ctypes.c_int(0)
    ops._call("pcb_dgrad_bn_rows_bf16", gy.device, gy.data_ptr(), gy.stride(0), wt.data_ptr(), wt.stride(0), M, N, N, K,
              yprev.data_ptr(), yprev.stride(0), mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), Cv,
              relu, dy.data_ptr(), dy.stride(0), None if deferred else sums.data_ptr(), work.data_ptr(), tick.data_ptr(),
              gparts.data_ptr() if deferred else None, ctypes.byref(groups))
    torch.cuda.synchronize()
    assert int(tick.abs().sum()) == 0
    gz = gy.float() @ wt.float().t()
    yh = (yf - mean) * invstd
    z = yh * gamma + beta
    mask = (z > 0) if relu else torch.ones_like(z, dtype=torch.bool)
    ref_dy = torch.zeros(M, N, device=DEV)
    ref_dy[:, :Cv] = gz[:, :Cv] * mask
    # elements whose activation sits within rounding of zero may be masked either way
    edge = torch.zeros(M, N, device=DEV, dtype=torch.bool)
    edge[:, :Cv] = z.abs() < 1e-5 * (1 + yf.abs())
    d = dy.float()
    tol = ref_dy.abs() * 2.0 ** -7 + 1e-5 * float(gz.abs().max())
    assert not (((d - ref_dy).abs() > tol) & ~edge).any()
    dq = d[:, :Cv]
    # elementwise half, in place (deferred: it also adds up the group partials into `sums`)
    out = dy.clone()
    ops._call("pcb_bn_bwd_apply_rows", gy.device, out.data_ptr(), yprev.data_ptr(), 1, M, N, Cv, mean.data_ptr(),
              invstd.data_ptr(), gamma.data_ptr(), sums.data_ptr(), out.data_ptr(), gparts.data_ptr() if deferred else None,
              groups.value)
    torch.cuda.synchronize()
    assert torch.allclose(sums[0, :Cv], dq.sum(0), rtol=1e-4, atol=1e-4 * float(dq.abs().sum(0).max()))
    assert torch.allclose(sums[1, :Cv], (dq * yh).sum(0), rtol=1e-4, atol=1e-4 * float((dq * yh).abs().sum(0).max()))
    assert float(sums[2].abs().max()) == 0.0
    if Cv < N:
        assert float(sums[:, Cv:].abs().max()) == 0.0 and float(d[:, Cv:].abs().max()) == 0.0
    ref = gamma * invstd * (dq - sums[0, :Cv] / M - yh * sums[1, :Cv] / M)
    o = out.float()
    assert torch.allclose(o[:, :Cv], ref, rtol=2.0 ** -7, atol=2.0 ** -8 * float(ref.abs().max()))
    if Cv < N:
        assert float(o[:, Cv:].abs().max()) == 0.0


@pytest.mark.parametrize("M,C,Cv,pool_k", [(4096, 64, 64, 1), (8192, 200, 196, 32), (1024, 512, 512, 16), (300, 8, 5, 1),
                                            (4096, 512, 512, 32), (16384, 256, 256, 16), (2 * 20 * 3, 32, 32, 20),
                                            (65536, 32, 32, 16), (33 * 7, 16, 12, 33), (128 * 128, 128, 128, 128)])
def test_bn_apply_rows(M, C, Cv, pool_k):
    g = torch.Generator(device=DEV).manual_seed(M + C)
    y = torch.zeros(M, C, device=DEV)
    y[:, :Cv] = torch.randn(M, Cv, device=DEV, generator=g)
    y = y.to(torch.bfloat16)
    mean = torch.randn(C, device=DEV, generator=g) * 0.2
    invstd = torch.rand(C, device=DEV, generator=g) + 0.5
    gamma = torch.rand(Cv, device=DEV, generator=g) + 0.5
    beta = torch.randn(Cv, device=DEV, generator=g) * 0.3
    wide = torch.full((M // pool_k, C + 16), 7.0, device=DEV, dtype=torch.bfloat16)
    out = wide[:, 8:8 + C]
    am = torch.empty(M // pool_k, C, device=DEV, dtype=torch.uint8) if pool_k > 1 else None
    ymax = torch.full((M // pool_k, C), 3.0, device=DEV, dtype=torch.bfloat16) if pool_k > 1 else None
    ops._call("pcb_bn_apply_rows", y.device, y.data_ptr(), 1, M, C, Cv, pool_k, mean.data_ptr(), invstd.data_ptr(),
              gamma.data_ptr(), beta.data_ptr(), 1, out.data_ptr(), out.stride(0), am.data_ptr() if am is not None else None,
              None, None, 0.0, None, None, None, 0, 1e-5, ymax.data_ptr() if ymax is not None else None)
    zfull = torch.relu((y.float()[:, :Cv] - mean[:Cv]) * invstd[:Cv] * gamma + beta)
    z = zfull
    if pool_k > 1:
        z, idx = zfull.view(-1, pool_k, Cv).max(dim=1)
        # the recorded argmax points at a row holding the maximum and is the FIRST such row (torch.max's tie rule on
        # the kernel's own values: ReLU produces many exact ties at 0)
        a = am[:, :Cv].long()
        zk = torch.relu(torch.addcmul((beta - mean[:Cv] * (invstd[:Cv] * gamma)), y.float()[:, :Cv], invstd[:Cv] * gamma))
        zk = zk.view(-1, pool_k, Cv)
        picked = zk.gather(1, a[:, None, :]).squeeze(1)
        assert torch.equal(picked, zk.max(dim=1)[0])
        first = (zk == zk.max(dim=1, keepdim=True)[0]).float().argmax(dim=1)
        assert torch.equal(a, first)
        # ymax: the stored pre-activation of the winning row, bit for bit
        yw = y[:, :Cv].reshape(-1, pool_k, Cv).gather(1, a[:, None, :]).squeeze(1)
        assert torch.equal(ymax[:, :Cv], yw)
        # backward of the pooled layer in two ordinary launches == the cooperative kernel up to summation order
        lib = _lib.lib()
        gz_w = torch.randn(M // pool_k, C + 8, device=DEV, generator=g).to(torch.bfloat16)
        gz = gz_w[:, :C]
        gam, bet = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        gam[:Cv], bet[:Cv] = gamma, beta
        res = []
        for entry in ("pcb_bn_bwd_rows", "pcb_bn_pool_bwd_rows"):
            work = torch.full((int(lib.pcb_bn_work_floats(C)),), float("nan"), device=DEV)
            gy = torch.full_like(y, 5.0)
            if entry == "pcb_bn_bwd_rows":
                ops._call(entry, y.device, gz.data_ptr(), gz.stride(0), y.data_ptr(), am.data_ptr(), 1, M, C, Cv, pool_k,
                          mean.data_ptr(), invstd.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1, work.data_ptr(), gy.data_ptr())
            else:
                ops._call(entry, y.device, gz.data_ptr(), gz.stride(0), ymax.data_ptr(), y.data_ptr(), am.data_ptr(), 1, M, C,
                          Cv, pool_k, mean.data_ptr(), invstd.data_ptr(), gam.data_ptr(), bet.data_ptr(), 1, work.data_ptr(),
                          gy.data_ptr())
            torch.cuda.synchronize()
            res.append((work[:3 * C].view(3, C).clone(), gy.float()))
        (s_a, gy_a), (s_b, gy_b) = res
        scale = float(s_a[:2, :Cv].abs().max()) + 1e-6
        assert torch.allclose(s_a[:2, :Cv], s_b[:2, :Cv], rtol=1e-4, atol=1e-5 * scale)
        assert float(s_b[2].abs().max()) == 0.0 and float(s_b[:, Cv:].abs().max() if Cv < C else 0.0) == 0.0
        assert torch.allclose(gy_a, gy_b, rtol=2.0 ** -7, atol=2.0 ** -8 * float(gy_a.abs().max()) + 1e-6)
    o = out.float()
    assert torch.allclose(o[:, :Cv], z, rtol=2.0 ** -7, atol=1e-6)
    assert float(o[:, Cv:].abs().max()) == 0.0 if Cv < C else True
    assert float(wide[:, :8].float().min()) == 7.0 and float(wide[:, 8 + C:].float().min()) == 7.0


class _Ref(nn.Module):
    def __init__(self, widths, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.convs = nn.ModuleList(nn.Conv1d(a, b, 1) for a, b in zip(widths[:-1], widths[1:]))
        self.bns = nn.ModuleList(nn.BatchNorm1d(b) for b in widths[1:])
        with torch.no_grad():
            for bn in self.bns:
                bn.weight.uniform_(0.5, 1.5)
                bn.bias.normal_(0, 0.2)


@pytest.mark.parametrize("widths,M,pool_k", [([16, 32, 32, 64], 16384, 32), ([104, 64, 96, 128], 8192, 16),
                                              ([264, 128, 196, 256], 2048, 32), ([128, 128], 4096, 1),
                                              ([1536, 256, 256], 1024, 1)])
def test_fused_mlp_matches_pytorch_training_mlp(widths, M, pool_k, monkeypatch):
    """Forward, input gradient and parameter gradients of the fused training MLP (a) against the round-1 composition
    with the same bf16 rounding points (library GEMM + cooperative BN row kernels, themselves checked against PyTorch
    in test_gpu_bn_rows.py): only summation order differs; (b) against conv1d / batch_norm / relu / max in fp32 on
    the bf16-rounded weights, where a bf16 rounding can hand the max-pool to another neighbour (loose bound)."""
    from pointcloud_bridge_b200.partsize import pointnet_util as pu
    nets = [_Ref(widths, 5).to(DEV).train() for _ in range(3)]
    with torch.no_grad():
        for net in nets:
            for m in net.convs:
                m.weight.copy_(m.weight.to(torch.bfloat16).float())
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(M, widths[0], device=DEV, generator=g).to(torch.bfloat16)
    xs = [x.clone().requires_grad_(True), x.clone().requires_grad_(True), x.float().clone().requires_grad_(True)]
    monkeypatch.setattr(ops, "_OWN_GEMM", True)
    assert ops.mlp_rows_fused_supported(xs[0], nets[0].convs, nets[0].bns, pool_k)
    out = ops.mlp_rows_fused(xs[0], nets[0].convs, nets[0].bns, pool_k)[:, :widths[-1]]
    monkeypatch.setattr(ops, "_OWN_GEMM", False)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        old = pu.mlp_rows(xs[1], nets[1].convs, nets[1].bns, pool_k)
    monkeypatch.undo()
    h = xs[2]
    for conv, bn in zip(nets[2].convs, nets[2].bns):
        h = F.relu(F.batch_norm(F.linear(h, conv.weight.flatten(1), conv.bias), bn.running_mean, bn.running_var, bn.weight,
                                bn.bias, True, bn.momentum, bn.eps))
    if pool_k > 1:
        h = h.view(-1, pool_k, h.shape[-1]).max(dim=1)[0]
    rel = lambda a, b: float((a.detach().float() - b.detach().float()).norm() / (b.detach().float().norm() + 1e-12))
    print("forward: vs round-1 path %.2e, vs fp32 %.2e" % (rel(out, old), rel(out, h)))
    assert rel(out, old) < 5e-3 and rel(out, h) < 2e-2
    gout = torch.randn(h.shape, device=DEV, generator=g).to(torch.bfloat16)
    out.backward(gout)
    old.backward(gout.to(old.dtype))
    h.backward(gout.float())
    loose = 0.35 if pool_k > 1 else 8e-2
    print("input gradient: vs round-1 path %.2e, vs fp32 %.2e" % (rel(xs[0].grad, xs[1].grad), rel(xs[0].grad, xs[2].grad)))
    assert rel(xs[0].grad, xs[1].grad) < 4e-2 and rel(xs[0].grad, xs[2].grad) < loose
    for (n1, p1), (_, p2), (_, p3) in zip(*[net.named_parameters() for net in nets]):
        if n1.startswith("convs") and n1.endswith("bias"):
            continue                                   # bias before a training-mode BN: zero gradient up to rounding
        assert p1.grad is not None, n1
        print(n1, "vs round-1 path %.2e, vs fp32 %.2e" % (rel(p1.grad, p2.grad), rel(p1.grad, p3.grad)))
        assert rel(p1.grad, p2.grad) < 4e-2 and rel(p1.grad, p3.grad) < loose, n1
    for b1, b3 in zip(nets[0].bns, nets[2].bns):
        assert torch.allclose(b1.running_mean, b3.running_mean, rtol=2e-2, atol=2e-3)
        assert torch.allclose(b1.running_var, b3.running_var, rtol=2e-2, atol=2e-3)
        assert int(b1.num_batches_tracked) == 1
