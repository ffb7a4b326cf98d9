"""Shape sweep of the tcgen05 training GEMM (csrc/gemm_rows.cu) against an fp32 matmul with a mismatch DIAGNOSIS per
shape (which rows / columns / K ranges are wrong) -- the first thing to run after touching operand layouts or
descriptors.  `python tools/check_gemm.py [plain|stats|dgrad]...`; stops at the first CUDA error."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
SHAPES = [(128, 64, 64), (128, 16, 16), (128, 32, 32), (128, 8, 8), (129, 8, 8), (1000, 16, 16), (4096, 32, 64),
          (4096, 64, 128), (4096, 128, 128), (300, 72, 24), (4173, 104, 200), (1024, 1536, 256), (300, 520, 512),
          (8192, 264, 128), (128, 64, 1024), (40000, 200, 256), (65536, 32, 64), (262144, 16, 16), (65536, 96, 128), (524288, 32, 64), (131072, 104, 64)]
modes = sys.argv[1:] or ["plain", "stats", "dgrad"]


def diagnose(got, ref, what):
    got, ref = got.float(), ref.float()
    tol = ref.abs() * 2.0 ** -7 + 2e-3 * ref.abs().max().clamp_min(1e-6) * 2.0 ** -7 + 1e-6
    bad = (got - ref).abs() > tol
    if not bad.any():
        return True
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    print("   %s MISMATCH: %d of %d elements; rows %d..%d (%d), cols %d..%d (%d); max err %.3g (ref max %.3g); nan %d" % (
        what, int(bad.sum()), bad.numel(), int(rows[0]), int(rows[-1]), rows.numel(), int(cols[0]), int(cols[-1]),
        cols.numel(), float((got - ref).abs().nan_to_num(1e9).max()), float(ref.abs().max()), int(got.isnan().sum())))
    r = int(rows[0])
    print("     row %d got %s" % (r, [round(v, 3) for v in got[r, :12].tolist()]))
    print("     row %d ref %s" % (r, [round(v, 3) for v in ref[r, :12].tolist()]))
    return False


ok_all = True
for M, K, N in SHAPES:
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(torch.bfloat16)
    ref = x.float() @ w.float().t()
    line = "M=%d K=%d N=%d:" % (M, K, N)
    if "plain" in modes:
        y = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        ops.gemm_rows(x, w, N, out=y)
        torch.cuda.synchronize()
        ok = diagnose(y, ref, "plain")
        line += " plain %s" % ("ok" if ok else "BAD")
        ok_all &= ok
    if "stats" in modes:
        y = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        stats = torch.full((3, N), float("nan"), device=dev)
        work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=dev)
        tick = ops._tickets(dev)
        ops._call("pcb_linear_bn_stats_rows_bf16", dev, x.data_ptr(), K, w.data_ptr(), K, M, N, N, K, y.data_ptr(), N, N,
                  1e-5, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), work.data_ptr(), tick.data_ptr(),
                  None, None)
        torch.cuda.synchronize()
        ok = diagnose(y, ref, "stats-y")
        yf = y.float()
        mean, var = yf.mean(0), yf.var(0, unbiased=False)
        ok_s = bool(torch.allclose(stats[0], mean, rtol=1e-4, atol=1e-5 * float(yf.abs().max()))) and \
            bool(torch.allclose(stats[2], var, rtol=2e-4, atol=1e-7)) and int(tick.abs().sum()) == 0
        if not ok_s:
            print("   stats MISMATCH: mean err %.3g var rel err %.3g tick %d" % (
                float((stats[0] - mean).abs().max()), float(((stats[2] - var).abs() / var.clamp_min(1e-12)).max()),
                int(tick.abs().sum())))
        line += " stats %s" % ("ok" if ok and ok_s else "BAD")
        ok_all &= ok and ok_s
    if "dgrad" in modes:
        # gz [M, N] = gy [M, K] . wt [N, K]^T through BN + ReLU of yprev [M, N]
        yprev = (torch.randn(M, N, device=dev, generator=g) * 1.3 + 0.2).to(torch.bfloat16)
        yf = yprev.float()
        mean, invstd = yf.mean(0).contiguous(), torch.rsqrt(yf.var(0, unbiased=False) + 1e-5).contiguous()
        gamma = (torch.rand(N, device=dev, generator=g) + 0.5).contiguous()
        beta = (torch.randn(N, device=dev, generator=g) * 0.3).contiguous()
        dy = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        sums = torch.full((3, N), float("nan"), device=dev)
        work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=dev)
        tick = ops._tickets(dev)
        ops._call("pcb_dgrad_bn_rows_bf16", dev, x.data_ptr(), K, w.data_ptr(), K, M, N, N, K, yprev.data_ptr(), N,
                  mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), N, 1, dy.data_ptr(), N,
                  sums.data_ptr(), work.data_ptr(), tick.data_ptr(), None, None)
        torch.cuda.synchronize()
        yh = (yf - mean) * invstd
        z = yh * gamma + beta
        edge = z.abs() < 1e-5 * (1 + yf.abs())
        ref_dy = ref * (z > 0)
        d = dy.float()
        ok = diagnose(torch.where(edge, ref_dy, d), ref_dy, "dgrad-dy")
        ok_s = bool(torch.allclose(sums[0], d.sum(0), rtol=1e-4, atol=1e-4 * float(d.abs().sum(0).max()))) and \
            bool(torch.allclose(sums[1], (d * yh).sum(0), rtol=1e-4, atol=1e-4 * float((d * yh).abs().sum(0).max())))
        if not ok_s:
            print("   sums MISMATCH: %.3g %.3g" % (float((sums[0] - d.sum(0)).abs().max()),
                                                   float((sums[1] - (d * yh).sum(0)).abs().max())))
        line += " dgrad %s" % ("ok" if ok and ok_s else "BAD")
        ok_all &= ok and ok_s
    print(line, flush=True)
print("ALL OK" if ok_all else "FAILURES")
sys.exit(0 if ok_all else 1)
