"""Per-layer timing of the tcgen05 training GEMMs (csrc/gemm_rows.cu) on the layer shapes of the MSG train step
(B = 16), next to torch.mm (cuBLAS) on the same operands.  CUDA events over `reps` back-to-back launches."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
B = 16
LAYERS = [  # (name, M, K, N)
    ("sa1.0 l1", B * 1024 * 16, 16, 16), ("sa1.0 l3", B * 1024 * 16, 16, 32), ("sa1.1 l1", B * 1024 * 32, 16, 32),
    ("sa1.1 l2", B * 1024 * 32, 32, 32), ("sa1.1 l3", B * 1024 * 32, 32, 64), ("sa2.0 l1", B * 256 * 16, 104, 64),
    ("sa2.1 l1", B * 256 * 32, 104, 64), ("sa2.1 l2", B * 256 * 32, 64, 96), ("sa2.1 l3", B * 256 * 32, 96, 128),
    ("sa3.1 l1", B * 64 * 32, 264, 128), ("sa3.1 l2", B * 64 * 32, 128, 200), ("sa3.1 l3", B * 64 * 32, 200, 256),
    ("sa4.1 l1", B * 16 * 32, 520, 256), ("sa4.1 l2", B * 16 * 32, 256, 384), ("sa4.1 l3", B * 16 * 32, 384, 512),
    ("fp4 l1", B * 64, 1536, 256), ("fp3 l1", B * 256, 512, 256), ("fp2 l1", B * 1024, 352, 256),
    ("fp1 l1", B * 4096, 128, 128), ("head conv2", B * 4096, 128, 8),
]
reps = int(os.environ.get("REPS", "20"))
only = os.environ.get("ONLY")                  # e.g. ONLY="sa1.1 l3": one layer (ncu runs)
if only:
    LAYERS = [l for l in LAYERS if l[0] == only]
nograph = os.environ.get("NOGRAPH", "0") == "1"


def timed(fn):
    """GPU time per launch: `reps` launches captured in one CUDA graph (no host launch overhead in the number)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if nograph:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3          # us


print("%-11s %8s %5s %5s | %8s %8s %8s | %8s %8s | %8s %8s" % ("layer", "M", "K", "N", "plain", "stats", "cuBLAS", "GB/s st", "frac",
                                                                 "dgradBN", "cuBLASd"))
tot = {"plain": 0.0, "stats": 0.0, "mm": 0.0, "dg": 0.0, "mmd": 0.0}
for name, M, K, N in LAYERS:
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    wt = w.t().contiguous()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(3, N, device=dev)
    work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=dev)
    tick = ops._tickets(dev)
    t_plain = timed(lambda: ops.gemm_rows(x, w, N, out=y))
    # as the training path runs it: the GEMM stops after its first fold level (group partials), the elementwise kernel
    # that follows merges the groups (PCB_GEMM_FULL_FOLD=1: the GEMM folds to the end itself)
    full = os.environ.get("PCB_GEMM_FULL_FOLD", "0") == "1"
    gparts = torch.empty(lib.pcb_gemm_max_groups(), 3, N, device=dev)
    groups = ctypes.c_int(0)
    t_stats = timed(lambda: ops._call("pcb_linear_bn_stats_rows_bf16", dev, x.data_ptr(), K, w.data_ptr(), K, M, N, N, K,
                                      y.data_ptr(), N, N, 1e-5, stats[0].data_ptr(), stats[1].data_ptr(),
                                      stats[2].data_ptr(), work.data_ptr(), tick.data_ptr(),
                                      None if full else gparts.data_ptr(), ctypes.byref(groups)))
    t_mm = timed(lambda: torch.mm(x, w.t(), out=y))
    # data gradient of this layer through the previous layer's BN: gz [M, K] = gy [M, N] . W; y_prev [M, K]
    t_dg = t_mmd = float("nan")
    if K % 8 == 0 and K >= 16:
        gy = torch.randn(M, N, device=dev).to(torch.bfloat16)
        yprev = torch.randn(M, K, device=dev).to(torch.bfloat16)
        dy = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        mean, invstd = torch.zeros(K, device=dev), torch.ones(K, device=dev)
        gamma, beta = torch.ones(K, device=dev), torch.zeros(K, device=dev)
        sums = torch.empty(3, K, device=dev)
        work2 = torch.empty(max(int(lib.pcb_gemm_work_floats(M, K, N)), 1), device=dev)
        gparts2 = torch.empty(lib.pcb_gemm_max_groups(), 3, K, device=dev)
        t_dg = timed(lambda: ops._call("pcb_dgrad_bn_rows_bf16", dev, gy.data_ptr(), N, wt.data_ptr(), N, M, K, K, N,
                                       yprev.data_ptr(), K, mean.data_ptr(), invstd.data_ptr(), gamma.data_ptr(),
                                       beta.data_ptr(), K, 1, dy.data_ptr(), K, sums.data_ptr(), work2.data_ptr(),
                                       tick.data_ptr(), None if full else gparts2.data_ptr(), ctypes.byref(groups)))
        t_mmd = timed(lambda: torch.mm(gy, w, out=dy))
    gbs = 2 * M * (K + N) / t_stats / 1e3
    print("%-11s %8d %5d %5d | %8.1f %8.1f %8.1f | %8.0f %8.3f | %8.1f %8.1f" % (name, M, K, N, t_plain, t_stats, t_mm, gbs,
                                                                                 gbs / 6540, t_dg, t_mmd))
    tot["plain"] += t_plain
    tot["stats"] += t_stats
    tot["mm"] += t_mm
    if t_dg == t_dg:
        tot["dg"] += t_dg
        tot["mmd"] += t_mmd
print("sum (us):", {k: round(v, 1) for k, v in tot.items()})
