"""Per-shape timing of the row weight-gradient kernel (csrc/wgrad.cu) against the batched-GEMM + reduction path
it replaces, on the layer shapes of the PointNet++ MSG train step (B=16).  CUDA events, 20 launches back to back."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import ops  # noqa: E402

dev = "cuda:0"
SHAPES = [(524288, 64, 32), (524288, 32, 16), (262144, 32, 16), (131072, 128, 96), (131072, 64, 104), (65536, 128, 128),
          (32768, 128, 264), (16384, 256, 352), (8192, 512, 384), (4096, 256, 520), (1024, 256, 1536)]
only = [int(a) for a in sys.argv[1:]]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for i, (M, N, K) in enumerate(SHAPES):
    if only and i not in only:
        continue
    gy = torch.randn(M, N, device=dev).to(torch.bfloat16)
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    out = torch.zeros(N, K, device=dev)
    t_own = timed(lambda: ops.wgrad_rows(gy, x, K, out=out))

    def lib_path():
        c = 2048
        if M % c == 0 and M // c >= 8:
            p = M // c
            return torch.bmm(gy.view(p, c, -1).transpose(1, 2), x.view(p, c, -1)).sum(dim=0, dtype=torch.float32)
        return torch.mm(gy.t(), x).float()
    t_lib = timed(lib_path)
    nbytes = 2 * M * (N + K)
    print(json.dumps({"M": M, "N": N, "K": K, "own_us": round(t_own, 1), "own_GBps": round(nbytes / t_own / 1e3, 0),
                      "lib_us": round(t_lib, 1), "MB": round(nbytes / 1e6, 1)}))
