"""Role timeline of CTA (0, 0) of the tcgen05 GEMM (needs a -DPCB_GEMM_TRACE build:
python tools/build_variant.py gemmtrace gemm_rows.cu -DPCB_GEMM_TRACE; PCB_LIB_PATH=.../variants/gemmtrace.so)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import _lib, ops  # noqa: E402

lib = _lib.lib()
dev = torch.device("cuda:0")
buf = np.zeros(3 * 512, dtype=np.int64)
SHAPES = [(524288, 32, 64), (131072, 96, 128), (4096, 512, 256)]
which = sys.argv[1] if len(sys.argv) > 1 else "plain"
for (M, K, N) in SHAPES:
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(3, N, device=dev)
    work = torch.empty(max(int(lib.pcb_gemm_work_floats(M, N, K)), 1), device=dev)
    tick = ops._tickets(dev)
    for _ in range(3):
        if which == "plain":
            ops.gemm_rows(x, w, N, out=y)
        else:
            ops._call("pcb_linear_bn_stats_rows_bf16", dev, x.data_ptr(), K, w.data_ptr(), K, M, N, N, K, y.data_ptr(), N, N,
                      1e-5, stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), work.data_ptr(), tick.data_ptr())
    torch.cuda.synchronize()
    assert lib.pcb_gemm_debug_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(3, 512).copy()
    t0 = min(int(t[0, 0]), int(t[2, 0]))
    us = lambda v: (int(v) - t0) / 1.9e3
    print(f"--- {which} M={M} K={K} N={N}")
    print(" producer (acquire, issued) per slab:", [(round(us(t[0, 2 * i]), 2), round(us(t[0, 2 * i + 1]), 2)) for i in range(10)])
    print(" mma (full seen, committed) per slab:", [(round(us(t[1, 2 * i]), 2), round(us(t[1, 2 * i + 1]), 2)) for i in range(10)])
    print(" epilogue (enter, acc full, tile in smem, stored) per tile:",
          [tuple(round(us(t[2, 4 * i + j]), 2) for j in range(4)) for i in range(16) if t[2, 4 * i] > t0 - 1])
    buf[:] = 0
