"""Per-CTA phase timeline of the fused BN kernels (needs a -DPCB_BN_TRACE build, see tools/build_variant.py)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
buf = np.zeros(8 * 1024, dtype=np.uint64)


def report(tag, ncta):
    assert lib.pcb_bn_debug_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    t = buf.reshape(1024, 8)[:ncta, :6].astype(np.int64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1e3
    names = ["start", "p1_end", "sync1", "fold_end", "sync2", "end"]
    print(tag, "ctas", ncta)
    for i, n in enumerate(names):
        c = t[:, i]
        print(f"   {n:9s} min {c.min():7.2f} med {np.median(c):7.2f} p90 {np.percentile(c, 90):7.2f} max {c.max():7.2f} us")


SHAPES = [(524288, 32, 1), (524288, 64, 32), (131072, 96, 1), (65536, 128, 1), (16384, 256, 1), (4096, 256, 1),
          (8192, 512, 32), (1024, 256, 1)]
for (M, C, K) in SHAPES:
    y = torch.randn(M, C, device=dev).to(torch.bfloat16)
    work = torch.zeros(lib.pcb_bn_work_floats(C), device=dev)
    stats = torch.zeros(2, C, device=dev)
    g = torch.ones(C, device=dev)
    b = torch.zeros(C, device=dev)
    out = torch.empty(M // K, C, device=dev, dtype=torch.bfloat16)
    am = torch.zeros(M // K, C, device=dev, dtype=torch.uint8)
    gz = torch.randn(M // K, C, device=dev).to(torch.bfloat16)
    gy = torch.empty_like(y)
    st = torch.cuda.current_stream().cuda_stream
    for it in range(3):
        buf[:] = 0
        lib.pcb_bn_fwd_rows(y.data_ptr(), 1, M, C, C, K, None, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None, 1,
                            stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(), 0, am.data_ptr() if K > 1 else None,
                            work.data_ptr(), st)
        torch.cuda.synchronize()
        if it == 2:
            assert lib.pcb_bn_debug_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
            n = int((buf.reshape(1024, 8)[:, 0] > 0).sum())
            report(f"fwd M={M} C={C} K={K}", n)
        lib.pcb_bn_bwd_rows(gz.data_ptr(), 0, y.data_ptr(), am.data_ptr() if K > 1 else None, 1, M, C, C, K,
                            stats[0].data_ptr(), stats[1].data_ptr(), g.data_ptr(), b.data_ptr(), 1,
                            work.data_ptr(), gy.data_ptr(), st)
        torch.cuda.synchronize()
        if it == 2:
            assert lib.pcb_bn_debug_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
            n = int((buf.reshape(1024, 8)[:, 0] > 0).sum())
            report(f"bwd M={M} C={C} K={K}", n)

    # back-to-back launch time (warm L2 for the small shapes), CUDA events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("fwd", lambda: lib.pcb_bn_fwd_rows(y.data_ptr(), 1, M, C, C, K, None, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None, 1,
                                                        stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(), 0,
                                                        am.data_ptr() if K > 1 else None, work.data_ptr(), st)),
                     ("bwd", lambda: lib.pcb_bn_bwd_rows(gz.data_ptr(), 0, y.data_ptr(), am.data_ptr() if K > 1 else None, 1, M, C, C, K,
                                                        stats[0].data_ptr(), stats[1].data_ptr(), g.data_ptr(), b.data_ptr(), 1,
                                                        work.data_ptr(), gy.data_ptr(), st))):
        for _ in range(5):
            fn()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"   {name} back-to-back: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us/launch")
