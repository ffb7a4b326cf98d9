"""Timing of the fused BN+ReLU(+max) row kernels on the layer shapes of the MSG train step
(bf16 activations).  Back-to-back launches (20 per measurement) so that launch gaps do not count."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import _lib, ops  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def t(fn, n=20):
    """GPU time per call: n launches captured into one CUDA graph (no host launch gaps), with a
    256 MB L2 flush inside the graph before every call whose duration is measured separately."""
    warm = os.environ.get("MB_WARM") == "1"               # no L2 flush between calls
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def graph_ms(with_fn):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(n):
                    if not warm:
                        flush.fill_(0.0)
                    if with_fn:
                        fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            g.replay()
            e1.record(s)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    return graph_ms(True) - (0.0 if warm else graph_ms(False))


def main():
    dev = "cuda:0"
    lib = _lib.lib()
    for (M, C, K) in [(524288, 32, 1), (524288, 64, 32), (262144, 16, 1), (131072, 96, 1), (131072, 128, 32), (65536, 128, 1),
                      (32768, 256, 32), (8192, 512, 32), (1024, 256, 1)]:
        y = torch.randn(M, C, device=dev).to(torch.bfloat16)
        work = torch.zeros(lib.pcb_bn_work_floats(C), device=dev)
        stats = torch.zeros(2, C, device=dev)
        g = torch.ones(C, device=dev)
        b = torch.zeros(C, device=dev)
        out = torch.empty(M // K, C, device=dev, dtype=torch.bfloat16)
        am = torch.empty(M // K, C, device=dev, dtype=torch.uint8)
        gz = torch.randn(M // K, C, device=dev).to(torch.bfloat16)
        gy = torch.empty_like(y)
        nb = y.numel() * 2

        def fwd():
            return lib.pcb_bn_fwd_rows(y.data_ptr(), 1, M, C, C, K, None, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None, 1,
                                       stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(), 0,
                                       am.data_ptr() if K > 1 else None, work.data_ptr(), torch.cuda.current_stream().cuda_stream)

        def bwd():
            return lib.pcb_bn_bwd_rows(gz.data_ptr(), 0, y.data_ptr(), am.data_ptr() if K > 1 else None, 1, M, C, C, K,
                                       stats[0].data_ptr(), stats[1].data_ptr(), g.data_ptr(), b.data_ptr(), 1,
                                       work.data_ptr(), gy.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert fwd() == 0 and bwd() == 0
        mf, mb = t(fwd), t(bwd)
        bf = nb + out.numel() * 2 + (am.numel() if K > 1 else 0)
        bb = 2 * nb + gz.numel() * 2 + (am.numel() if K > 1 else 0)
        print(json.dumps({"M": M, "C": C, "pool_k": K, "MB": round(nb / 1e6, 1),
                          "fwd_us": round(mf * 1e3, 1), "fwd_frac": round(bf / mf / 1e6 / HBM, 3),
                          "bwd_us": round(mb * 1e3, 1), "bwd_frac": round(bb / mb / 1e6 / HBM, 3)}), flush=True)


if __name__ == "__main__":
    main()
