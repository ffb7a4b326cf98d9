"""Driver for an ncu capture of the fused BN row kernels on three layer shapes of the MSG train
step (bf16): `ncu --set full -k regex:bn_.*fused python tools/ncu_bn.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pointcloud_bridge_b200 import _lib  # noqa: E402

lib = _lib.lib()
dev = "cuda:0"
flush = torch.empty(64 * 1024 * 1024, device=dev)
for (M, C, K) in [(524288, 64, 32), (131072, 96, 1), (524288, 32, 1)]:
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
    for _ in range(2):
        flush.fill_(0)
        assert lib.pcb_bn_fwd_rows(y.data_ptr(), 1, M, C, C, K, None, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None, 1,
                                   stats[0].data_ptr(), stats[1].data_ptr(), out.data_ptr(), 0,
                                   am.data_ptr() if K > 1 else None, work.data_ptr(), st) == 0
        flush.fill_(0)
        assert lib.pcb_bn_bwd_rows(gz.data_ptr(), 0, y.data_ptr(), am.data_ptr() if K > 1 else None, 1, M, C, C, K,
                                   stats[0].data_ptr(), stats[1].data_ptr(), g.data_ptr(), b.data_ptr(), 1,
                                   work.data_ptr(), gy.data_ptr(), st) == 0
    torch.cuda.synchronize()
print("ok")
