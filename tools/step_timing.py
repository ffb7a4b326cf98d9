"""Where the captured MSG train step spends its time when nothing else runs: the step graph replayed alone (no index
chain of the next batch on the side stream, no input copies), next to the pipelined loop that bench.py times.

    [PCB_OWN_GEMM=0] python tools/step_timing.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, N = 16, 4096
batches = []
for i in range(2):
    xyz, rgb, lab = synthetic.bridge_batch(100 + i, B, N)
    batches.append((torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev), torch.from_numpy(lab).to(dev)))
net = msg.get_model(5).to(dev).train()
tr = Trainer(net, amp=True, graph=True)
tr.prefetch(batches[0][0], labels=batches[0][1])
for i in range(8):
    tr.step_prefetched()
    tr.prefetch(batches[(i + 1) % 2][0], labels=batches[(i + 1) % 2][1])
torch.cuda.synchronize()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def timed(fn, n=20, do_flush=True):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in evs:
        if do_flush:
            flush.fill_(0.0)
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
    return t[len(t) // 2], t[0], t[-1]


print("graph replay alone (L2 flushed)   median %.3f  min %.3f  max %.3f ms" % timed(lambda: tr._g.replay()))
print("graph replay alone (warm L2)      median %.3f  min %.3f  max %.3f ms" % timed(lambda: tr._g.replay(), do_flush=False))


def pipelined():
    tr.step_prefetched()
    tr.prefetch(batches[1][0], labels=batches[1][1])


print("pipelined step (as bench.py)      median %.3f  min %.3f  max %.3f ms" % timed(pipelined))
print("kernels of ours per replay:", tr.kernel_launches_per_replay, " own GEMM:", os.environ.get("PCB_OWN_GEMM", "0"))
