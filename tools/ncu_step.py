"""Driver for ncu captures of ONE MSG train step (bench.py's workload) between cudaProfilerStart/Stop.

    python tools/ncu_step.py            # the step as benchmarked: CUDA-graph replay + prefetch pipeline (one replay and
                                        # one index chain of the next batch fall into the capture)
    python tools/ncu_step.py eager      # eager launches, sampling inside the step (no graph, no prefetch)

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/step_launches.csv python tools/ncu_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointcloud_bridge_b200 import synthetic  # noqa: E402
from pointcloud_bridge_b200.engine import Trainer  # noqa: E402
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg  # noqa: E402

eager = len(sys.argv) > 1 and sys.argv[1] == "eager"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, N = 16, 4096
batches = []
for i in range(2):
    xyz, rgb, lab = synthetic.bridge_batch(100 + i, B, N)
    batches.append((torch.from_numpy(synthetic.sem_seg_input(xyz, rgb)).to(dev), torch.from_numpy(lab).to(dev)))
net = msg.get_model(5).to(dev).train()
tr = Trainer(net, amp=True, graph=not eager)
if eager:
    tr._use_chain = False
    for _ in range(2):
        tr.step(batches[0][0], labels=batches[0][1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step(batches[0][0], labels=batches[0][1])
else:
    tr.prefetch(batches[0][0], labels=batches[0][1])
    for i in range(6):                                   # 3 eager warm-up steps, capture, 2 replays
        tr.step_prefetched()
        tr.prefetch(batches[(i + 1) % 2][0], labels=batches[(i + 1) % 2][1])
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr.step_prefetched()
    tr.prefetch(batches[1][0], labels=batches[1][1])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")
